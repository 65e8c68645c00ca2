"""Optimisation history as a rooted forest (mirror of morl/opt_graph.py:8-27).

Node i keeps the unit-L2 scalarisation weight it was trained with, the objectives it reached,
the gain over its parent and the parent / children links. ``prediction_guided_selection`` reads
``weights``, ``objs``, ``delta_objs`` and ``succ``; morl/morl.py:57-58,115 calls ``insert``.
"""
import numpy as np


class OptGraph:
    def __init__(self):
        self.weights = []      # unit L2 norm
        self.objs = []
        self.delta_objs = []   # objs - objs[prev]; zeros for roots
        self.prev = []
        self.succ = []
        self._flat_n = 0       # nodes already mirrored in the dense arrays of `flat()`
        self._flat = None

    def insert(self, weights, objs, prev):
        """Append a node and return its id (opt_graph.py:16-27). `weights` may be a torch tensor
        (roots, morl.py:58) or a numpy array (children, morl.py:111-115); it is stored divided
        by its L2 norm, in its own type, as the reference does."""
        node = len(self.objs)
        w = _clone(weights)
        self.weights.append(w / np.linalg.norm(weights))
        o = _clone(objs)
        self.objs.append(o)
        self.prev.append(prev)
        if prev == -1:
            self.delta_objs.append(np.zeros_like(o))
        else:
            self.delta_objs.append(o - self.objs[prev])
            self.succ[prev].append(node)
        self.succ.append([])
        return node

    # ---- dense views used by the device path (float64) ----
    def flat(self):
        """Dense float64 mirror of the node lists, extended incrementally (nodes are only ever appended):
        -> (weights [n,M], objs [n,M], delta_objs [n,M], prev [n] int64). The arrays are over-allocated and sliced;
        callers must not write into them."""
        n = len(self.objs)
        if n == 0:
            return np.zeros((0, 0)), np.zeros((0, 0)), np.zeros((0, 0)), np.zeros(0, dtype=np.int64)
        if getattr(self, '_flat', None) is None or self._flat_n > n:      # (graphs unpickled from before the cache existed)
            self._flat, self._flat_n = None, 0
        M = len(np.asarray(self.objs[0]).reshape(-1))
        if self._flat is None or self._flat[0].shape[0] < n:
            cap = max(64, 2 * n)
            new = [np.zeros((cap, M)), np.zeros((cap, M)), np.zeros((cap, M)), np.zeros(cap, dtype=np.int64)]
            if self._flat is not None:
                for a, b in zip(new, self._flat):
                    a[:self._flat_n] = b[:self._flat_n]
            self._flat = new
        W, O, D, P = self._flat
        for i in range(self._flat_n, n):
            W[i] = np.asarray(self.weights[i], dtype=np.float64)
            O[i] = np.asarray(self.objs[i], dtype=np.float64)
            D[i] = np.asarray(self.delta_objs[i], dtype=np.float64)
            P[i] = self.prev[i]
        self._flat_n = n
        return W[:n], O[:n], D[:n], P[:n]

    def arrays(self):
        """-> (weights [n,M], objs [n,M], delta_objs [n,M]) as float64 numpy arrays."""
        f = lambda seq: np.array([np.asarray(v, dtype=np.float64) for v in seq], dtype=np.float64)
        return f(self.weights), f(self.objs), f(self.delta_objs)


def _clone(v):
    return v.clone() if hasattr(v, "clone") else np.array(v, copy=True)
