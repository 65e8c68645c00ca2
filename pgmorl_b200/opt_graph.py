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
    def arrays(self):
        """-> (weights [n,M], objs [n,M], delta_objs [n,M]) as float64 numpy arrays."""
        f = lambda seq: np.array([np.asarray(v, dtype=np.float64) for v in seq], dtype=np.float64)
        return f(self.weights), f(self.objs), f(self.delta_objs)


def _clone(v):
    return v.clone() if hasattr(v, "clone") else np.array(v, copy=True)
