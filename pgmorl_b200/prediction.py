"""Hyperbolic prediction model of PG-MORL (mirror of predict_hyperbolic / collect_nearest_data,
morl/population_2d.py:12-118 and morl/population_3d.py:13-113), batched over the whole population.

Device side: the neighbourhood search in the opt-graph with the widening (threshold, sigma) schedule and the
gather of the training points of every model (csrc/k4_inputs.cu), then ALL fits of a selection call (n_pop x M
bounded robust least-squares problems) in one launch of K4 (csrc/k4_fit.cu).
Host side (numpy): the Gaussian point weights and the model evaluations -- the two places where the reference's
numbers come out of numpy's exp, which no device routine reproduces bit for bit -- as whole-population array
expressions built from the row-wise helpers of utils.py that give the bits of the scalar calls they replace.
Predictions are then objs + f(test weight)."""
import numpy as np

from . import kernels as K
from ._nvtx import rng as _nvtx
from .utils import pow2, rowdot


class GraphView:
    """Flat float64 arrays over an OptGraph. Edges (source node -> successor) in the data order of
    collect_nearest_data: by source node, then by successor (successors are appended in node order)."""

    def __init__(self, opt_graph):
        W, O, D, prev = opt_graph.flat()
        self.weights, self.objs = W, O
        child = np.nonzero(prev >= 0)[0]
        child = child[np.argsort(prev[child], kind="stable")]
        self.child, self.parent = child, prev[child]
        wc = W[child]
        self.edge_w = wc / wc.sum(axis=1, keepdims=True)        # successor weights normalised to sum 1 (population_2d.py:19)
        self.edge_dy = D[child]                                  # ... and their objective gains
        # successors of node i: child[estart[i]:estart[i + 1]]
        self.estart = np.searchsorted(self.parent, np.arange(len(O) + 1))

    def successors_of(self, nodes):
        """-> (member index [P], successor node [P]) for the successor lists of `nodes`, members in order."""
        nodes = np.asarray(nodes, dtype=np.int64)
        start, cnt = self.estart[nodes], self.estart[nodes + 1] - self.estart[nodes]
        member = np.repeat(np.arange(len(nodes)), cnt)
        offs = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        return member, self.child[np.repeat(start, cnt) + offs]


def gaussian_weights(view, node_ids, steps, source):
    """Point weights exp(-(dist / sigma)^2 / 2) of every listed edge (population_2d.py:88-96): dist = L2 norm of the
    source node's relative objective distance to the member's node, sigma = 0.03 doubled `steps` times.
    source [n, Kmax] (padding entries give meaningless values that K4 never reads)."""
    ok = view.objs[np.asarray(node_ids, dtype=np.int64)]
    with np.errstate(all="ignore"):
        q = np.abs(ok[:, None, :] - view.objs[source]) / np.abs(ok)[:, None, :]
        dist = np.sqrt(rowdot(q, q))
        sigma = np.ldexp(0.03, np.asarray(steps, dtype=np.int32))          # exact doublings
        return np.exp(-pow2(dist / sigma[:, None]) / 2.0)


def model(x, A, a, b, c):
    return A * (np.exp(a * (x - b)) - 1) / (np.exp(a * (x - b)) + 1) + c


_FIT_STREAMS = {}


def fit_stream():
    """The fit chain (front-end kernels, K4 and the copies around them) runs on its own CUDA stream: a launch issued early
    (`prefetch`) is then not held up by -- and does not hold up -- the small kernels and synchronising copies the caller
    issues on the current stream meanwhile (the archive's dominance filter). Inputs come from and results go to the host
    inside this stream, so there is no device-side dependency on any other stream."""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _FIT_STREAMS:
        _FIT_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _FIT_STREAMS[dev]


class FitRecord(dict):
    """Record of the most recent K4 launch (diagnostics / tests): theta, status, nfev, cost [F] as numpy arrays;
    the ragged inputs x, y, w (lists of 1-D arrays) and ub are copied back from the device on first access."""

    def __init__(self, lazy, **kw):
        super().__init__(**kw)
        self._lazy = lazy

    def __missing__(self, key):
        if key in ("x", "y", "w", "ub") and self._lazy is not None:
            lazy, self._lazy = self._lazy, None
            self.update(lazy())
            return self[key]
        raise KeyError(key)


def launch_fits(opt_graph, node_ids, obj_num, cap_threshold):
    """Build the training data of every node's model and launch all n x M fits (K4) without waiting for them.
    Returns a handle for `finish_predictions`; the caller may do host work that does not need the fits meanwhile."""
    import torch
    with torch.cuda.stream(fit_stream()):
        with _nvtx("selection.fit_inputs"):
            view = GraphView(opt_graph)
            node_ids = np.asarray(list(node_ids), dtype=np.int64)
            front = K.fit_inputs_launch(view.objs, view.parent, view.edge_w, view.edge_dy, node_ids, cap_threshold)
            coef = gaussian_weights(view, node_ids, front["steps"], front["source"]) if len(node_ids) else np.zeros((0, 1))
        with _nvtx("selection.k4_fits"):
            fits = K.fit_hyperbolic_launch_packed(front, coef)                  # all fits in one launch
    return dict(view=view, node_ids=node_ids, obj_num=obj_num, front=front, fits=fits, n_nodes=len(view.objs),
                cap_threshold=bool(cap_threshold))


def _ragged_inputs(handle, keep):
    """x, y, w, ub of the kept fits as the lists the scalar code used to hold (device -> host copy of the pack)."""
    import torch
    front = handle["front"]
    with torch.cuda.stream(fit_stream()):
        pack, ub, klen = front["pack"].cpu().numpy(), front["ub"].cpu().numpy(), front["klen"]
    M = handle["obj_num"]
    out = dict(x=[], y=[], w=[], ub=[])
    for f in keep:
        k = int(klen[f // M])
        out["x"].append(pack[0, f, :k].copy()); out["y"].append(pack[1, f, :k].copy())
        out["w"].append(pack[2, f, :k].copy()); out["ub"].append(ub[f].copy())
    return out


def finish_predictions(handle, tests, counts, zero_if_degenerate=False):
    """Second half of `predict_population`. tests [n, T, M] holds counts[i] test weights for node_ids[i] in its first
    rows (any positive scaling; normalised to sum 1 here, as population_2d.py:28-32); a member without test weights
    contributes neither predictions nor fit records (the reference never fits a sample without test weights).
    Returns (pred [n, T, M], fit record). `zero_if_degenerate`: the fork copy's fallback
    (WorkingMorl/morl/population_2d.py:112-117)."""
    import torch
    with torch.cuda.stream(fit_stream()):
        theta, status, nfev, cost = K.fit_hyperbolic_collect(handle["fits"])
    view, M = handle["view"], handle["obj_num"]
    n = len(handle["node_ids"])
    counts = np.asarray(counts, dtype=np.int64)
    tests = np.asarray(tests, dtype=np.float64).reshape(n, -1, M)
    with np.errstate(all="ignore"):
        tw = tests / tests.sum(axis=2, keepdims=True)              # row /= np.sum(row), population_2d.py:28-32 (same bits)
        th = theta.reshape(n, 1, M, 4)
        delta = model(tw, th[..., 0], th[..., 1], th[..., 2], th[..., 3])
    if zero_if_degenerate:
        # the fork copy predicts no change without usable data (:112-117): fewer than two distinct training weights
        front = handle["front"]
        with torch.cuda.stream(fit_stream()):
            x = front["pack"][0].cpu().numpy()
        for f in range(n * M):
            k = int(front["klen"][f // M])
            if k == 0 or len(np.unique(x[f, :k])) < 2:
                delta[f // M, :, f % M] = 0.0
    pred = view.objs[handle["node_ids"]][:, None, :] + delta
    keep = np.nonzero(np.repeat(counts > 0, M))[0]
    record = FitRecord(lambda: _ragged_inputs(handle, keep.tolist()), theta=theta[keep], status=status[keep],
                       nfev=nfev[keep], cost=cost[keep])
    return pred, record


def _my_members(n):
    """Indices of the population members whose fits this rank computes (all of them in a single process)."""
    from . import dist as pdist
    rank, W = pdist.world()
    return np.arange(n) if (W == 1 or n < W) else np.arange(rank, n, W)


def prefetch(opt_graph, samples, obj_num, cap_threshold):
    """Launch this rank's share of the fits of `samples` NOW and return the handle; `predict_candidates(..., pending=...)`
    picks it up if population and opt-graph are still the ones it was launched for. Lets the caller put the fit chain
    (milliseconds of dependent FP64 latency) under host work that does not need it, e.g. the archive update."""
    ids = np.array([s.optgraph_id for s in samples], dtype=np.int64)
    pending = launch_fits(opt_graph, ids[_my_members(len(ids))], obj_num, cap_threshold)
    pending["all_ids"] = ids
    return pending


def _usable(pending, opt_graph, ids, cap_threshold):
    return (pending is not None and pending["n_nodes"] == len(opt_graph.objs) and pending["cap_threshold"] == bool(cap_threshold)
            and np.array_equal(pending["all_ids"], ids))


def predict_candidates(opt_graph, samples, make_tests, obj_num, cap_threshold, max_tests, zero_if_degenerate=False,
                       tests_in_lockstep=False, pending=None, while_fitting=None):
    """Test weights and predicted objectives of every population member.
    `make_tests(view, node_ids) -> (tests [n, max_tests, M], counts [n])` enumerates the test weights of the given
    members (rows past counts[i] are padding, any finite positive numbers).
    -> (tests [n_pop, max_tests, M], counts [n_pop], pred [n_pop, max_tests, M], fit record).

    Single process: all fits in one K4 launch, issued before the host enumerates the test weights so that it runs under
    that work. Under torch.distributed (sharded runs) the members are split over the ranks
    (member i on rank i % W): each rank builds the fit inputs, runs K4 and evaluates the models for ITS members only, then
    one all-gather of a padded float64 table (count, test weights, predictions per member) gives every rank the identical
    candidate set -- the per-rank cost of the selection front-end stays flat as tasks and GPUs grow together.
    `tests_in_lockstep`: the test weights consume numpy's global RNG (3 objectives), so every rank enumerates them for
    every member to keep the streams identical; only fits and predictions are split.
    `pending`: a handle from `prefetch` (used if it was launched for this population and opt-graph);
    `while_fitting`: host work of the caller to run once the test weights are enumerated and before the fits are awaited."""
    from . import dist as pdist
    rank, W = pdist.world()
    n = len(samples)
    ids = np.array([s.optgraph_id for s in samples], dtype=np.int64)
    M = obj_num
    mine = _my_members(n)
    if not _usable(pending, opt_graph, ids, cap_threshold):
        pending = launch_fits(opt_graph, ids[mine], obj_num, cap_threshold)
    if W == 1 or n < W:
        with _nvtx("selection.test_weights"):
            tests, counts = make_tests(pending["view"], ids)
        if while_fitting is not None:
            while_fitting()
        pred, fits = finish_predictions(pending, tests, counts, zero_if_degenerate=zero_if_degenerate)
        return tests, counts, pred, fits
    with _nvtx("selection.test_weights"):
        if tests_in_lockstep:
            tests, counts = make_tests(pending["view"], ids)
            my_tests, my_counts = tests[mine], counts[mine]
        else:
            my_tests, my_counts = make_tests(pending["view"], ids[mine])
    if while_fitting is not None:
        while_fitting()
    my_pred, fits = finish_predictions(pending, my_tests, my_counts, zero_if_degenerate=zero_if_degenerate)
    T = max_tests
    rows = np.zeros((len(mine), 1 + 2 * T * M))
    rows[:, 0] = my_counts
    valid = (np.arange(T)[None, :] < np.asarray(my_counts)[:, None])[:, :, None]
    rows[:, 1:1 + T * M] = np.where(valid, my_tests, 0.0).reshape(len(mine), -1)
    rows[:, 1 + T * M:] = np.where(valid, my_pred, 0.0).reshape(len(mine), -1)
    table = pdist.all_gather_rows(rows, n)
    counts = table[:, 0].astype(np.int64)
    tests = table[:, 1:1 + T * M].reshape(n, T, M).copy()
    pred = table[:, 1 + T * M:].reshape(n, T, M).copy()
    tests[~(np.arange(T)[None, :] < counts[:, None])] = 1.0          # padding stays finite and positive
    return tests, counts, pred, fits


class Candidates:
    """The (sample, weight, prediction) triples of a selection call, kept as arrays; indexable like the reference's
    list of dicts (`candidates[i]['prediction']`)."""

    def __init__(self, samples, tests, counts, pred):
        valid = np.arange(tests.shape[1])[None, :] < np.asarray(counts)[:, None]
        self.member, slot = np.nonzero(valid)                       # member order, then test order: the reference's order
        self.weight = tests[self.member, slot]
        self.prediction = pred[self.member, slot]
        self._samples = samples

    def __len__(self):
        return len(self.member)

    def __getitem__(self, i):
        return {'sample': self._samples[int(self.member[i])], 'weight': self.weight[i], 'prediction': self.prediction[i]}

    def __iter__(self):
        return (self[i] for i in range(len(self)))
