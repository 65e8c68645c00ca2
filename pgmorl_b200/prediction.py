"""Hyperbolic prediction model of PG-MORL (mirror of predict_hyperbolic / collect_nearest_data,
morl/population_2d.py:12-118 and morl/population_3d.py:13-113), batched over the whole population.

Host side (numpy, same expressions as the reference so the fit inputs are bit-identical): neighbourhood
search in the opt-graph with the widening (threshold, sigma) schedule and the Gaussian point weights.
Device side: ALL fits of a selection call (n_pop x M bounded robust least-squares problems) in one
launch of K4 (csrc/k4_fit.cu). Predictions are then objs + f(test weight)."""
import numpy as np

from . import kernels as K
from ._nvtx import rng as _nvtx
from .utils import norm2


class GraphView:
    """Flat arrays over an OptGraph for repeated neighbourhood queries."""

    def __init__(self, opt_graph):
        self.objs = np.array([np.asarray(o, dtype=np.float64) for o in opt_graph.objs])
        parents, children = [], []
        for i, succ in enumerate(opt_graph.succ):           # data order of collect_nearest_data: by node, then by successor
            for s in succ:
                parents.append(i); children.append(s)
        self.parent = np.array(parents, dtype=np.int64)
        self.child = np.array(children, dtype=np.int64)
        # successor weights normalised to sum 1 (population_2d.py:19) and their objective gains
        self.edge_w = np.array([np.asarray(opt_graph.weights[s], dtype=np.float64) / np.sum(np.asarray(opt_graph.weights[s], dtype=np.float64))
                                for s in children]).reshape(len(children), -1)
        self.edge_dy = np.array([np.asarray(opt_graph.delta_objs[s], dtype=np.float64) for s in children]).reshape(len(children), -1)

    def nearest_edges(self, k, threshold):
        """Edges (i -> s) whose source node i lies within `threshold` (relative, per objective) of node k."""
        ok = self.objs[k]
        near = np.all(np.abs(ok - self.objs) < np.abs(ok) * threshold, axis=1)
        return np.nonzero(near[self.parent])[0] if len(self.parent) else np.zeros(0, dtype=np.int64)


def _enough_distinct(weights):
    """More than 3 pairwise-distinct weights (L2 distance >= 1e-5), first-occurrence scan (population_2d.py:39-49)."""
    cnt = 0
    for i in range(len(weights)):
        distinct = True
        for j in range(i):
            if norm2(weights[i] - weights[j]) < 1e-5:
                distinct = False
                break
        if distinct:
            cnt += 1
            if cnt > 3:
                return True
    return False


def fit_inputs(view, k, obj_num, cap_threshold):
    """Training data of the model of node k: per objective (x, y, w, ub). `cap_threshold` reproduces the
    3-objective variant's stop at threshold >= 1 (population_3d.py:46); the 2-objective one widens until
    more than 3 distinct weights are found (population_2d.py:50)."""
    threshold, sigma = 0.1, 0.03
    ok = view.objs[k]
    aok = np.abs(ok)
    rel = np.abs(ok - view.objs)                              # |objs_k - objs_i| for every node i, reused by each widening
    has_edges = len(view.parent) > 0
    near_block = None
    step = 0
    while True:
        if step == 0:
            lt = rel < aok * threshold
            near = lt[:, 0]
            for m in range(1, lt.shape[1]):
                near = near & lt[:, m]
        else:
            if (step - 1) % 4 == 0:                           # neighbourhoods of the next four thresholds in one comparison
                thr = threshold * np.array([1.0, 2.0, 4.0, 8.0])     # exact doublings: the values `threshold *= 2` goes through
                lt = rel[None, :, :] < aok[None, None, :] * thr[:, None, None]
                near_block = lt[:, :, 0]
                for m in range(1, lt.shape[2]):
                    near_block = near_block & lt[:, :, m]
            near = near_block[(step - 1) % 4]
        e = np.nonzero(near[view.parent])[0] if has_edges else np.zeros(0, dtype=np.int64)
        wd = view.edge_w[e]
        if _enough_distinct(wd) or (cap_threshold and threshold >= 1.0):
            break
        if not np.isfinite(threshold):          # the reference would widen forever here (fewer than 4 distinct weights
            break                               # in the whole graph); fits are launched for every sample, so stop instead
        threshold *= 2.0
        sigma *= 2.0
        step += 1
    q = rel / aok                                             # same element-wise operations as population_2d.py:92-93
    coef = np.empty(len(e))
    node_coef = {}
    parents = view.parent[e].tolist()
    for r, i in enumerate(parents):                           # one Gaussian weight per source node, shared by its edges
        c = node_coef.get(i)
        if c is None:
            dist = norm2(q[i])
            c = node_coef[i] = np.exp(-((dist / sigma) ** 2) / 2.0)
        coef[r] = c
    out = []
    dy = view.edge_dy[e]
    for dim in range(obj_num):
        x = wd[:, dim].copy()
        y = dy[:, dim].copy()
        span = y.max() - y.min()                              # np.clip(max - min, 1, 500) of population_2d.py:100
        ub = np.array([min(max(span, 1.0), 500.0), 20.0, 5.0, 500.0])
        out.append((x, y, coef.copy(), ub))
    return out


def model(x, A, a, b, c):
    return A * (np.exp(a * (x - b)) - 1) / (np.exp(a * (x - b)) + 1) + c


def launch_fits(opt_graph, node_ids, obj_num, cap_threshold):
    """Build the training data of every node's model and launch all n x M fits (K4) without waiting for them.
    Returns a handle for `finish_predictions`; the caller may do host work that does not need the fits meanwhile."""
    with _nvtx("selection.fit_inputs"):
        view = GraphView(opt_graph)
        xs, ys, ws, ubs = [], [], [], []
        for k in node_ids:
            for x, y, w, ub in fit_inputs(view, k, obj_num, cap_threshold):
                xs.append(x); ys.append(y); ws.append(w); ubs.append(ub)
    with _nvtx("selection.k4_fits"):
        fits = K.fit_hyperbolic_launch(xs, ys, ws, ubs)                    # all fits in one launch
    return dict(view=view, node_ids=list(node_ids), obj_num=obj_num, x=xs, y=ys, w=ws, ub=ubs, fits=fits)


def finish_predictions(handle, test_weights_per_node, zero_if_degenerate=False):
    """Second half of `predict_population`. test_weights_per_node[i] is an array [n_i, M] for node_ids[i] (any
    positive scaling; normalised to sum 1 here, as population_2d.py:28-32) or None / empty: that node then
    contributes neither predictions nor fit records (the reference never fits a sample without test weights).
    `zero_if_degenerate`: the fork copy's fallback (WorkingMorl/morl/population_2d.py:112-117)."""
    theta, status, nfev, cost = K.fit_hyperbolic_collect(handle["fits"])
    view, M = handle["view"], handle["obj_num"]
    preds, keep = [], []
    for i, k in enumerate(handle["node_ids"]):
        if test_weights_per_node[i] is None or len(test_weights_per_node[i]) == 0:
            continue
        keep.extend(range(i * M, (i + 1) * M))
        tw = np.array(test_weights_per_node[i], dtype=np.float64)
        tw = tw / tw.sum(axis=1, keepdims=True)                   # row /= np.sum(row), population_2d.py:28-32 (same bits)
        cols = []
        for dim in range(M):
            x = handle["x"][i * M + dim]
            if zero_if_degenerate and (len(x) == 0 or len(np.unique(x)) < 2):
                cols.append(np.zeros(len(tw)))       # the fork copy predicts no change without usable data (:112-117)
            else:
                cols.append(model(tw.T[dim], *theta[i * M + dim]))
        delta = np.transpose(np.array(cols))
        preds.append(view.objs[k][None, :] + delta)
    pick = lambda seq: [seq[j] for j in keep]
    return preds, dict(x=pick(handle["x"]), y=pick(handle["y"]), w=pick(handle["w"]), ub=pick(handle["ub"]),
                       theta=theta[keep], status=status[keep], nfev=nfev[keep], cost=cost[keep])


def predict_candidates(opt_graph, samples, make_tests, obj_num, cap_threshold, max_tests, zero_if_degenerate=False,
                       tests_in_lockstep=False):
    """Test weights and predicted objectives of every population member: -> (all_tests: one [n_i, M] list per member,
    preds: one [n_i, M] array per member WITH test weights, fit record).

    Single process: all fits in one K4 launch, issued before the host enumerates the test weights so that it runs under
    that work. Under torch.distributed (sharded runs) the members are split over the ranks
    (member i on rank i % W): each rank builds the fit inputs, runs K4 and evaluates the models for ITS members only, then
    one all-gather of a padded float64 table (count, test weights, predictions per member) gives every rank the identical
    candidate set -- the per-rank cost of the selection front-end stays flat as tasks and GPUs grow together.
    `tests_in_lockstep`: the test weights consume numpy's global RNG (3 objectives), so every rank enumerates them for
    every member to keep the streams identical; only fits and predictions are split."""
    from . import dist as pdist
    rank, W = pdist.world()
    n = len(samples)
    if W == 1 or n < W:
        pending = launch_fits(opt_graph, [s.optgraph_id for s in samples], obj_num, cap_threshold)
        all_tests = [make_tests(s) for s in samples]
        preds, fits = finish_predictions(pending, all_tests, zero_if_degenerate=zero_if_degenerate)
        return all_tests, preds, fits
    mine = list(range(rank, n, W))
    pending = launch_fits(opt_graph, [samples[i].optgraph_id for i in mine], obj_num, cap_threshold)
    if tests_in_lockstep:
        everyone = [make_tests(s) for s in samples]
        my_tests = [everyone[i] for i in mine]
    else:
        my_tests = [make_tests(samples[i]) for i in mine]
    my_preds, fits = finish_predictions(pending, my_tests, zero_if_degenerate=zero_if_degenerate)
    M = obj_num
    rows = np.zeros((len(mine), 1 + 2 * max_tests * M))
    it = iter(my_preds)
    for j, tw in enumerate(my_tests):
        k = len(tw)
        assert k <= max_tests
        rows[j, 0] = k
        if k:
            rows[j, 1:1 + k * M] = np.asarray(tw, dtype=np.float64).reshape(-1)
            rows[j, 1 + max_tests * M:1 + max_tests * M + k * M] = np.asarray(next(it), dtype=np.float64).reshape(-1)
    table = pdist.all_gather_rows(rows, n)
    all_tests, preds = [], []
    for i in range(n):
        k = int(table[i, 0])
        tw = table[i, 1:1 + k * M].reshape(k, M)
        all_tests.append([tw[j].copy() for j in range(k)])
        if k:
            preds.append(table[i, 1 + max_tests * M:1 + max_tests * M + k * M].reshape(k, M).copy())
    return all_tests, preds, fits


def predict_population(opt_graph, node_ids, test_weights_per_node, obj_num, cap_threshold):
    """For every node: predicted objectives objs + delta(test weight) for each of its test weights.
    Returns (list of [n_i, M] prediction arrays, fit record dict)."""
    return finish_predictions(launch_fits(opt_graph, node_ids, obj_num, cap_threshold), test_weights_per_node)
