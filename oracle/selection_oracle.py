"""ORACLE (test infrastructure, NOT product code) -- float64 restatement of the arithmetic of the
reference's prediction-guided selection that the CUDA kernels K4/K5 replace: Pareto filtering,
exact 2-D / 3-D hypervolume, sparsity, the greedy candidate pick, and the hyperbolic model fit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this. Pinned by tests/test_oracle_selection.py against tests/golden/selection_*.npz, which hold
outputs of the UNMODIFIED reference (tests/golden/make_golden_selection.py).

Third-party arithmetic: the model fit is `scipy.optimize.least_squares` (pinned 1.4.1 in the
reference's environment.yml:103; 1.18.1 in this image): `fit_scipy` calls it exactly as
morl/population_2d.py:106 does; `trf_fit` is a step-for-step numpy restatement of its
`trf_bounds` path (scipy/optimize/_lsq/{least_squares,trf,common}.py) that the CUDA port follows.

Reference anchors (paths relative to /root/reference/morl):
  get_ep_indices / check_dominated   utils.py:24-39
  update_ep                          utils.py:42-65
  2-D hypervolume / sparsity         population_2d.py:185-202
  M-D sparsity                       utils.py:87-100
  3-D hypervolume                    hypervolume.py:41-153 (round(hv, 4) at :74)
  greedy selection                   population_2d.py:264-304, population_3d.py:294-333
  hyperbolic model                   population_2d.py:56-108
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# Pareto filtering
# ------------------------------------------------------------------------------------------------
def get_ep_indices(objs):
    """Indices, in np.argsort(obj0) order, of points with all objs >= 0 that no other point
    weakly dominates with one strict inequality (utils.py:24-39)."""
    objs = np.asarray(objs, dtype=np.float64)
    if len(objs) == 0:
        return []
    keep = []
    for idx in np.argsort(objs.T[0]):
        p = objs[idx]
        dominated = np.logical_and((objs >= p).all(axis=1), (objs > p).any(axis=1)).any()
        if (p >= 0).all() and not dominated:
            keep.append(int(idx))
    return keep


def update_ep(ep, new):
    """utils.py:42-65: incremental front update with the reference's 1e-5 tolerances; returns a list
    of points ordered by objective 0."""
    new = np.asarray(new, dtype=np.float64)
    ep = [np.asarray(p, dtype=np.float64) for p in ep]
    if (new < 0).any():
        return ep
    out, on_ep = [], True
    for p in ep:
        if (p >= new - 1e-5).all() and (p > new + 1e-5).any():
            on_ep = False
        if not (new >= p).all():
            out.append(p)
    if on_ep:
        pos = len(out)
        for i, p in enumerate(out):
            if new[0] < p[0]:
                pos = i
                break
        out.insert(pos, new)
    return out


# ------------------------------------------------------------------------------------------------
# metrics
# ------------------------------------------------------------------------------------------------
def hv2d(objs):
    """population_2d.py:185-192: sum over the sorted front of (x_i - x_prev) * y_i, reference point 0."""
    objs = np.asarray(objs, dtype=np.float64)
    x, hv = 0.0, 0.0
    for i in get_ep_indices(objs):
        hv += (max(0.0, objs[i, 0]) - x) * (max(0.0, objs[i, 1]) - 0.0)
        x = max(0.0, objs[i, 0])
    return hv


def sparsity2d(objs):
    """population_2d.py:194-202: mean squared distance between neighbours on the sorted front."""
    objs = np.asarray(objs, dtype=np.float64)
    idx = get_ep_indices(objs)
    if len(idx) < 2:
        return 0.0
    s = 0.0
    for a, b in zip(idx[1:], idx[:-1]):
        d = objs[a] - objs[b]
        s += np.sum(np.square(d))
    return s / (len(idx) - 1)


def sparsity_md(front):
    """utils.py:87-100: per dimension, sum of squared gaps between sorted values; / (n-1)."""
    front = np.asarray(front, dtype=np.float64)
    if len(front) < 2:
        return 0.0
    s = 0.0
    for dim in range(front.shape[1]):
        v = np.sort(front[:, dim])
        for i in range(1, len(v)):
            s += np.square(v[i] - v[i - 1])
    return s / (len(front) - 1)


def _area2_presorted(q):
    """Inner 2-D sweep of hvRecursive (hypervolume.py:92-105) over negated points (all coords <= 0)
    already in the y-list order: h = running min of n_x."""
    h = q[0][0]
    acc = 0.0
    prev = q[0]
    for p in q[1:]:
        acc += h * (prev[1] - p[1])
        if p[0] < h:
            h = p[0]
        prev = p
    acc += h * prev[1]
    return acc


def hv3d_raw(front):
    """Un-rounded 3-D hypervolume in the reference's summation order (hypervolume.py:106-153 with
    dimIndex = 2). Points are negated and those outside the reference orthant dropped (:55-61).
    preProcess (:156-164) sorts the node list by x, then (stably) by y, then (stably) by z and
    records each order, so the y-list is the stable y-sort of the x-sorted list and the z-list the
    stable z-sort of that. The sweep inserts points in z-list order; after inserting point k the
    2-D area of the inserted points (walked in y-list order) multiplies the gap to the next z."""
    pts = [[-float(v) for v in p] for p in front]
    pts = [p for p in pts if all(v <= 0.0 for v in p)]
    if not pts:
        return 0.0
    ylist = sorted(sorted(range(len(pts)), key=lambda i: pts[i][0]), key=lambda i: pts[i][1])
    zlist = sorted(ylist, key=lambda i: pts[i][2])
    hv, area_prev = 0.0, 0.0
    inserted = set()
    for k, i in enumerate(zlist):
        if k > 0:
            hv += area_prev * (pts[i][2] - pts[zlist[k - 1]][2])
        inserted.add(i)
        area_prev = _area2_presorted([pts[j] for j in ylist if j in inserted])
    hv -= area_prev * pts[zlist[-1]][2]
    return hv


def hv3d(front):
    """utils.compute_hypervolume for 3 objectives: InnerHyperVolume(zeros).compute -> round(hv, 4)."""
    return round(hv3d_raw(front), 4)


# ------------------------------------------------------------------------------------------------
# greedy candidate pick
# ------------------------------------------------------------------------------------------------
def greedy_select_2d(ep_objs, preds, alpha, num_tasks):
    """population_2d.py:264-304. Returns (best ids, per-round hv arrays, per-round sparsity arrays)."""
    vep = [np.asarray(p, dtype=np.float64) for p in ep_objs]
    preds = np.asarray(preds, dtype=np.float64)
    mask = np.ones(len(preds), dtype=bool)
    best_ids, hvs, sps = [], [], []
    for _ in range(num_tasks):
        hv = np.zeros(len(preds)); sp = np.zeros(len(preds))
        for i in range(len(preds)):
            if mask[i]:
                batch = np.array(vep + [preds[i]])
                hv[i] = hv2d(batch); sp[i] = sparsity2d(batch)
        hvs.append(hv); sps.append(sp)
        best, best_v = -1, -np.inf
        for i in range(len(preds)):
            if mask[i] and hv[i] - alpha * sp[i] > best_v:
                best, best_v = i, hv[i] - alpha * sp[i]
        if best == -1:
            break
        best_ids.append(best)
        mask[best] = False
        batch = np.array(vep + [preds[best]])
        vep = [batch[i] for i in get_ep_indices(batch)]
    return best_ids, hvs, sps


def greedy_select_3d(ep_objs, preds, alpha, num_tasks):
    """population_3d.py:294-333 with the serial scorer (:206-214)."""
    vep = [np.asarray(p, dtype=np.float64) for p in ep_objs]
    preds = np.asarray(preds, dtype=np.float64)
    mask = np.ones(len(preds), dtype=bool)
    best_ids, hvs, sps = [], [], []
    for _ in range(num_tasks):
        hv = np.zeros(len(preds)); sp = np.zeros(len(preds))
        for i in range(len(preds)):
            if mask[i]:
                new_ep = update_ep(vep, preds[i])
                hv[i] = hv3d(new_ep); sp[i] = sparsity_md(new_ep)
        hvs.append(hv); sps.append(sp)
        best, best_v = -1, -np.inf
        for i in range(len(preds)):
            if mask[i] and hv[i] - alpha * sp[i] > best_v:
                best, best_v = i, hv[i] - alpha * sp[i]
        if best == -1:
            break
        best_ids.append(best)
        mask[best] = False
        vep = update_ep(vep, preds[best])
    return best_ids, hvs, sps


# ------------------------------------------------------------------------------------------------
# hyperbolic prediction model (population_2d.py:56-108)
# ------------------------------------------------------------------------------------------------
def model(x, A, a, b, c):
    e = np.exp(a * (x - b))
    return A * (e - 1) / (e + 1) + c


def residual(p, x, y, w):
    e = np.exp(p[1] * (x - p[2]))
    return (p[0] * (e - 1.0) / (e + 1) + p[3] - y) * w


def jacobian(p, x, y, w):
    A, a, b = p[0], p[1], p[2]
    e = np.exp(a * (x - b))
    J = np.zeros((4, len(x)))
    J[0] = ((e - 1) / (e + 1)) * w
    J[1] = (A * (x - b) * (2.0 * e) / ((e + 1) ** 2)) * w
    J[2] = (A * (-a) * (2.0 * e) / ((e + 1) ** 2)) * w
    J[3] = w
    return J.T


LB = np.array([0.0, 0.1, -5.0, -500.0])


def upper_bounds(y):
    return np.array([np.clip(np.max(y) - np.min(y), 1.0, 500.0), 20.0, 5.0, 500.0])


def fit_scipy(x, y, w, ub=None):
    """The reference's call (population_2d.py:106): soft_l1 loss, f_scale 20, bounded TRF from ones(4)."""
    from scipy.optimize import least_squares
    ub = upper_bounds(y) if ub is None else ub
    return least_squares(lambda p, xx, yy: residual(p, xx, yy, w), np.ones(4), loss="soft_l1", f_scale=20.0,
                         args=(x, y), jac=lambda p, xx, yy: jacobian(p, xx, yy, w), bounds=(LB, ub))
