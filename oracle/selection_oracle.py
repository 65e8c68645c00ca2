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


def hv2d_inner(front):
    """utils.compute_hypervolume (InnerHyperVolume, round(hv, 4)) of a 2-objective front, as the fork copy's scorer
    calls it (WorkingMorl/morl/population_2d.py:213-214). The dimension sweep on 2-D points is the per-slice area
    routine of the 3-D case; lifting the points to z = 1 gives the same bits (checked against the reference's
    InnerHyperVolume in both dimensions when the fork golden was made)."""
    return hv3d([[float(p[0]), float(p[1]), 1.0] for p in front])


def greedy_select_2d_fork(ep_objs, preds, alpha, num_tasks):
    """The fork copy's 2-objective greedy loop (WorkingMorl/morl/population_2d.py:207-226, 266-306): candidates scored
    with update_ep + InnerHyperVolume + M-D sparsity; the virtual front is still rebuilt with get_ep_indices."""
    vep = [np.asarray(p, dtype=np.float64) for p in ep_objs]
    preds = np.asarray(preds, dtype=np.float64)
    mask = np.ones(len(preds), dtype=bool)
    best_ids, hvs, sps = [], [], []
    for _ in range(num_tasks):
        hv = np.zeros(len(preds)); sp = np.zeros(len(preds))
        for i in range(len(preds)):
            if mask[i]:
                new_ep = update_ep(vep, preds[i])
                hv[i] = hv2d_inner(new_ep); sp[i] = sparsity_md(new_ep)
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


# ------------------------------------------------------------------------------------------------
# numpy restatement of scipy's bounded Trust Region Reflective path for this 4-parameter fit
# (scipy/optimize/_lsq/least_squares.py:900-1030, trf.py:129-412, common.py). With numpy's own
# LAPACK SVD it reproduces scipy bit for bit (tests/test_oracle_selection.py); the CUDA port
# (csrc/k4_fit.cu) follows it line by line with a one-sided Jacobi SVD.
# ------------------------------------------------------------------------------------------------
EPS = np.finfo(float).eps
F_SCALE = 20.0


def _soft_l1(f, cost_only=False):
    z = (f / F_SCALE) ** 2
    t = 1 + z
    rho0 = 2 * (t ** 0.5 - 1)
    if cost_only:
        return 0.5 * F_SCALE ** 2 * np.sum(rho0)
    return rho0 * F_SCALE ** 2, t ** -0.5, (-0.5 * t ** -1.5) / F_SCALE ** 2


def _scale_robust(J, f, rho1, rho2):
    js = rho1 + 2 * rho2 * f ** 2
    js[js < EPS] = EPS
    js = js ** 0.5
    return J * js[:, None], f * (rho1 / js)


def _cl_scaling(x, g, lb, ub):
    v = np.ones_like(x); dv = np.zeros_like(x)
    m = g < 0
    v[m] = ub[m] - x[m]; dv[m] = -1
    m = g > 0
    v[m] = x[m] - lb[m]; dv[m] = 1
    return v, dv


def _strictly_feasible(x, lb, ub, rstep):
    xn = x.copy()
    lower_dist, upper_dist = x - lb, ub - x
    if rstep == 0:
        lower, upper = x <= lb, x >= ub
        xn[lower] = np.nextafter(lb[lower], ub[lower]); xn[upper] = np.nextafter(ub[upper], lb[upper])
    else:
        lt, ut = rstep * np.maximum(1, np.abs(lb)), rstep * np.maximum(1, np.abs(ub))
        lower = lower_dist <= np.minimum(upper_dist, lt)
        upper = upper_dist <= np.minimum(lower_dist, ut)
        xn[lower] = lb[lower] + rstep * np.maximum(1, np.abs(lb[lower]))
        xn[upper] = ub[upper] - rstep * np.maximum(1, np.abs(ub[upper]))
    tight = (xn < lb) | (xn > ub)
    xn[tight] = 0.5 * (lb[tight] + ub[tight])
    return xn


def _step_to_bound(x, s, lb, ub):
    steps = np.full_like(x, np.inf)
    nz = s != 0
    with np.errstate(over="ignore"):
        steps[nz] = np.maximum((lb - x)[nz] / s[nz], (ub - x)[nz] / s[nz])
    mn = np.min(steps)
    return mn, np.equal(steps, mn) * np.sign(s).astype(int)


def _quad1d(Jh, g, s, diag, s0=None):
    v = Jh.dot(s)
    a = (np.dot(v, v) + np.dot(s * diag, s)) * 0.5
    b = np.dot(g, s)
    if s0 is None:
        return a, b
    u = Jh.dot(s0)
    b += np.dot(u, v)
    c = 0.5 * np.dot(u, u) + np.dot(g, s0)
    b += np.dot(s0 * diag, s)
    c += 0.5 * np.dot(s0 * diag, s0)
    return a, b, c


def _min_quad1d(a, b, lo, hi, c=0.0):
    t = [lo, hi]
    if a != 0:
        ex = -0.5 * b / a
        if lo < ex < hi:
            t.append(ex)
    t = np.asarray(t)
    yv = t * (a * t + b) + c
    i = np.argmin(yv)
    return t[i], yv[i]


def _eval_quad(Jh, g, s, diag):
    Js = Jh.dot(s)
    return 0.5 * (np.dot(Js, Js) + np.dot(s * diag, s)) + np.dot(s, g)


def _solve_tr(n, m, uf, s, V, Delta, alpha0):
    def phi_d(alpha):
        denom = s ** 2 + alpha
        pn = np.linalg.norm(suf / denom)
        return pn - Delta, -np.sum(suf ** 2 / denom ** 3) / pn

    suf = s * uf
    full_rank = m >= n and s[-1] > EPS * m * s[0]
    if full_rank:
        p = -V.dot(uf / s)
        if np.linalg.norm(p) <= Delta:
            return p, 0.0
    alpha_upper = np.linalg.norm(suf) / Delta
    if full_rank:
        phi, phip = phi_d(0.0)
        alpha_lower = -phi / phip
    else:
        alpha_lower = 0.0
    if not full_rank and alpha0 == 0:
        alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
    else:
        alpha = alpha0
    for _ in range(10):
        if alpha < alpha_lower or alpha > alpha_upper:
            alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
        phi, phip = phi_d(alpha)
        if phi < 0:
            alpha_upper = alpha
        ratio = phi / phip
        alpha_lower = max(alpha_lower, alpha - ratio)
        alpha -= (phi + Delta) * ratio / Delta
        if np.abs(phi) < 0.01 * Delta:
            break
    p = -V.dot(suf / (s ** 2 + alpha))
    p *= Delta / np.linalg.norm(p)
    return p, alpha


def _select_step(x, Jh, diag, gh, p, ph, d, Delta, lb, ub, theta):
    if np.all((x + p >= lb) & (x + p <= ub)):
        return p, ph, -_eval_quad(Jh, gh, ph, diag)
    p_stride, hits = _step_to_bound(x, p, lb, ub)
    rh = np.copy(ph)
    rh[hits.astype(bool)] *= -1
    r = d * rh
    p = p * p_stride; ph = ph * p_stride
    x_on = x + p
    a_ = np.dot(rh, rh); b_ = np.dot(ph, rh); c_ = np.dot(ph, ph) - Delta ** 2
    dd = np.sqrt(b_ * b_ - a_ * c_)
    q = -(b_ + np.copysign(dd, b_))
    t1, t2 = q / a_, c_ / q
    to_tr = max(t1, t2)
    to_bound, _ = _step_to_bound(x_on, r, lb, ub)
    r_stride = min(to_bound, to_tr)
    if r_stride > 0:
        lo = (1 - theta) * p_stride / r_stride
        hi = theta * to_bound if r_stride == to_bound else to_tr
    else:
        lo, hi = 0, -1
    if lo <= hi:
        a, b, c = _quad1d(Jh, gh, rh, diag, s0=ph)
        r_stride, r_value = _min_quad1d(a, b, lo, hi, c=c)
        rh = rh * r_stride + ph
        r = rh * d
    else:
        r_value = np.inf
    p = p * theta; ph = ph * theta
    p_value = _eval_quad(Jh, gh, ph, diag)
    agh = -gh
    ag = d * agh
    to_tr = Delta / np.linalg.norm(agh)
    to_bound, _ = _step_to_bound(x, ag, lb, ub)
    ag_stride = theta * to_bound if to_bound < to_tr else to_tr
    a, b = _quad1d(Jh, gh, agh, diag)
    ag_stride, ag_value = _min_quad1d(a, b, 0, ag_stride)
    agh = agh * ag_stride; ag = ag * ag_stride
    if p_value < r_value and p_value < ag_value:
        return p, ph, -p_value
    if r_value < p_value and r_value < ag_value:
        return r, rh, -r_value
    return ag, agh, -ag_value


def trf_fit(x, y, w, ub, svd=None, max_nfev=400, ftol=1e-8, xtol=1e-8, gtol=1e-8):
    """Restated least_squares(..., method='trf', loss='soft_l1', f_scale=20, bounds=(LB, ub)) from ones(4).
    Returns (theta, status, nfev, cost)."""
    svd = svd or (lambda A: np.linalg.svd(A, full_matrices=False))
    lb = LB
    fun = lambda p: residual(p, x, y, w)
    jac = lambda p: jacobian(p, x, y, w)
    xk = _strictly_feasible(np.ones(4), lb, ub, 1e-10)
    f = fun(xk); nfev = 1
    J = jac(xk)
    m, n = J.shape
    rho0, rho1, rho2 = _soft_l1(f)
    cost = 0.5 * np.sum(rho0)
    J, f = _scale_robust(J, f, rho1, rho2)
    g = J.T.dot(f)
    v, dv = _cl_scaling(xk, g, lb, ub)
    Delta = np.linalg.norm(xk / v ** 0.5)
    if Delta == 0:
        Delta = 1.0
    alpha = 0.0
    status = None
    while True:
        v, dv = _cl_scaling(xk, g, lb, ub)
        g_norm = np.linalg.norm(g * v, ord=np.inf)
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        d = v ** 0.5
        diag = g * dv
        gh = d * g
        f_aug = np.zeros(m + n); f_aug[:m] = f
        J_aug = np.empty((m + n, n))
        J_aug[:m] = J * d
        Jh = J_aug[:m].copy()
        J_aug[m:] = np.diag(diag ** 0.5)
        U, s, Vt = svd(J_aug)
        V = Vt.T
        uf = U.T.dot(f_aug)
        theta = max(0.995, 1 - g_norm)
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            ph, alpha = _solve_tr(n, m, uf, s, V, Delta, alpha)
            p = d * ph
            step, step_h, predicted = _select_step(xk, Jh, diag, gh, p, ph, d, Delta, lb, ub, theta)
            x_new = _strictly_feasible(xk + step, lb, ub, 0)
            f_new = fun(x_new); nfev += 1
            shn = np.linalg.norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * shn
                continue
            cost_new = _soft_l1(f_new, cost_only=True)
            actual = cost - cost_new
            if predicted > 0:
                ratio = actual / predicted
            elif predicted == actual == 0:
                ratio = 1
            else:
                ratio = 0
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * shn
            elif ratio > 0.75 and shn > 0.95 * Delta:
                Delta_new = Delta * 2.0
            step_norm = np.linalg.norm(step)
            ft = actual < ftol * cost and ratio > 0.25
            xt = step_norm < xtol * (xtol + np.linalg.norm(xk))
            status = 4 if (ft and xt) else 2 if ft else 3 if xt else None
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual > 0:
            xk = x_new
            f = f_new
            cost = cost_new
            J = jac(xk)
            rho0, rho1, rho2 = _soft_l1(f)
            J, f = _scale_robust(J, f, rho1, rho2)
            g = J.T.dot(f)
    if status is None:
        status = 0
    return xk, status, nfev, cost


def jacobi_svd(A, sweeps=60):
    """One-sided (Hestenes) Jacobi SVD, the algorithm of the CUDA port: returns (U, s, Vt) with s sorted
    in descending order like LAPACK."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[1]
    V = np.eye(n)
    for _ in range(sweeps):
        rotated = False
        for p in range(n - 1):
            for q in range(p + 1, n):
                al, be, ga = A[:, p] @ A[:, p], A[:, q] @ A[:, q], A[:, p] @ A[:, q]
                if ga == 0.0 or abs(ga) <= 1e-16 * np.sqrt(al * be):
                    continue
                rotated = True
                zeta = (be - al) / (2.0 * ga)
                t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                c = 1.0 / np.sqrt(1.0 + t * t); sn = c * t
                ap, aq = A[:, p].copy(), A[:, q].copy()
                A[:, p], A[:, q] = c * ap - sn * aq, sn * ap + c * aq
                vp, vq = V[:, p].copy(), V[:, q].copy()
                V[:, p], V[:, q] = c * vp - sn * vq, sn * vp + c * vq
        if not rotated:
            break
    s = np.sqrt((A * A).sum(0))
    order = np.argsort(-s, kind="stable")
    s = s[order]; A = A[:, order]; V = V[:, order]
    U = np.where(s > 0, A / np.where(s > 0, s, 1.0), 0.0)
    return U, s, V.T


def qr_jacobi_svd(A, sweeps=60):
    """The SVD of the CUDA port (csrc/k4_fit.cu::qr_jacobi_svd): Householder QR of the tall augmented Jacobian,
    then a one-sided Jacobi SVD of the 4 x 4 triangular factor with round-robin pair order
    (0,1)(2,3) | (0,2)(1,3) | (0,3)(1,2). Returns (U, s, Vt), s sorted in descending order like LAPACK.
    Same role as LAPACK's gesdd inside scipy's trf (trf.py:312-318); not bit-identical to it."""
    A = np.array(A, dtype=np.float64)
    m, n = A.shape
    assert n == 4
    Q = np.eye(m)
    for j in range(n):
        x = A[j:, j]
        nx = np.sqrt(x @ x)
        if nx == 0.0:
            continue
        v = x.copy()
        v[0] -= -np.copysign(nx, x[0])
        beta = 1.0 / (nx * (nx + abs(x[0])))
        A[j:, j:] -= np.outer(v, beta * (v @ A[j:, j:]))
        Q[:, j:] -= np.outer(Q[:, j:] @ v, beta * v)
    B = np.triu(A[:n, :n])
    V = np.eye(n)
    for _ in range(sweeps):
        rotated = False
        for p, q in ((0, 1), (2, 3), (0, 2), (1, 3), (0, 3), (1, 2)):
            al, be, ga = B[:, p] @ B[:, p], B[:, q] @ B[:, q], B[:, p] @ B[:, q]
            if ga == 0.0 or ga * ga <= 1e-32 * (al * be):          # |ga| <= 1e-16 sqrt(al be)
                continue
            rotated = True
            d, h = be - al, 2.0 * ga
            t = np.copysign(1.0, d) * h / (abs(d) + np.sqrt(d * d + h * h))
            c = 1.0 / np.sqrt(1.0 + t * t); sn = c * t
            bp, bq = B[:, p].copy(), B[:, q].copy()
            B[:, p], B[:, q] = c * bp - sn * bq, sn * bp + c * bq
            vp, vq = V[:, p].copy(), V[:, q].copy()
            V[:, p], V[:, q] = c * vp - sn * vq, sn * vp + c * vq
        if not rotated:
            break
    s = np.sqrt((B * B).sum(0))
    order = np.argsort(-s, kind="stable")
    s = s[order]; B = B[:, order]; V = V[:, order]
    UR = np.where(s > 0, B / np.where(s > 0, s, 1.0), 0.0)
    return Q[:, :n] @ UR, s, V.T


# ------------------------------------------------------------------------------------------------
# training data of the prediction models: collect_nearest_data + the widening loop
# (population_2d.py:12-21,37-54,90-104; population_3d.py:13-21,33-49), one population member at a
# time with the reference's scalar numpy expressions. Pinned bit for bit by the recorded fit inputs
# of tests/golden/selection_{2d,3d}.npz (tests/test_host_selection.py); the product computes the same
# data for the whole population at once (csrc/k4_inputs.cu + pgmorl_b200/prediction.py).
# ------------------------------------------------------------------------------------------------
def _norm2(v):
    """np.linalg.norm of a real 1-D vector: sqrt of the BLAS dot product (numpy/linalg/_linalg.py, ord=None)."""
    v = np.asarray(v)
    return float(np.sqrt(v.dot(v)))


class GraphArrays:
    """Flat arrays over an opt-graph (lists `objs`, `weights`, `delta_objs`, `succ` as morl/opt_graph.py keeps them),
    edges in the order collect_nearest_data walks them: by node, then by successor."""

    def __init__(self, opt_graph):
        self.objs = np.array([np.asarray(o, dtype=np.float64) for o in opt_graph.objs])
        parents, children = [], []
        for i, succ in enumerate(opt_graph.succ):
            for s in succ:
                parents.append(i); children.append(s)
        self.parent = np.array(parents, dtype=np.int64)
        self.child = np.array(children, dtype=np.int64)
        # successor weights normalised to sum 1 (population_2d.py:19) and their objective gains
        self.edge_w = np.array([np.asarray(opt_graph.weights[s], dtype=np.float64) / np.sum(np.asarray(opt_graph.weights[s], dtype=np.float64))
                                for s in children]).reshape(len(children), -1)
        self.edge_dy = np.array([np.asarray(opt_graph.delta_objs[s], dtype=np.float64) for s in children]).reshape(len(children), -1)


def _enough_distinct(weights):
    """More than 3 pairwise-distinct weights (L2 distance >= 1e-5), first-occurrence scan (population_2d.py:39-49)."""
    cnt = 0
    for i in range(len(weights)):
        distinct = True
        for j in range(i):
            if _norm2(weights[i] - weights[j]) < 1e-5:
                distinct = False
                break
        if distinct:
            cnt += 1
            if cnt > 3:
                return True
    return False


def fit_inputs(view, k, obj_num, cap_threshold, with_steps=False):
    """Training data of the model of node k: per objective (x, y, w, ub). `cap_threshold` reproduces the
    3-objective variant's stop at threshold >= 1 (population_3d.py:46); the 2-objective one widens until
    more than 3 distinct weights are found (population_2d.py:50). Where the reference would widen forever
    (fewer than 4 distinct weights in the whole graph) the loop stops once the threshold has overflowed."""
    threshold, sigma = 0.1, 0.03
    ok = view.objs[k]
    aok = np.abs(ok)
    rel = np.abs(ok - view.objs)
    steps = 0
    with np.errstate(all="ignore"):
        while True:
            near = np.all(rel < aok * threshold, axis=1)
            e = np.nonzero(near[view.parent])[0] if len(view.parent) else np.zeros(0, dtype=np.int64)
            wd = view.edge_w[e]
            if _enough_distinct(wd) or (cap_threshold and threshold >= 1.0):
                break
            if not np.isfinite(threshold):
                break
            threshold *= 2.0
            sigma *= 2.0
            steps += 1
        q = rel / aok                                             # same element-wise operations as population_2d.py:92-93
        coef = np.empty(len(e))
        for r, i in enumerate(view.parent[e].tolist()):
            dist = _norm2(q[i])
            coef[r] = np.exp(-((dist / sigma) ** 2) / 2.0)
    out = []
    dy = view.edge_dy[e]
    for dim in range(obj_num):
        x = wd[:, dim].copy()
        y = dy[:, dim].copy()
        span = (y.max() - y.min()) if len(y) else 1.0            # np.clip(max - min, 1, 500) of population_2d.py:100
        ub = np.array([min(max(span, 1.0), 500.0), 20.0, 5.0, 500.0])
        out.append((x, y, coef.copy(), ub))
    return (out, steps, e) if with_steps else out
