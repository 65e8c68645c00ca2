"""GPU parity tests of K5 (and K4) against the oracle and the reference goldens, through the C ABI.
Bar: bit-exact for hypervolume, sparsity, Pareto membership and selected indices (float64, same
summation order as the reference)."""
import os

import numpy as np
import pytest

from oracle import selection_oracle as so

import json

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# The recorded fits K4 does not reproduce to rtol 1e-6, pinned by index, next to the fits on which scipy ITSELF moves by more
# than that under a 1e-13 relative perturbation of x0 (tests/golden/make_k4_allowlist.py, run on a B200; SURVEY.md section 7,
# protocol (ii)): 18 / 32 / 5 fits of 474 / 996 / 300, of which 18 / 31 / 4 are among scipy's own 19 / 38 / 6 chaotic fits.
K4_PINNED = json.load(open(os.path.join(GOLDEN, "k4_disagreeing_fits.json")))


def test_ep_filter_and_metrics_known_answers():
    from pgmorl_b200 import kernels as K
    z = np.load(os.path.join(GOLDEN, "selection_kats.npz"))
    for i in range(int(z["n_kat"])):
        pts = z[f"kat{i}_pts"]
        idx = K.ep_filter(pts)
        ref = z[f"kat{i}_ep_idx"]
        # same set; same order wherever objective 0 is not tied (numpy's unstable argsort breaks ties arbitrarily)
        assert sorted(idx.tolist()) == sorted(ref.tolist())
        assert np.array_equal(pts[idx][:, 0], pts[ref][:, 0])
        if pts.shape[1] == 2:
            hv, sp = K.front_metrics(pts)
            assert hv == float(z[f"kat{i}_hv"]) and sp == float(z[f"kat{i}_sp"])
        elif f"kat{i}_hv" in z:
            ep = pts[ref]
            hv, sp = K.front_metrics(ep)
            assert hv == float(z[f"kat{i}_hv"]) and sp == float(z[f"kat{i}_sp"])
            hv_all, _ = K.front_metrics(pts[(pts >= 0).all(1)])      # dominated points inside the set
            assert hv_all == float(z[f"kat{i}_hv_all"])


def test_metrics_random_sets_match_oracle():
    from pgmorl_b200 import kernels as K
    rng = np.random.RandomState(7)
    for n in (1, 2, 3, 17, 64, 300):
        p2 = rng.uniform(0, 50, (n, 2)); p3 = rng.uniform(0, 50, (n, 3))
        if n > 3:
            p2[1] = p2[0]; p3[2, 0] = p3[1, 0]; p3[3, 1:] = p3[1, 1:]
        assert K.front_metrics(p2) == (so.hv2d(p2), so.sparsity2d(p2))
        assert K.front_metrics(p3) == (so.hv3d(p3), so.sparsity_md(p3))
    assert K.front_metrics(np.zeros((0, 2))) == (0.0, 0.0)


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_greedy_selection_bit_exact_vs_reference(name, M):
    """Every round's hv / sparsity array and every chosen candidate equal the reference's, for every
    generation of the recorded history (candidate predictions taken from the reference, so K5 is
    tested independently of the K4 fit)."""
    from pgmorl_b200 import kernels as K
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    alpha = float(z["args_f"][0]); num_tasks = int(z["args"][0])
    for g in range(gens):
        pred = z[f"g{g}_cand_pred"]
        best, hv, sp, front = K.select_greedy(z[f"g{g}_round0_vep"], pred, alpha, num_tasks)
        n_rounds = int(z[f"g{g}_n_rounds"])
        assert n_rounds == num_tasks
        for r in range(n_rounds):
            assert np.array_equal(hv[r], z[f"g{g}_round{r}_hv"]), (g, r)
            assert np.array_equal(sp[r], z[f"g{g}_round{r}_sparsity"]), (g, r)
            assert np.array_equal(pred[best[r]], z[f"g{g}_predicted"][r]), (g, r)
            if r + 1 < n_rounds:
                mask = z[f"g{g}_round{r + 1}_mask"]
                assert not mask[best[r]] and mask.sum() == len(pred) - (r + 1)


def test_greedy_edge_cases():
    from pgmorl_b200 import kernels as K
    # fewer candidates than tasks: stops with -1 ("Too few candidates", population_2d.py:287-289)
    ep = np.array([[1.0, 5.0], [3.0, 2.0]])
    cand = np.array([[2.0, 4.0], [-1.0, 9.0]])
    best, hv, sp, front = K.select_greedy(ep, cand, 1.0, 4)
    ob, ohv, osp = so.greedy_select_2d(ep, cand, 1.0, 4)
    assert best.tolist() == ob + [-1] * (4 - len(ob))
    for r in range(len(ob)):
        assert np.array_equal(hv[r], ohv[r]) and np.array_equal(sp[r], osp[r])
    # empty archive
    best, hv, sp, front = K.select_greedy(np.zeros((0, 2)), cand, 1.0, 1)
    ob, ohv, osp = so.greedy_select_2d([], cand, 1.0, 1)
    assert best.tolist() == ob and np.array_equal(hv[0], ohv[0]) and np.array_equal(sp[0], osp[0])
    # 3 objectives, duplicates and a candidate equal to an archive point
    ep3 = np.array([[1.0, 2.0, 3.0], [2.0, 2.0, 2.0], [3.0, 2.0, 1.0]])
    c3 = np.array([[2.0, 2.0, 2.0], [2.5, 2.5, 0.5], [5.0, -1.0, 5.0], [0.5, 0.5, 4.0]])
    best, hv, sp, front = K.select_greedy(ep3, c3, 1e6, 3)
    ob, ohv, osp = so.greedy_select_3d(ep3, c3, 1e6, 3)
    assert best.tolist() == ob
    for r in range(3):
        assert np.array_equal(hv[r], ohv[r]) and np.array_equal(sp[r], osp[r])


@pytest.mark.parametrize("name", ["selection_2d.npz", "selection_3d.npz", "selection_2d_fork.npz"])
def test_k4_fits_match_scipy(name):
    """K4 follows scipy's TRF step for step; the only substitution is the SVD algorithm, and ~2-4 % of these
    fits are chaotic even against scipy itself (hard part 3). Gate: the set of fits whose theta is
    NOT within rtol 1e-6 of scipy's is exactly the pinned allow-list (tests/golden/k4_disagreeing_fits.json: 3.8 % / 3.2 % /
    1.7 % of the fits, all but 0 / 1 / 1 of them fits on which scipy itself is sensitive to a 1e-13 perturbation of x0);
    every fit finite and inside its bounds, and the disagreeing fits are not worse than scipy's by more than the solver
    tolerance in cost ... reported, not hidden."""
    from pgmorl_b200 import kernels as K
    z = np.load(os.path.join(GOLDEN, name))
    xs, ys, ws, ubs, ref, ref_cost, ref_status = [], [], [], [], [], [], []
    for g in range(int(z["meta"][1])):
        for i in range(int(z[f"g{g}_n_fits"])):
            pre = f"g{g}_fit{i}_"
            xs.append(z[pre + "x"]); ys.append(z[pre + "y"]); ws.append(z[pre + "w"]); ubs.append(z[pre + "ub"])
            ref.append(z[pre + "theta"]); ref_cost.append(float(z[pre + "cost"])); ref_status.append(int(z[pre + "status"]))
    theta, status, nfev, cost = K.fit_hyperbolic(xs, ys, ws, ubs)
    ref = np.array(ref); ref_cost = np.array(ref_cost); ubs = np.array(ubs)
    assert np.isfinite(theta).all() and (theta >= so.LB - 1e-12).all() and (theta <= ubs + 1e-12).all()
    close = np.isclose(theta, ref, rtol=1e-6, atol=1e-9).all(axis=1)
    frac = close.mean()
    pin = K4_PINNED[name]
    dis = np.nonzero(~close)[0].tolist()
    print(f"{name}: {len(ref)} fits, theta within 1e-6 of scipy: {100 * frac:.1f} %, status equal: "
          f"{100 * (status == np.array(ref_status)).mean():.1f} %; disagreeing fits {dis}; scipy's own sensitivity to a "
          f"1e-13 perturbation of x0: {len(pin['sensitive'])} fits, {len(set(dis) & set(pin['sensitive']))} of the "
          f"{len(dis)} disagreeing ones among them")
    assert len(ref) == pin["n_fits"]
    assert set(dis) <= set(pin["disagree"]), f"new disagreeing fits: {sorted(set(dis) - set(pin['disagree']))}"
    assert len(dis) >= len(pin["disagree"]) - 1 or len(dis) == 0      # the list is the measurement, not slack
    # where theta agrees the bookkeeping agrees too
    assert (status[close] == np.array(ref_status)[close]).mean() > 0.98
    assert np.allclose(cost[close], ref_cost[close], rtol=1e-8, atol=1e-12)
    # disagreeing fits: both are local solutions of the same problem; costs stay comparable
    assert np.all(cost[~close] <= ref_cost[~close] * 1.5 + 1e-6)


def test_k4_ragged_fits_up_to_the_size_limit_match_scipy():
    """The recorded histories only hold small neighbourhoods (4-16 points, one row per lane). Well-conditioned
    synthetic fits with 4 ... 2 500 points (up to 79 rows per lane, ragged batch, all in one launch) must follow
    scipy step for step: same status, same number of function evaluations, theta within 1e-9 relative.
    Also: an empty batch returns empty arrays, and a fit beyond the shared-memory limit raises."""
    from pgmorl_b200 import kernels as K
    r = np.random.RandomState(7)
    xs, ys, ws, ubs, ref = [], [], [], [], []
    for k in (5, 33, 64, 200, 620, 17, 100, 32, 31, 97, 4, 333, 2500):
        for rep in range(2):
            x = np.sort(r.uniform(0, 1, k))
            A, a, b, c = r.uniform(5, 60), r.uniform(1, 8), r.uniform(.2, .8), r.uniform(-20, 20)
            y = so.model(x, A, a, b, c) + r.normal(0, 0.5, k)
            w = r.uniform(0.05, 1, k)
            xs.append(x); ys.append(y); ws.append(w); ubs.append(so.upper_bounds(y))
            ref.append(so.fit_scipy(x, y, w, ubs[-1]))
    theta, status, nfev, cost = K.fit_hyperbolic(xs, ys, ws, ubs)
    bad = []
    for i, res in enumerate(ref):
        same = (np.allclose(theta[i], res.x, rtol=1e-9, atol=1e-11) and status[i] == res.status
                and nfev[i] == res.nfev and np.isclose(cost[i], res.cost, rtol=1e-10))
        if not same:
            bad.append((len(xs[i]), theta[i], res.x, status[i], res.status, nfev[i], res.nfev))
    print(f"{len(ref) - len(bad)}/{len(ref)} synthetic fits identical to scipy (status, nfev, theta to 1e-9)")
    assert len(bad) <= 2, bad                      # K = 4 with noise can be one of the chaotic cases
    e = K.fit_hyperbolic([], [], [], [])
    assert e[0].shape == (0, 4) and len(e[1]) == 0
    with pytest.raises(Exception):
        K.fit_hyperbolic([np.linspace(0, 1, 2600)], [np.zeros(2600)], [np.ones(2600)], [so.upper_bounds(np.zeros(2600))])


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_prediction_guided_selection_end_to_end(name, M):
    """Full product path (host candidate generation -> K4 fits -> K5 greedy pick) on the exact state the
    reference had in every generation of the recorded history. Selected (sample, weight) pairs must be
    bit-exact unless a chaotic fit (one that already disagrees with scipy beyond 1e-6, see the K4 test)
    feeds a candidate whose score is involved; any such generation is reported and must be explained by
    a flagged fit."""
    import torch
    from tests.helpers import rebuild_selection_state
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    exact, explained = 0, 0
    torch.set_default_dtype(torch.float64)      # the reference runs with float64 as torch's default (morl/morl.py:33)
    try:
        _run_generations(z, gens, M, name)
    finally:
        torch.set_default_dtype(torch.float32)


def _run_generations(z, gens, M, name):
    from tests.helpers import rebuild_selection_state
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    exact, explained = 0, 0
    for g in range(gens):
        args, graph, pop, ep = rebuild_selection_state(z, g, M)
        np.random.seed(1000 + g)
        template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
        elites, scals, predicted = pop.prediction_guided_selection(args, g, ep, graph, template)
        ids = [s.optgraph_id for s in elites]
        w = np.array([sc.weights.numpy() for sc in scals])
        ref_theta = np.array([z[f"g{g}_fit{i}_theta"] for i in range(int(z[f"g{g}_n_fits"]))])
        fits_ok = np.isclose(pop.last_fits["theta"], ref_theta, rtol=1e-6, atol=1e-9).all(axis=1)
        same = ids == z[f"g{g}_elite_ids"].tolist() and np.array_equal(w, z[f"g{g}_elite_w"])
        if same and fits_ok.all():
            # every fit agrees -> predictions agree to ~1e-9 and the picks are identical
            assert np.allclose(np.array(predicted), z[f"g{g}_predicted"], rtol=1e-6, atol=1e-8)
        if same:
            exact += 1
        else:
            assert not fits_ok.all(), f"generation {g}: selection differs although every fit matches scipy"
            explained += 1
        print(f"{name} gen {g}: picks {'identical' if same else 'DIFFER'}; fits matching scipy "
              f"{int(fits_ok.sum())}/{len(fits_ok)}")
    print(f"{name}: {exact}/{gens} generations bit-exact, {explained} explained by chaotic fits")
    assert exact == gens            # measured on a B200: 7/7 (2 objectives) and 5/5 (3 objectives) generations identical


@pytest.mark.parametrize("M,n_pop,n_ep", [(2, 200, 300), (3, 420, 500)])
def test_selection_at_full_population_size(M, n_pop, n_ep):
    """SURVEY.md section 8(d) sizes: every performance buffer full (2 objectives: 200 samples, ~1 400 candidates,
    archive of 300; 3 objectives: 420 samples, ~2 900 candidates, archive of 500). The product's whole selection runs;
    its scoring is then checked bit for bit against the oracle on the product's own predictions: all rounds and all
    picks with 2 objectives, the first round on a random subset of candidates with 3 (the Python oracle needs ~20 ms per
    candidate there). A sample of the fits is compared with scipy."""
    import torch
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    from synth_envs import make_selection_state
    torch.set_default_dtype(torch.float64)
    try:
        args, graph, pop, ep = make_selection_state(M, n_pop, n_ep, seed=3)
        np.random.seed(7)
        template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
        elites, scals, predicted = pop.prediction_guided_selection(args, 0, ep, graph, template)
    finally:
        torch.set_default_dtype(torch.float32)
    cand = np.array([c["prediction"] for c in pop.last_candidates])
    assert len(pop.sample_batch) == n_pop and len(cand) >= 4 * n_pop and len(elites) == args.num_tasks
    assert np.isfinite(cand).all()
    hv, sp = np.asarray(pop.last_hv), np.asarray(pop.last_sparsity)
    if M == 2:
        best, ohv, osp = so.greedy_select_2d(ep.obj_batch, cand, args.sparsity, args.num_tasks)
        for r in range(args.num_tasks):
            assert np.array_equal(hv[r], ohv[r]) and np.array_equal(sp[r], osp[r]), r
            assert np.array_equal(predicted[r], cand[best[r]]), r
    else:
        idx = np.random.RandomState(1).choice(len(cand), 150, replace=False)
        best, ohv, osp = so.greedy_select_3d(ep.obj_batch, cand[idx], args.sparsity, 1)
        assert np.array_equal(hv[0][idx], ohv[0]) and np.array_equal(sp[0][idx], osp[0])
        # the product's first pick is the arg-max of its own first-round scores (first index on ties)
        score = hv[0] - args.sparsity * sp[0]
        assert np.array_equal(predicted[0], cand[int(np.argmax(score))])
    f = pop.last_fits
    assert len(f["x"]) == n_pop * M
    sel = np.random.RandomState(2).choice(len(f["x"]), 60, replace=False)
    ok = sum(bool(np.isclose(f["theta"][i], so.fit_scipy(f["x"][i], f["y"][i], f["w"][i], f["ub"][i]).x,
                             rtol=1e-6, atol=1e-9).all()) for i in sel)
    print(f"M={M}: {len(cand)} candidates, archive {n_ep}; {ok}/60 sampled fits within 1e-6 of scipy")
    assert ok >= 54


def test_fork_scoring_variant_matches_the_fork_reference():
    """the 2-objective scorer of the reference's fork copy (WorkingMorl/morl/population_2d.py:
    update_ep + InnerHyperVolume + M-D sparsity), selected with args.fork_scoring. (1) With the fork's own candidate
    predictions every round's hv / sparsity array and every pick is bit-exact; (2) the whole product path (host
    candidates -> K4 -> scoring) reproduces the fork's elites and weights on the recorded generations, with the same
    allowance for chaotic fits as the main end-to-end test."""
    import torch
    from tests.helpers import rebuild_selection_state
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    z = np.load(os.path.join(GOLDEN, "selection_2d_fork.npz"))
    gens = int(z["meta"][1])
    alpha = float(z["args_f"][0]); num_tasks = int(z["args"][0])
    exact = 0
    torch.set_default_dtype(torch.float64)
    try:
        for g in range(gens):
            args, graph, pop, ep = rebuild_selection_state(z, g, 2)
            args.fork_scoring = True
            pred = z[f"g{g}_cand_pred"]
            best, hv, sp = pop._greedy_fork(z[f"g{g}_round0_vep"], pred, alpha, num_tasks)
            for r in range(num_tasks):
                assert np.array_equal(hv[r], z[f"g{g}_round{r}_hv"]), (g, r)
                assert np.array_equal(sp[r], z[f"g{g}_round{r}_sparsity"]), (g, r)
                assert np.array_equal(pred[best[r]], z[f"g{g}_predicted"][r]), (g, r)
            np.random.seed(1000 + g)
            template = WeightedSumScalarization(num_objs=2, weights=np.ones(2) / 2)
            elites, scals, predicted = pop.prediction_guided_selection(args, g, ep, graph, template)
            ids = [s.optgraph_id for s in elites]
            w = np.array([sc.weights.numpy() for sc in scals])
            ref_theta = np.array([z[f"g{g}_fit{i}_theta"] for i in range(int(z[f"g{g}_n_fits"]))])
            assert len(pop.last_fits["theta"]) == len(ref_theta)
            fits_ok = np.isclose(pop.last_fits["theta"], ref_theta, rtol=1e-6, atol=1e-9).all(axis=1)
            same = ids == z[f"g{g}_elite_ids"].tolist() and np.array_equal(w, z[f"g{g}_elite_w"])
            if same:
                exact += 1
            else:
                assert not fits_ok.all(), f"generation {g}: selection differs although every fit matches scipy"
            print(f"fork gen {g}: picks {'identical' if same else 'DIFFER'}; fits matching scipy {int(fits_ok.sum())}/{len(fits_ok)}")
    finally:
        torch.set_default_dtype(torch.float32)
    assert exact == gens            # measured on a B200: 5/5 generations identical


def _front_end_vs_oracle(graph, ids, M, cap, force_global_scratch=False):
    """csrc/k4_inputs.cu (through kernels.fit_inputs_launch) + prediction.gaussian_weights against the scalar oracle
    scan of every member: same edge lists, widening steps, x / y / w / ub, bit for bit."""
    from pgmorl_b200 import kernels as K
    from pgmorl_b200.prediction import GraphView, gaussian_weights
    view, ref = GraphView(graph), so.GraphArrays(graph)
    ids = np.asarray(ids, dtype=np.int64)
    front = K.fit_inputs_launch(view.objs, view.parent, view.edge_w, view.edge_dy, ids, cap,
                                force_global_scratch=force_global_scratch)
    coef = gaussian_weights(view, ids, front["steps"], front["source"])
    pack, ub = front["pack"].cpu().numpy(), front["ub"].cpu().numpy()
    klen_f = front["klen_f"].cpu().numpy()
    for b, k in enumerate(ids):
        out, steps, e = so.fit_inputs(ref, int(k), M, cap, with_steps=True)
        assert front["klen"][b] == len(e) and front["steps"][b] == steps, (b, k, front["klen"][b], len(e), front["steps"][b], steps)
        assert np.array_equal(front["source"][b, :len(e)], ref.parent[e])
        for dim, (x, y, w, u) in enumerate(out):
            f = b * M + dim
            assert klen_f[f] == len(e)
            assert np.array_equal(pack[0, f, :len(e)], x) and np.array_equal(pack[1, f, :len(e)], y)
            assert np.array_equal(coef[b, :len(e)], w, equal_nan=True) and np.array_equal(ub[f], u)
    return front


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_fit_input_kernels_match_reference_goldens(name, M):
    """The device front-end on the recorded histories: every generation's fit inputs equal the oracle's (which
    tests/test_host_selection.py pins to the reference's recorded x / y / w / ub)."""
    from tests.helpers import rebuild_selection_state
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    for g in range(gens):
        args, graph, pop, ep = rebuild_selection_state(z, g, M)
        ids = [s.optgraph_id for s in pop.sample_batch]
        front = _front_end_vs_oracle(graph, ids, M, cap=(M != 2))
        if g == gens - 1:          # and directly against the recorded arrays, where every member has test weights
            n_fits = int(z[f"g{g}_n_fits"])
            if n_fits == len(ids) * M:
                pack = front["pack"].cpu().numpy()
                for f in range(n_fits):
                    k = int(front["klen"][f // M])
                    assert np.array_equal(pack[0, f, :k], z[f"g{g}_fit{f}_x"]) and np.array_equal(pack[1, f, :k], z[f"g{g}_fit{f}_y"])


def test_fit_input_kernels_edge_cases():
    """Synthetic opt-graphs that drive every exit of the widening loop: plenty of distinct weights at the first
    threshold; several doublings; fewer than four distinct weights in the whole graph (threshold runs to +inf, every
    edge listed); a member with a zero objective (its neighbourhood stays empty until the product |objs| * inf is
    formed, NaN for the zero component: empty edge list); a graph without edges; duplicated weights that only differ
    below the 1e-5 tolerance; the 3-objective stop at threshold >= 1; more edges than one scan pass (256) handles."""
    from pgmorl_b200.opt_graph import OptGraph
    rng = np.random.RandomState(11)

    def family_graph(M, roots, kids, spread, weights=None, zero_root=False):
        g = OptGraph()
        members = []
        for r in range(roots):
            base = rng.uniform(40.0, 60.0, M) * (1.0 + spread * r)
            if zero_root and r == 0:
                base[0] = 0.0
            root = g.insert(np.ones(M) / M, base.copy(), -1)
            for j in range(kids):
                w = rng.dirichlet(np.ones(M)) if weights is None else np.asarray(weights[j % len(weights)], dtype=np.float64)
                members.append(g.insert(w.copy(), base + rng.normal(0, 2.0, M), root))
        return g, members

    for M in (2, 3):
        for cap in (False, True):
            g, members = family_graph(M, 6, 5, 0.02)                       # dense: done at the first threshold(s)
            _front_end_vs_oracle(g, members + [0], M, cap)
            g, members = family_graph(M, 8, 2, 0.9)                        # sparse: several doublings
            _front_end_vs_oracle(g, members, M, cap)
            two = [np.eye(M)[0] * 0.7 + 0.3 / M, np.eye(M)[1] * 0.7 + 0.3 / M]
            g, members = family_graph(M, 4, 3, 0.5, weights=two)           # 2 distinct weights: runs to +inf unless capped
            front = _front_end_vs_oracle(g, members[:5], M, cap)
            if not cap:
                assert (front["steps"] > 1000).all() and (front["klen"] == 12).all()
            near_dup = [np.full(M, 1.0 / M), np.full(M, 1.0 / M) + 1e-7 * np.arange(M), np.eye(M)[0] * 0.5 + 0.5 / M,
                        np.eye(M)[1] * 0.5 + 0.5 / M, np.eye(M)[1] * 0.2 + 0.8 / M]
            g, members = family_graph(M, 5, 5, 0.05, weights=near_dup)
            _front_end_vs_oracle(g, members, M, cap)
            g, members = family_graph(M, 3, 4, 0.3, zero_root=True)        # member 0 = the root with a zero objective
            _front_end_vs_oracle(g, [0] + members, M, cap)
            g, members = family_graph(M, 2, 400, 0.01)                     # 800 edges: several scan passes per step
            _front_end_vs_oracle(g, members[::37], M, cap)
            _front_end_vs_oracle(g, members[::53], M, cap, force_global_scratch=True)     # per-member scratch in global memory
    # an opt-graph beyond the shared-memory scratch (13 000 nodes): the library switches to the global workspace by itself
    from pgmorl_b200._lib import lib
    g, members = family_graph(2, 1000, 12, 0.002)
    assert lib().pgm_fit_neighbours_workspace_bytes(len(g.objs), 2, 12000, 7) > 0
    _front_end_vs_oracle(g, members[::1801], 2, False)
    g = OptGraph()                                                         # roots only: no edge at all
    for r in range(3):
        g.insert(np.ones(2) / 2, rng.uniform(10, 20, 2), -1)
    from pgmorl_b200 import kernels as K
    from pgmorl_b200.prediction import GraphView
    view = GraphView(g)
    front = K.fit_inputs_launch(view.objs, view.parent, view.edge_w, view.edge_dy, np.arange(3), False)
    assert (front["klen"] == 0).all() and (front["steps"] > 1000).all()
    assert np.array_equal(front["ub"].cpu().numpy(), np.tile([1.0, 20.0, 5.0, 500.0], (6, 1)))


def test_incremental_3d_scorer_matches_oracle_on_adversarial_fronts():
    """K5's 3-objective scorer reuses the round's base front (sorted lists, slice areas, head of the hypervolume sum)
    for every candidate. Fronts and candidates built to hit its special cases -- ties in every coordinate, candidates
    equal to a front point, candidates that dominate many points (removals below and above their own z rank), dominated
    and negative candidates, an empty archive -- scored over several greedy rounds: every round's hv / sparsity array and
    every pick must equal the python restatement of the reference (update_ep + InnerHyperVolume + compute_sparsity)."""
    from pgmorl_b200 import kernels as K
    rng = np.random.RandomState(5)

    def shell(n, r, quant=None):
        v = np.abs(rng.normal(size=(n, 3))) + 0.05
        v = v / np.linalg.norm(v, axis=1, keepdims=True) * r
        if quant:
            v = np.round(v / quant) * quant              # coordinates on a coarse grid: many exact ties
        return v

    cases = []
    ep = shell(40, 50.0)
    cand = np.concatenate([shell(30, 50.0), shell(20, 56.0), shell(10, 44.0), ep[:5].copy(), ep[5:8] + [0.0, 1e-6, 0.0],
                           np.array([[60.0, 60.0, 60.0], [-1.0, 70.0, 70.0], [0.0, 0.0, 0.0], [49.0, 49.0, 0.5]])])
    cases.append((ep, cand, 4))
    ep = shell(60, 30.0, quant=2.0)
    ep = ep[so.get_ep_indices(ep)]
    cand = np.concatenate([shell(40, 31.0, quant=2.0), shell(20, 34.0, quant=1.0), ep[:6].copy()])
    cases.append((ep, cand, 5))
    cases.append((np.zeros((0, 3)), shell(12, 10.0), 3))                                  # empty archive
    cases.append((shell(1, 20.0), np.concatenate([shell(8, 20.0), shell(4, 25.0)]), 3))
    for ep, cand, rounds in cases:
        ep = ep[np.argsort(ep[:, 0], kind="stable")] if len(ep) else ep
        for alpha in (0.0, 1.0):
            best, hv, sp, front = K.select_greedy(ep, cand, alpha, rounds)
            obest, ohv, osp = so.greedy_select_3d(ep, cand, alpha, rounds)
            assert best.tolist() == list(obest), (best, obest)
            for r in range(rounds):
                assert np.array_equal(hv[r], ohv[r]), (r, np.nonzero(hv[r] != ohv[r])[0][:5])
                assert np.array_equal(sp[r], osp[r]), (r, np.nonzero(sp[r] != osp[r])[0][:5])
