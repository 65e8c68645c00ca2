"""Host-side selection logic (no GPU): performance buffers, candidate weights, opt-graph neighbourhood
search and fit inputs reproduce the unmodified reference bit for bit (goldens from
tests/golden/make_golden_selection.py)."""
import os

import numpy as np
import pytest

import synth_envs
from pgmorl_b200 import synthetic
from oracle import selection_oracle as so
from pgmorl_b200.prediction import GraphView, gaussian_weights
from tests.helpers import rebuild_selection_state

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_generate_weights_known_answers():
    from pgmorl_b200.utils import generate_weights_batch_dfs
    z = np.load(os.path.join(GOLDEN, "selection_kats.npz"))
    for name, (M, d) in {"w2": (2, 0.2), "w3a": (3, 0.25), "w3b": (3, 0.125), "w3c": (3, 1.0 / 19)}.items():
        wb = []
        generate_weights_batch_dfs(0, M, 0.0, 1.0, d, [], wb)
        assert np.array_equal(np.array(wb), z["weights_" + name])
        assert np.array_equal(synthetic.simplex_weights(M, d), z["weights_" + name])
    assert len(z["weights_w2"]) == 6 and len(z["weights_w3a"]) == 15 and len(z["weights_w3b"]) == 45 and len(z["weights_w3c"]) == 210


def test_opt_graph_insert_semantics():
    import torch
    from pgmorl_b200.opt_graph import OptGraph
    g = OptGraph()
    r = g.insert(torch.tensor([0.2, 0.8], dtype=torch.float64), np.array([1.0, 2.0]), -1)
    c = g.insert(np.array([0.2, 0.8]), np.array([1.5, 4.0]), r)
    assert (r, c) == (0, 1) and g.succ == [[1], []] and g.prev == [-1, 0]
    assert np.allclose(np.asarray(g.weights[0]), np.array([0.2, 0.8]) / np.linalg.norm([0.2, 0.8]))
    assert np.array_equal(g.delta_objs[0], [0, 0]) and np.array_equal(g.delta_objs[1], [0.5, 2.0])


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_population_update_matches_reference(name, M):
    """Re-binning population(g-1) + offspring(g) gives the reference's population(g), in order."""
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    for g in range(1, gens):
        args, graph, pop, ep = rebuild_selection_state(z, g - 1, M)
        n_prev_nodes = len(z[f"g{g - 1}_graph_objs"])
        O = z[f"g{g}_graph_objs"]
        offspring = [synth_envs.ObjSample(O[i].copy(), i) for i in range(n_prev_nodes, len(O))]   # one node per task
        pop.update(offspring)
        assert [s.optgraph_id for s in pop.sample_batch] == z[f"g{g}_pop_ids"].tolist()


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_candidate_weights_and_fit_inputs_bit_exact(name, M):
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    for g in (0, gens - 1):
        args, graph, pop, ep = rebuild_selection_state(z, g, M)
        np.random.seed(1000 + g)
        weights, nodes, fit_nodes = [], [], []
        for s in pop.sample_batch:
            if M == 2:
                tw = pop._test_weights(graph, s, args.num_weight_candidates)
            else:
                from pgmorl_b200.utils import generate_weights_batch_dfs
                grid = []
                generate_weights_batch_dfs(0, M, 0.0, 1.0, args.delta_weight / 2.0, [], grid)
                tw = pop._test_weights(args, graph, s, grid)
            if len(tw):
                fit_nodes.append(s.optgraph_id)
            for w in tw:
                weights.append(np.asarray(w, dtype=np.float64)); nodes.append(s.optgraph_id)
        assert nodes == z[f"g{g}_cand_node"].tolist()
        assert np.array_equal(np.array(weights), z[f"g{g}_cand_weight"])
        view = so.GraphArrays(graph)
        i = 0
        for k in fit_nodes:
            for x, y, w, ub in so.fit_inputs(view, k, M, cap_threshold=(M != 2)):
                pre = f"g{g}_fit{i}_"
                assert np.array_equal(x, z[pre + "x"]) and np.array_equal(y, z[pre + "y"])
                assert np.array_equal(w, z[pre + "w"]) and np.array_equal(ub, z[pre + "ub"])
                i += 1
        assert i == int(z[f"g{g}_n_fits"])


@pytest.mark.parametrize("name,M", [("selection_2d.npz", 2), ("selection_3d.npz", 3)])
def test_batched_front_end_matches_reference_bits(name, M):
    """The whole-population array forms the product uses (flat opt-graph view, batched test weights, Gaussian point
    weights from the listed edges) against the reference's recorded candidates and fit inputs, bit for bit."""
    from pgmorl_b200.utils import generate_weights_batch_dfs, rownorm
    z = np.load(os.path.join(GOLDEN, name))
    gens = int(z["meta"][1])
    for g in (0, gens // 2, gens - 1):
        args, graph, pop, ep = rebuild_selection_state(z, g, M)
        np.random.seed(1000 + g)
        view, ref = GraphView(graph), so.GraphArrays(graph)
        for a in ("objs", "parent", "child", "edge_w", "edge_dy"):
            assert np.array_equal(getattr(view, a), getattr(ref, a)), a
        ids = np.array([s.optgraph_id for s in pop.sample_batch], dtype=np.int64)
        if M == 2:
            tests, counts = pop._test_weights_batch(view, ids, args.num_weight_candidates)
        else:
            grid = []
            generate_weights_batch_dfs(0, M, 0.0, 1.0, args.delta_weight / 2.0, [], grid)
            grid_arr = np.array(grid, dtype=np.float64)
            tests, counts = pop._test_weights_batch(args, view, ids, grid_arr, rownorm(grid_arr))
        nodes = np.repeat(ids, counts)
        weights = np.concatenate([tests[b, :counts[b]] for b in range(len(ids))])
        assert nodes.tolist() == z[f"g{g}_cand_node"].tolist()
        assert np.array_equal(weights, z[f"g{g}_cand_weight"])
        # Gaussian point weights from (steps, source nodes of the listed edges) -- here taken from the oracle's scan,
        # on the device they come from csrc/k4_inputs.cu (tests/test_gpu_selection.py)
        fit_ids = ids[counts > 0]
        scans = [so.fit_inputs(ref, int(k), M, cap_threshold=(M != 2), with_steps=True) for k in fit_ids]
        kmax = max(len(e) for _, _, e in scans)
        source = np.zeros((len(fit_ids), kmax), dtype=np.int64)
        for b, (_, _, e) in enumerate(scans):
            source[b, :len(e)] = ref.parent[e]
        coef = gaussian_weights(view, fit_ids, [st for _, st, _ in scans], source)
        i = 0
        for b, (out, _, e) in enumerate(scans):
            for dim in range(M):
                assert np.array_equal(coef[b, :len(e)], z[f"g{g}_fit{i}_w"])
                assert np.array_equal(view.edge_w[e, dim], z[f"g{g}_fit{i}_x"])
                i += 1
        assert i == int(z[f"g{g}_n_fits"])


def test_rowwise_helpers_give_the_scalar_bits():
    """utils.rowdot / rownorm / pow2 against the scalar numpy / Python expressions of the reference."""
    from pgmorl_b200 import utils
    rng = np.random.RandomState(3)
    for m in (2, 3):
        a = rng.rand(2000, m) * rng.choice([1e-4, 1.0, 80.0], size=(2000, 1))
        b = rng.rand(2000, m) - 0.3
        assert np.array_equal(utils.rowdot(a, b), np.array([np.dot(p, q) for p, q in zip(a, b)]))
        assert np.array_equal(utils.rownorm(a), np.array([np.linalg.norm(p) for p in a]))
        assert np.array_equal(utils.rowdot(a[:7, None, :], b[None, :5, :]),
                              np.array([[np.dot(p, q) for q in b[:5]] for p in a[:7]]))
    x = np.abs(rng.randn(5000)) * 4.0
    assert np.array_equal(utils.pow2(x), np.array([float(v) ** 2 for v in x]))
    assert np.array_equal(np.exp(-utils.pow2(x) / 2.0), np.array([np.exp(-(float(v) ** 2) / 2.0) for v in x]))
    # both implementations behind the switch agree with each other
    assert np.array_equal(utils._rowdot_loop(a, b), utils.rowdot(a, b)) and np.array_equal(utils._pow2_loop(x), utils.pow2(x))


def test_opt_graph_flat_cache_and_graph_view_stay_in_step_with_the_lists():
    """OptGraph.flat() mirrors the node lists incrementally (the selection front-end reads it every generation);
    GraphView's edge order is the reference's walk over opt_graph.succ."""
    from pgmorl_b200.opt_graph import OptGraph
    rng = np.random.RandomState(0)
    g = OptGraph()
    for step in range(5):
        for _ in range(40):                                  # grows past the initial capacity: the arrays are re-allocated
            n = len(g.objs)
            prev = -1 if n == 0 or rng.rand() < 0.2 else int(rng.randint(n))
            g.insert(rng.dirichlet(np.ones(3)), rng.uniform(1, 9, 3), prev)
        W, O, D, P = g.flat()
        assert np.array_equal(O, np.array(g.objs)) and np.array_equal(D, np.array(g.delta_objs))
        assert np.array_equal(W, np.array(g.weights)) and P.tolist() == g.prev
        view, ref = GraphView(g), so.GraphArrays(g)
        for a in ("objs", "parent", "child", "edge_w", "edge_dy"):
            assert np.array_equal(getattr(view, a), getattr(ref, a)), a
        nodes = rng.randint(len(g.objs), size=7)
        member, succ = view.successors_of(nodes)
        assert succ.tolist() == [s for k in nodes for s in g.succ[k]]
        assert member.tolist() == [i for i, k in enumerate(nodes) for _ in g.succ[k]]


def test_candidates_container_and_lazy_fit_record():
    from pgmorl_b200.prediction import Candidates, FitRecord
    tests = np.arange(2 * 3 * 2, dtype=np.float64).reshape(2, 3, 2)
    pred = tests + 100.0
    c = Candidates(["a", "b"], tests, [2, 0], pred)
    assert len(c) == 2 and c[1]["sample"] == "a" and np.array_equal(c[1]["weight"], tests[0, 1])
    assert np.array_equal(c.prediction, pred[0, :2]) and [x["sample"] for x in c] == ["a", "a"]
    calls = []
    rec = FitRecord(lambda: calls.append(1) or dict(x=[1], y=[2], w=[3], ub=[4]), theta=np.zeros((1, 4)))
    assert "x" not in rec and rec["theta"].shape == (1, 4) and not calls
    assert rec["x"] == [1] and rec["ub"] == [4] and calls == [1]          # one device -> host copy, on first access
    with pytest.raises(KeyError):
        rec["nope"]
