"""Host-side selection logic (no GPU): performance buffers, candidate weights, opt-graph neighbourhood
search and fit inputs reproduce the unmodified reference bit for bit (goldens from
tests/golden/make_golden_selection.py)."""
import os

import numpy as np
import pytest

import synth_envs
from pgmorl_b200 import synthetic
from pgmorl_b200.prediction import GraphView, fit_inputs
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
        view = GraphView(graph)
        i = 0
        for k in fit_nodes:
            for x, y, w, ub in fit_inputs(view, k, M, cap_threshold=(M != 2)):
                pre = f"g{g}_fit{i}_"
                assert np.array_equal(x, z[pre + "x"]) and np.array_equal(y, z[pre + "y"])
                assert np.array_equal(w, z[pre + "w"]) and np.array_equal(ub, z[pre + "ub"])
                i += 1
        assert i == int(z[f"g{g}_n_fits"])
