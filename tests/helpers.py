"""Shared helpers for parity tests: golden loading and the tolerance gate."""
import os

import numpy as np

from pgmorl_b200 import synthetic
from pgmorl_b200.layout import NetDims

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_mopg_case(name):
    z = np.load(os.path.join(GOLDEN, name))
    O, A, M, T, N, E, B, n_tasks, traj_seed, total = [int(x) for x in z["meta"]]
    return z, dict(dims=NetDims(O, A, M), T=T, N=N, E=E, B=B, n_tasks=n_tasks, traj_seed=traj_seed,
                   total_num_updates=total, iters=[int(j) for j in z["iters"]],
                   gamma=float(z["gamma_lam"][0]), lam=float(z["gamma_lam"][1]))


def task_traj(meta, j, task):
    traj = synthetic.make_trajectories(meta["n_tasks"], meta["T"], meta["N"], meta["dims"],
                                       seed=meta["traj_seed"] + j)
    return {k: v[task].numpy() for k, v in traj.items()}


def rel_err(a, b):
    """Norm-wise gate of SURVEY.md section 7 (hard part 2): max|a-b| / max|b| per tensor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def rebuild_selection_state(z, g, M):
    """Product-side OptGraph / Population / EP holding exactly the state the reference had when it ran
    prediction_guided_selection in generation g of a golden history."""
    from pgmorl_b200.ep import EP
    from pgmorl_b200.opt_graph import OptGraph
    from synth_envs import ObjSample, SelectionArgs
    from pgmorl_b200 import population_2d, population_3d
    args = SelectionArgs(M)
    graph = OptGraph()
    W, O, prev = z[f"g{g}_graph_w"], z[f"g{g}_graph_objs"], z[f"g{g}_graph_prev"]
    for i in range(len(O)):
        graph.weights.append(W[i].copy()); graph.objs.append(O[i].copy()); graph.prev.append(int(prev[i]))
        graph.delta_objs.append(np.zeros_like(O[i]) if prev[i] == -1 else O[i] - O[int(prev[i])])
        graph.succ.append([])
        if prev[i] != -1:
            graph.succ[int(prev[i])].append(i)
    pop = (population_2d if M == 2 else population_3d).Population(args)
    pop.sample_batch = [ObjSample(o.copy(), int(i)) for o, i in zip(z[f"g{g}_pop_objs"], z[f"g{g}_pop_ids"])]
    ep = EP()
    ep.obj_batch = z[f"g{g}_ep_objs"].copy()
    ep.sample_batch = np.array([ObjSample(o.copy()) for o in ep.obj_batch], dtype=object)
    return args, graph, pop, ep
