import os, sys, time
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tests.helpers import rebuild_selection_state
from pgmorl_b200.scalarization_methods import WeightedSumScalarization
from pgmorl_b200 import kernels as K, prediction
torch.set_default_dtype(torch.float64)
import cProfile, pstats
for name, M in (("selection_2d.npz", 2), ("selection_3d.npz", 3)):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    g = int(z["meta"][1]) - 1
    for rep in range(3):
        args_s, graph, pop, ep = rebuild_selection_state(z, g, M)
        np.random.seed(1000 + g)
        template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
        torch.cuda.synchronize()
        if rep == 2:
            pr = cProfile.Profile(); pr.enable()
        t0 = time.perf_counter()
        pop.prediction_guided_selection(args_s, g, ep, graph, template)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if rep == 2:
            pr.disable()
            print(name, "total ms", 1e3 * (t1 - t0))
            pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# full performance buffers (SURVEY.md section 8(d) sizes), built by synthetic.make_selection_state
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from synth_envs import make_selection_state
for M, n_pop, n_ep in ((2, 200, 300), (3, 420, 500)):
    for rep in range(3):
        args_s, graph, pop, ep = make_selection_state(M, n_pop, n_ep, seed=3)
        np.random.seed(7)
        template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
        torch.cuda.synchronize()
        if rep == 2:
            pr = cProfile.Profile(); pr.enable()
        t0 = time.perf_counter()
        pop.prediction_guided_selection(args_s, 0, ep, graph, template)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if rep == 2:
            pr.disable()
            print(f"full size M={M} n_pop={n_pop} archive={n_ep} candidates={len(pop.last_candidates)} total ms", 1e3 * (t1 - t0))
            pstats.Stats(pr).sort_stats("cumulative").print_stats(16)
