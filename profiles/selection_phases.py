"""Wall-clock phases of prediction_guided_selection on the four selection workloads of bench.py's `selection` leg
(recorded 2-D / 3-D histories, full-size 2-D / 3-D buffers): where a generation's selection time goes.
    python profiles/selection_phases.py            (on a GPU box)
Phases (host clock, the device work they wait for included): fit inputs (opt-graph view, neighbourhood kernels, Gaussian
weights, K4 launch), test weights (host, runs under K4), wait for K4, model evaluation, K5 greedy loop, elites."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import rebuild_selection_state            # noqa: E402
from synth_envs import make_selection_state            # noqa: E402
import pgmorl_b200.kernels as K                        # noqa: E402
import pgmorl_b200.prediction as P                     # noqa: E402
from pgmorl_b200.scalarization_methods import WeightedSumScalarization   # noqa: E402

T = {}


def timed(mod, name, label, sync=False):
    f = getattr(mod, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        if sync:
            torch.cuda.synchronize()
        T[label] = T.get(label, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, g)


timed(P, "launch_fits", "fit_inputs+launch")
timed(K, "fit_inputs_launch", "  of which neighbour kernels + copies")
timed(P, "gaussian_weights", "  of which gaussian weights (host)")
timed(K, "fit_hyperbolic_collect", "wait for K4")
timed(K, "select_greedy", "K5 greedy")
timed(P, "finish_predictions", "finish_predictions (incl. wait)")


def run(label, make):
    best = None
    for rep in range(5):
        args_s, graph, pop, ep = make()
        M = args_s.obj_num
        np.random.seed(7)
        template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
        orig = pop._test_weights_batch

        def tw(*a, **k):
            t0 = time.perf_counter(); r = orig(*a, **k); T["test weights (host)"] = time.perf_counter() - t0; return r
        pop._test_weights_batch = tw
        T.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pop.prediction_guided_selection(args_s, 0, ep, graph, template)
        torch.cuda.synchronize()
        tt = time.perf_counter() - t0
        if rep and (best is None or tt < best[0]):
            best = (tt, dict(T), pop)
    nfev = best[2].last_fits["nfev"]
    print(f"{label}: {1e3 * best[0]:.2f} ms  (n_pop {len(best[2].sample_batch)}, candidates {len(best[2].last_candidates)}, "
          f"fits {len(nfev)}, nfev max {int(nfev.max())} mean {nfev.mean():.0f})")
    for k, v in best[1].items():
        print(f"    {k:42s} {1e3 * v:7.2f} ms")


if __name__ == "__main__":
    torch.set_default_dtype(torch.float64)
    for name, M in (("selection_2d.npz", 2), ("selection_3d.npz", 3)):
        z = np.load(os.path.join(ROOT, "tests", "golden", name))
        g = int(z["meta"][1]) - 1
        run(f"{M}d recorded", lambda: rebuild_selection_state(z, g, M))
    for M, n_pop, n_ep in ((2, 200, 300), (3, 420, 500)):
        run(f"{M}d full", lambda: make_selection_state(M, n_pop, n_ep, seed=3))
