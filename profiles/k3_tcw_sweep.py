#!/usr/bin/env python
"""Humanoid-shape population iteration (K1 + K2 + K3) against tasks per GPU: shows where the wide tensor-core kernel's
cluster variants (8 / 4 / 2 CTAs per task) stop being co-resident. usage: k3_tcw_sweep.py [P ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402
from pgmorl_b200.population_state import PopulationMOPG  # noqa: E402

d = ENV_SHAPES["humanoid"]
T, N, E = 2048, 8, 10
for P in [int(x) for x in sys.argv[1:]] or [8, 12, 14, 15, 16, 18, 24, 32, 37]:
    pop = PopulationMOPG(d, P, T, N, cluster=0)
    traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
    for p in range(P):
        pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
    pop.set_lr(3e-4)
    pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
    for _ in range(2):
        pop.step()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pop.step(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"P={P:3d}: {best:7.2f} ms per iteration, {P * T * N / best / 1e3:7.1f} M env-steps/s", flush=True)
    del pop
    torch.cuda.empty_cache()
