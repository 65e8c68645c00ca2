#!/usr/bin/env python
"""K3 time per MOPG iteration vs population size: best FP32 FFMA cluster configuration (the library's rule before the
tensor-core path existed) against the tensor-core path (cluster 32).
    gpurun -- python profiles/k3_sweep.py [P ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from pgmorl_b200 import kernels as K  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402
from pgmorl_b200.population_state import PopulationMOPG  # noqa: E402

Ps = [int(x) for x in sys.argv[1:]] or [6, 8, 12, 18, 24, 37, 64, 74, 128, 148, 256]
d = ENV_SHAPES["walker2d"]
T, N, E, B = 2048, 4, 10, 32


def ffma_cluster(P, sms=148):
    """CTAs per task the FFMA planner picks (csrc/k3_ppo.cu k3_plan)"""
    C, cmax = 2, (16 if P <= 7 else 8)
    while C < cmax and P * C * 2 <= sms:
        C *= 2
    return C


for P in Ps:
    traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
    row = []
    for cluster in (ffma_cluster(P), 32, 64):
        pop = PopulationMOPG(d, P, T, N, cluster=cluster)
        for p in range(P):
            pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
        pop.set_lr(3e-4)
        pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
        pop.step()
        torch.cuda.synchronize()
        # time K3 alone (pack + PPO update) on the state left by the step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(3):
            e0.record()
            K.ppo_update(pop.params, pop.adam_m, pop.adam_v, pop.adam_step, pop.lr, pop.obs, pop.action, pop.logp,
                         pop.value, pop.returns.view(P, T * N, d.obj), pop.adv.view(P, T * N), pop.perm, B, d,
                         hyper=pop.hyper, workspace=pop.workspace, cluster=cluster, losses=pop.losses)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(min(ts))
        del pop
    print(f"P={P:4d}  ffma(cluster {ffma_cluster(P):2d}) {row[0]:8.3f} ms   tensor-core {row[1]:8.3f} ms   tensor-core x2 CTAs {row[2]:8.3f} ms   ratio {row[0] / min(row[1], row[2]):.2f}   "
          f"TC env-steps/s (K3 only) {P * T * N / min(row[1], row[2]) * 1e3 / 1e6:.1f} M")
