#!/usr/bin/env python
"""Per-phase timeline of the wide-observation tensor-core K3 kernel (csrc/k3_tcw.cuh) from the instrumented build.

    python -m pgmorl_b200.build --trace
    PGM_LIB_PATH=pgmorl_b200/libpgmorl_b200_trace.so python profiles/k3_tcw_trace.py [P] [cluster]

clock64 marks of thread 0 (the MMA issuer) of every CTA of task 0, optimiser steps 8 and 9.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402
from pgmorl_b200.population_state import PopulationMOPG  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CL = int(sys.argv[2]) if len(sys.argv) > 2 else 128     # 32 / 64 / 128: 2 / 4 / 8 CTAs per task
d = ENV_SHAPES["humanoid"]
T, N, E, B = 2048, 8, 10, 32
pop = PopulationMOPG(d, P, T, N, gamma=0.99, cluster=CL)
traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
for p in range(P):
    pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
pop.set_lr(3e-4)
pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
for _ in range(2):
    pop.step()
torch.cuda.synchronize()
r256 = lambda b: (b + 255) // 256 * 256
off_tr = r256(P * 2 * 64 * 4) + r256(P * 16 * 4) + r256(P * 64 * 4)
off = (-pop.workspace.data_ptr()) % 256
NC = CL // 16
n = P * NC * 2 * 48
tr = pop.workspace[off + off_tr: off + off_tr + n * 8].cpu().numpy().view(np.int64).reshape(P * NC, 2, 48)
names = ["idx", "sync+issue loads 0,1"] + [f"G1 block {b}" for b in range(6)] + [
    "wait G1", "E1", "sync+issue G2", "wait G2", "E2", "sync+issue G3", "wait G3", "E3", "sync+issue G4,GWh", "wait G4", "E4+db2",
    "sync+issue G5,GW2", "wait G5", "wait GW2", "E5+x0 load+sync", "issue G1X 0,1", "wait G1X 0", "load x4,x5+sync", "issue G1X 2",
    "wait G1X", "tail sync", "TMEM->GR", "cluster bar 1", "slice reduce+ssq", "cluster bar 2", "Adam+publish", "cluster bar 3", "release fence"]
for cta in range(NC):
    c = tr[cta, 1, :35].astype(np.int64)
    ntile_marks = c[0] != 0
    print(f"cta {cta} ({'actor' if cta < NC // 2 else 'critic'}): step = {int(tr[cta, 1, 34] - tr[cta, 0, 34])} cycles"
          + ("" if ntile_marks else " (no tile)"))
    prev = tr[cta, 0, 34]
    for i in list(range(34)) + [35, 34]:          # mark 35 (after the release fence) sits between 33 and 34
        v = tr[cta, 1, i]
        if v == 0:
            continue
        print(f"    {names[i]:>24s} {int(v - prev):7d}")
        prev = v
