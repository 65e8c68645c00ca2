#!/usr/bin/env python
"""Per-phase timeline of the K3 tensor-core kernel from the instrumented build (profiling aid).

    python -m pgmorl_b200.build --trace
    PGM_LIB_PATH=pgmorl_b200/libpgmorl_b200_trace.so python profiles/k3_tc_trace.py [P]

clock64 marks of threads r == 0 (MMA issuer) and r == 64 of each 128-thread group, optimiser steps 8 and 9.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402
from pgmorl_b200.population_state import PopulationMOPG  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 6
CL = int(sys.argv[2]) if len(sys.argv) > 2 else 32      # 32: 2 CTAs per task, 64: 4 CTAs per task
d = ENV_SHAPES["halfcheetah"]
T, N, E, B = 2048, 4, 10, 32
S = T * N
pop = PopulationMOPG(d, P, T, N, cluster=CL)
traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
for p in range(P):
    pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
pop.set_lr(3e-4)
pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
for _ in range(3):
    pop.step()
torch.cuda.synchronize()
r256 = lambda b: (b + 255) // 256 * 256
off_tr = r256(P * S * 32 * 4) + r256(P * 2 * 64 * 4) + r256(P * 16 * 4) + r256(P * 64 * 4) + r256(P * 2 * 2 * 2 * 6144 * 4)
off = (-pop.workspace.data_ptr()) % 256
NC = CL // 16          # CTAs per task
n = P * NC * 2 * 2 * 2 * 48
raw = pop.workspace[off + off_tr: off + off_tr + n * 8].cpu().numpy().view(np.int64)
tr = raw.reshape(P * NC, 2, 2, 2, 48)      # cta, group, who, step, mark
names = {1: "waitB+X", 2: "sync+issueG1", 3: "waitG1(0)", 4: "E1(0)", 5: "sync+issueG2(0)", 6: "waitG1(1)", 7: "E1(1)",
         8: "sync+issueG2(1)", 9: "waitG2(0)", 10: "E2(0)", 11: "sync+issueG3(0)", 12: "waitG2(1)", 13: "E2(1)",
         14: "sync+issueG3(1)", 15: "waitG3", 16: "E3", 17: "sync+issueG4,GWh", 18: "waitG4(0)", 19: "E4(0)",
         20: "sync+issueG5,GW2(0)", 21: "waitG4(1)", 22: "E4(1)", 23: "sync+issueG5,GW2(1)", 24: "waitG5(0)", 25: "E5a(0)",
         26: "waitGW2(0)", 27: "E5b+sync+issueG1X(0)", 28: "waitG5(1)", 29: "E5a(1)", 30: "waitGW2(1)",
         31: "E5b+sync+issueG1X(1)", 32: "waitG1X(0)", 33: "drain dW2,dWh + waitG1X(1)", 34: "drain G1X + sums", 35: "ssq+ldmv", 36: "clusterbar", 37: "adam",
         38: "sync"}
for cta in range(NC):
    for g in (0, 1):
        for who in (0, 1):
            c = tr[cta, g, who, 1, :39].astype(np.int64)
            dur = np.diff(c)
            print(f"cta {cta} ({'actor' if cta < NC // 2 else 'critic'}) group {g} thread r={'0 (issuer)' if who == 0 else '64'}:"
                  f" step = {int(tr[cta, g, who, 1, 38] - tr[cta, g, who, 0, 38])} cycles")
            print("   ", " ".join(f"{names[i + 1]}={int(x)}" for i, x in enumerate(dur)))
            x = tr[cta, g, who, 1]
            print("    X phase detail: waits", int(x[39] - x[0]), "x-stores", int(x[40] - x[39]), "scalars", int(x[41] - x[40]),
                  "load_rec", int(x[42] - x[41]), "index+rest", int(x[1] - x[42]))
