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
d = ENV_SHAPES["halfcheetah"]
T, N, E, B = 2048, 4, 10, 32
S = T * N
pop = PopulationMOPG(d, P, T, N, cluster=32)
traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
for p in range(P):
    pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
pop.set_lr(3e-4)
pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
for _ in range(3):
    pop.step()
torch.cuda.synchronize()
r256 = lambda b: (b + 255) // 256 * 256
off_tr = r256(P * S * 32 * 4) + r256(P * 2 * 64 * 4) + r256(P * 16 * 4) + r256(P * 64 * 4) + r256(P * 2 * 2 * 6144 * 4)
off = (-pop.workspace.data_ptr()) % 256
n = P * 2 * 2 * 2 * 2 * 40
raw = pop.workspace[off + off_tr: off + off_tr + n * 8].cpu().numpy().view(np.int64)
tr = raw.reshape(P * 2, 2, 2, 2, 40)      # cta, group, who, step, mark
names = {0: "tile0", 1: "rec+waitB", 2: "x->tmem/smem", 3: "bar", 4: "issueG1", 5: "waitG1", 6: "E1", 7: "bar", 8: "issueG2",
         9: "waitG2", 10: "E2", 11: "bar", 12: "issueG3", 13: "waitG3", 14: "E3", 15: "bar", 16: "issueG4+GWh", 17: "waitG4",
         18: "E4", 19: "bar", 20: "issueG5+GW2", 21: "waitG5", 22: "E5a", 23: "waitGW2", 24: "E5b", 25: "bar",
         26: "issueG1X", 27: "-", 28: "waitG1X", 29: "sync", 30: "readout+ssq", 31: "ldmv", 32: "clusterbar", 33: "adam",
         34: "sync"}
for cta in (0, 1):
    for g in (0, 1):
        for who in (0, 1):
            c = tr[cta, g, who, 1, :35].astype(np.int64)
            dur = np.diff(c)
            print(f"cta {cta} ({'actor' if cta == 0 else 'critic'}) group {g} thread r={'0 (issuer)' if who == 0 else '64'}:"
                  f" step = {int(tr[cta, g, who, 1, 34] - tr[cta, g, who, 0, 34])} cycles")
            print("   ", " ".join(f"{names[i + 1]}={int(x)}" for i, x in enumerate(dur)))
