#!/usr/bin/env python
"""Per-phase timeline of the K3 PPO kernel from the instrumented build (profiling aid).

    python -m pgmorl_b200.build --trace
    PGM_LIB_PATH=pgmorl_b200/libpgmorl_b200_trace.so python profiles/k3_trace.py [cluster] [P]

Marks (clock64 of thread 0, steps 8..11 of the launch): 0 step start, 1 records landed, 2 forward done,
3 loss done, 4 phase A done, 5 phase B done, 6 phase C done, 7 partial written, 8 after barrier 1,
9 reduce + ssq done, 10 after barrier 2, 11 Adam done.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from pgmorl_b200 import kernels as K  # noqa: E402
from pgmorl_b200._lib import lib  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402
from pgmorl_b200.population_state import PopulationMOPG  # noqa: E402

C = int(sys.argv[1], 0) if len(sys.argv) > 1 else 8      # OR 0x100 in for the two-barrier tail
P = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d = ENV_SHAPES["halfcheetah"]
T, N, E, B = 2048, 4, 10, 32
pop = PopulationMOPG(d, P, T, N, cluster=C)
traj, eps, perm, w, ov, flats = synthetic_inputs(d, P, T, N, E, 1)
for p in range(P):
    pop.load_task(p, flats[p], weights=w[p], obj_var=ov[p])
pop.set_lr(3e-4)
pop.upload(traj["obs"], traj["rewards"], traj["masks"], traj["bad_masks"], eps.float(), perm.int())
for _ in range(3):
    pop.step()
torch.cuda.synchronize()
import ctypes
fn = lib().pgm_ppo_last_trace_offset            # exported by the instrumented build only
fn.restype = ctypes.c_size_t
t_off = fn()
tr_bytes = P * 16 * 4 * 16 * 2 * 8
off = (-pop.workspace.data_ptr()) % 256
raw = pop.workspace[off + t_off: off + t_off + tr_bytes].cpu().numpy().view(np.int64)
CF, C = C, C & 0xFF
tr = raw[: P * C * 4 * 16 * 2].reshape(P * C, 4, 16, 2)
names = ["gather", "fwd", "loss", "phA", "phB", "phC", "write", "bar1", "reduce", "bar2", "adam"]
print(f"cluster {C}: per-phase cycles (clock64), task 0, step index 9 (2nd traced step)")
for r in range(C):
    c = tr[r, 1, :12, 0]
    dur = np.diff(c)
    print(f"rank {r:2d} ({'actor ' if r < C // 2 else 'critic'}):", " ".join(f"{n}={int(x):5d}" for n, x in zip(names, dur)),
          f"| step={int(tr[r, 2, 0, 0] - tr[r, 1, 0, 0])}")
g = tr[:C, 1, :12, 1].astype(np.float64)
g -= g.min()
print("globaltimer (ns, relative) at marks 7 (arrive barrier 1) and 8 (leave):")
for r in range(C):
    print(f"rank {r:2d}: start={g[r,0]:7.0f} arrive1={g[r,7]:7.0f} leave1={g[r,8]:7.0f} arrive2={g[r,9]:7.0f} leave2={g[r,10]:7.0f} end={g[r,11]:7.0f}")
