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
