#!/usr/bin/env python
"""Device time of the K4 fit launch on the recorded histories (last generation of tests/golden/selection_{2d,3d}.npz):
CUDA events around the launch on torch's current stream, best of 5, plus the distribution of outer iterations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgmorl_b200 import kernels as K  # noqa: E402

for name in ("selection_2d.npz", "selection_3d.npz"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    g = int(z["meta"][1]) - 1
    xs, ys, ws, ubs = [], [], [], []
    for i in range(int(z[f"g{g}_n_fits"])):
        pre = f"g{g}_fit{i}_"
        xs.append(z[pre + "x"]); ys.append(z[pre + "y"]); ws.append(z[pre + "w"]); ubs.append(z[pre + "ub"])
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = K.fit_hyperbolic_launch(xs, ys, ws, ubs)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    theta, status, nfev, cost = K.fit_hyperbolic_collect(h)
    kl = np.array([len(x) for x in xs])
    print(f"{name}: {len(xs)} fits, points per fit {kl.min()}..{kl.max()} (median {int(np.median(kl))}), "
          f"nfev median {int(np.median(nfev))} max {int(nfev.max())}; upload + launch {best:.3f} ms (device, best of 5)")
