#!/usr/bin/env python
"""How far do two FP32 runs of the SAME long cut drift apart when only the summation order of K3 changes (cluster 16 vs 8)?
Puts the FP32-vs-float64 deviation of tests/test_gpu_run.py::test_longer_reference_cut_drift_is_bounded into perspective.
    gpurun -- python profiles/drift_probe.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_envs  # noqa: E402
from pgmorl_b200 import mopg, morl  # noqa: E402
from pgmorl_b200.layout import NetDims  # noqa: E402


def rows(path):
    return np.array([[float(x) for x in l.replace(";", ",").split(",")] for l in open(path).read().strip().splitlines()])


def run(cluster):
    d = NetDims(17, 6, 2)
    out = tempfile.mkdtemp()
    args = synth_envs.run_args_2d_long(out)
    mopg.set_env_hooks(make_vec_envs=lambda **kw: synth_envs.SeededReplayVecEnv(d, args.num_steps, args.num_processes, [1.3, 0.7], base_seed=500),
                       gym_make=lambda name: synth_envs.ToyEvalEnv(d))
    morl.run(args, device="cuda", cluster=cluster)
    return out


a, b = run(16), run(8)
gold = os.path.join(ROOT, "tests", "golden", "run_2d_long")
for f in ("16/elites/offsprings.txt", "16/ep/objs.txt"):
    x, y, g = rows(os.path.join(a, f)), rows(os.path.join(b, f)), rows(os.path.join(gold, f))
    print(f, "cluster16 vs cluster8: %.2e" % (np.abs(x - y).max() / np.abs(x).max()) if x.shape == y.shape else "shape differs",
          "| cluster16 vs reference: %.2e" % (np.abs(x - g).max() / np.abs(g).max()) if x.shape == g.shape else "shape differs",
          "| cluster8 vs reference: %.2e" % (np.abs(y - g).max() / np.abs(g).max()) if y.shape == g.shape else "shape differs")
