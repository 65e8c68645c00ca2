#!/usr/bin/env python
"""cProfile of the generation boundary (record handling + opt-graph / archive / population update + prediction-guided
selection) at the task counts of the weak-scaling runs, on ONE GPU (the redundant per-rank work does not depend on W).
    gpurun -- python profiles/boundary_profile.py [n_tasks ...]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import GenerationBoundary  # noqa: E402
from pgmorl_b200.layout import ENV_SHAPES  # noqa: E402

d = ENV_SHAPES["halfcheetah"]
for n_tasks in [int(a) for a in sys.argv[1:]] or [6, 48]:
    gb = GenerationBoundary(d, n_tasks, 1, 0, pop=None)
    for _ in range(3):
        gb.prepare(); gb.run()
    rows = []
    for _ in range(3):
        gb.prepare(); rows.append(gb.run())
    print(f"n_tasks {n_tasks}:", {k: round(1e3 * min(r[k] for r in rows), 2) for k in ("exchange_s", "bookkeeping_s", "selection_s")},
          "ms; n_pop", rows[-1]["n_pop"], "archive", rows[-1]["archive"])
    gb.prepare()
    pr = cProfile.Profile()
    pr.enable(); gb.run(); pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
    print("\n".join(l for l in s.getvalue().splitlines()[4:40]))
