"""Generate tests/golden/selection_2d_fork.npz: the 2-objective selection of the reference's FORK copy
(WorkingMorl/morl/population_2d.py), whose scorer differs from the top-level copy: candidates are scored with
utils.update_ep + InnerHyperVolume (round(hv, 4)) + utils.compute_sparsity instead of the closed 2-D forms
(WorkingMorl/morl/population_2d.py:207-226), the neighbourhood search stops at threshold >= 1 (:62) and models
without usable data predict no change (:112-117).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_selection_fork.py

The fork's modules are imported under their own names FIRST, so that the recording harness of
make_golden_selection.py (same file format as selection_2d.npz) binds to them unmodified.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FORK = "/root/reference/WorkingMorl/morl"
sys.dont_write_bytecode = True
sys.path.insert(0, FORK)
import population_2d as fork2d  # noqa: E402,F401
import ep  # noqa: E402,F401
import opt_graph  # noqa: E402,F401
import scalarization_methods  # noqa: E402,F401
import utils  # noqa: E402,F401

assert fork2d.__file__.startswith(FORK) and ep.__file__.startswith(FORK)
sys.path.insert(0, HERE)
import make_golden_selection as harness  # noqa: E402

assert harness.ref2d is fork2d

if __name__ == "__main__":
    gens = 5
    blob = harness.record_history(2, gens, 11)
    path = os.path.join(HERE, "selection_2d_fork.npz")
    np.savez_compressed(path, **blob)
    print("selection_2d_fork.npz", os.path.getsize(path) // 1024, "KiB", "fits/gen",
          [int(blob[f"g{g}_n_fits"]) for g in range(gens)])
