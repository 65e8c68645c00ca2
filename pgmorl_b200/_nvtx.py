"""Optional NVTX ranges around the stages of the path (the reference has no tracing at all).
Enabled with PGM_NVTX=1; a no-op otherwise, so the hot loop pays one attribute lookup. The ranges show up in Nsight
Systems / `ncu --nvtx` as mopg.k1_forward, mopg.k2_gae_adv, mopg.k3_ppo_update, rollout.step, selection.fit_inputs,
selection.k4_fits, selection.test_weights, selection.k5_greedy, dist.all_gather_records, dist.migrate_states."""
import contextlib
import os

ENABLED = os.environ.get("PGM_NVTX", "0") not in ("", "0")

if ENABLED:
    import torch

    @contextlib.contextmanager
    def rng(name):
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
else:
    _NULL = contextlib.nullcontext()

    def rng(name):
        return _NULL
