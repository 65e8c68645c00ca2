"""Pin oracle/mopg_oracle.py (float64 numpy restatement) against outputs of the
unmodified reference stored in tests/golden/mopg_*.npz (made by make_golden_mopg.py)."""
import numpy as np
import pytest

from oracle import mopg_oracle as orc
from pgmorl_b200 import synthetic
from tests.helpers import load_mopg_case, rel_err, task_traj

CASES = ["mopg_walker_small.npz", "mopg_hopper3_small.npz", "mopg_humanoid_small.npz"]


def run_case(name, tasks=None, tol=1e-9):
    z, meta = load_mopg_case(name)
    d = meta["dims"]
    dims = (d.obs, d.act, d.obj)
    for task in (tasks if tasks is not None else range(meta["n_tasks"])):
        flat = z[f"t{task}_init"].copy()
        # product-side init reproduces the reference's Policy init (LAPACK QR may differ by an
        # ulp with thread count / CPU, hence a tolerance instead of array_equal)
        assert rel_err(synthetic.init_policy_flat(d, seed=1000 + task).numpy(), flat) < 1e-14
        m = np.zeros_like(flat); v = np.zeros_like(flat); step = 0
        for k, j in enumerate(meta["iters"]):
            traj = task_traj(meta, j, task)
            eps, perm = synthetic.host_rng_streams(j, meta["T"], meta["N"], d.act, meta["E"])
            lr = synthetic.linear_lr(3e-4, j, 1.0, meta["total_num_updates"])
            out = orc.mopg_iteration(flat, m, v, step, lr, dims, traj, eps.numpy(), perm.numpy(),
                                     z[f"t{task}_weights"], z[f"t{task}_obj_var"],
                                     gamma=meta["gamma"], lam=meta["lam"], num_mini_batch=meta["B"])
            step = out["step"]
            pre = f"t{task}_i{k}_"
            assert abs(lr - float(z[pre + "lr"])) < 1e-18
            assert step == int(z[pre + "adam_step"])
            for key in ("value", "action", "logp", "returns", "adv"):
                if pre + key in z:
                    assert rel_err(out[key], z[pre + key]) < 1e-12, key
            assert rel_err(out["losses"], z[pre + "losses"]) < tol
            assert rel_err(flat, z[pre + "params"]) < tol
            assert rel_err(m, z[pre + "adam_m"]) < tol
            assert rel_err(v, z[pre + "adam_v"]) < tol


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_small(name):
    run_case(name)


def test_oracle_matches_reference_full_c2():
    # full HalfCheetah-shape iteration (320 Adam steps on 8192 samples), one task
    run_case("mopg_halfcheetah_full.npz", tasks=[1], tol=1e-8)


def test_torch_port_matches_reference():
    """oracle/mopg_torch_port.py (the CPU-baseline leg of bench.py) reproduces the reference too."""
    import torch
    from oracle.mopg_torch_port import PortPolicy, mopg_iteration_port
    z, meta = load_mopg_case("mopg_walker_small.npz")
    d = meta["dims"]
    task = 1
    pol = PortPolicy(z[f"t{task}_init"], d.obs, d.act, d.obj)
    opt = torch.optim.Adam(pol.ordered(), lr=3e-4, eps=1e-5)
    for k, j in enumerate(meta["iters"]):
        traj = task_traj(meta, j, task)
        lr = synthetic.linear_lr(3e-4, j, 1.0, meta["total_num_updates"])
        out = mopg_iteration_port(pol, opt, traj, j, lr, z[f"t{task}_weights"], z[f"t{task}_obj_var"],
                                  gamma=meta["gamma"], lam=meta["lam"], ppo_epoch=meta["E"],
                                  num_mini_batch=meta["B"])
        pre = f"t{task}_i{k}_"
        assert rel_err(out["returns"], z[pre + "returns"]) < 1e-12
        assert rel_err(out["adv"], z[pre + "adv"]) < 1e-12
        assert rel_err(out["losses"], z[pre + "losses"]) < 1e-9
        assert rel_err(pol.flat(), z[pre + "params"]) < 1e-9


def test_reference_runner_reproduces_golden():
    """oracle/ref_mopg_runner.py (bench.py's `kind: "reference"` CPU arm: the UNMODIFIED reference imported in place) gives
    the golden's first iteration. Runs only where /root/reference exists (the build container)."""
    import os
    if not os.path.isdir("/root/reference/morl"):
        pytest.skip("/root/reference is not present on this machine")
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import ref_mopg_runner as rr
from pgmorl_b200 import synthetic
from tests.helpers import load_mopg_case, task_traj, rel_err
z, meta = load_mopg_case("mopg_walker_small.npz")
d = meta["dims"]; task = 1; j = int(meta["iters"][0])
lr = synthetic.linear_lr(3e-4, j, 1.0, meta["total_num_updates"])
dt, flat, losses = rr._iteration(z[f"t{task}_init"], (d.obs, d.act, d.obj), task_traj(meta, j, task), j, lr,
                                 z[f"t{task}_weights"], z[f"t{task}_obj_var"], meta["gamma"], meta["lam"], meta["E"], meta["B"])
assert rel_err(flat, z[f"t{task}_i0_params"]) < 1e-12, rel_err(flat, z[f"t{task}_i0_params"])
assert rel_err(losses, z[f"t{task}_i0_losses"]) < 1e-12
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # separate process: importing the reference switches torch's default dtype to float64 for the whole interpreter
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
