"""CPU baseline runner over the UNMODIFIED reference (test / bench infrastructure, never imported by pgmorl_b200/).

Where /root/reference is present (the build container; the GPU box has no copy), `bench.py --impl reference` times the
reference's own classes -- `a2c_ppo_acktr.model.Policy`, `storage.RolloutStorage`, `algo.PPO.update`, driven by the loop of
morl/mopg.py:96-144 -- imported in place through tests/golden/ref_import.py (gym and a2c_ppo_acktr.envs are
stubbed, nothing else). Same call signature as oracle/mopg_torch_port.timed_population_iteration, which is the fallback
(`kind: "port"`) where the reference cannot be imported."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))

_REF = {}


def _load():
    if _REF:
        return _REF
    import ref_import
    ref_import.install()
    import torch
    from a2c_ppo_acktr import algo, utils as a2c_utils
    from a2c_ppo_acktr.model import Policy
    from a2c_ppo_acktr.storage import RolloutStorage
    from scalarization_methods import WeightedSumScalarization
    _REF.update(torch=torch, algo=algo, utils=a2c_utils, Policy=Policy, RolloutStorage=RolloutStorage,
                Scal=WeightedSumScalarization, Box=ref_import.Box)
    return _REF


def _iteration(flat, dims, traj, j, lr, w, ov, gamma, lam, ppo_epoch, num_mini_batch):
    """One MOPG iteration of one task with the reference's own objects (morl/mopg.py:96-144 on replayed trajectories)."""
    R = _load()
    torch = R["torch"]
    O, A, M = dims
    T, N = traj["rewards"].shape[:2]
    policy = R["Policy"]((O,), R["Box"](A), base_kwargs={"layernorm": False}, obj_num=M)
    policy.double()
    off = 0
    with torch.no_grad():
        for p in policy.parameters():                      # named_parameters() order = the flat layout (pgmorl_b200/layout.py)
            n = p.numel()
            p.copy_(torch.as_tensor(flat[off:off + n], dtype=torch.float64).reshape(p.shape))
            off += n
    agent = R["algo"].PPO(policy, 0.2, ppo_epoch, num_mini_batch, 0.5, 0.0, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    scal = R["Scal"](num_objs=M, weights=np.asarray(w, dtype=np.float64))
    rollouts = R["RolloutStorage"](num_steps=T, num_processes=N, obs_shape=(O,), action_space=R["Box"](A),
                                   recurrent_hidden_state_size=policy.recurrent_hidden_state_size, obj_num=M)
    obs_all = torch.as_tensor(traj["obs"])
    rew, tm, tb = torch.as_tensor(traj["rewards"]), torch.as_tensor(traj["masks"]), torch.as_tensor(traj["bad_masks"])
    t0 = time.perf_counter()
    rollouts.obs[0].copy_(obs_all[0])
    torch.manual_seed(j)
    for g in agent.optimizer.param_groups:
        g["lr"] = lr
    for step in range(T):
        with torch.no_grad():
            value, action, action_log_prob, rhs = policy.act(rollouts.obs[step], rollouts.recurrent_hidden_states[step],
                                                             rollouts.masks[step])
        obj_tensor = torch.zeros([N, M])
        obj_tensor.copy_(rew[step])
        rollouts.insert(obs_all[step + 1], rhs, action, action_log_prob, value, obj_tensor,
                        torch.FloatTensor(tm[step + 1].unsqueeze(-1)), torch.FloatTensor(tb[step + 1].unsqueeze(-1)))
    with torch.no_grad():
        next_value = policy.get_value(rollouts.obs[-1], rollouts.recurrent_hidden_states[-1], rollouts.masks[-1]).detach()
    rollouts.compute_returns(next_value, True, gamma, lam, True)
    losses = agent.update(rollouts, scal, np.asarray(ov, dtype=np.float64))
    rollouts.after_update()
    dt = time.perf_counter() - t0
    out = torch.cat([p.detach().reshape(-1) for p in policy.parameters()]).numpy().copy()
    return dt, out, np.array(losses)


def _worker(args):
    flat, dims, traj, j, lr, w, ov, kw = args
    _load()["torch"].set_num_threads(1)
    return _iteration(flat, dims, traj, j, lr, w, ov, **kw)


_WARM = False


def timed_population_iteration(flats, dims, trajs, j, lr, weights, obj_var, processes, gamma=0.995, lam=0.95, ppo_epoch=10,
                               num_mini_batch=32):
    """Run one MOPG iteration for every task with a pool of `processes` forked workers (process per task like
    morl/morl.py:84-88, one torch thread each, morl/morl.py:34). Returns (wall_seconds, per-task seconds, final flats)."""
    import multiprocessing as mp
    global _WARM
    kw = dict(gamma=gamma, lam=lam, ppo_epoch=ppo_epoch, num_mini_batch=num_mini_batch)
    if not _WARM:       # pay the lazy imports once in the parent so the forked workers inherit them (the reference's
        O, A, M = dims  # workers are forked from a parent that already built every policy and optimizer)
        tiny = {"obs": np.zeros((3, 2, O), np.float32), "rewards": np.zeros((2, 2, M), np.float32),
                "masks": np.ones((3, 2), np.float32), "bad_masks": np.ones((3, 2), np.float32)}
        _iteration(np.asarray(flats[0]), dims, tiny, 0, 3e-4, np.ones(M) / M, np.ones(M), gamma, lam, 1, 1)
        _WARM = True
    jobs = [(np.asarray(flats[p]), dims, trajs[p], j, lr, weights[p], obj_var[p], kw) for p in range(len(flats))]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return wall, [r[0] for r in res], [r[1] for r in res]
