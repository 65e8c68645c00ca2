"""ORACLE / CPU BASELINE (test infrastructure, NOT product code) -- torch-CPU float64 port of
the reference's per-task MOPG iteration that keeps the reference's OP SEQUENCE: a per-step
``act`` loop of T tiny forwards, a T-step Python GAE loop, and a PPO loop of E*B minibatch
steps with autograd, ``clip_grad_norm_`` and ``torch.optim.Adam`` -- so that its run time on the
host cores is representative of the reference's CPU path (bench.py ``cpu_baseline`` and
``--impl reference``; the reference itself cannot travel to the GPU box).

Pinned by tests/test_oracle_mopg.py against the goldens made from the unmodified reference.

Reference anchors: morl/mopg.py:96-144; a2c_ppo_acktr/model.py:57-82,201-256;
distributions.py:30-40,71-90; storage.py:10-154; algo/ppo.py:40-115; utils.py:32-50.
"""
import math

import numpy as np
import torch
import torch.nn as nn

HALF_LOG_2PI = math.log(math.sqrt(2 * math.pi))


class PortPolicy(nn.Module):
    """Same parameter set / order as the reference Policy with MOMLPBase + DiagGaussian."""

    def __init__(self, flat, O, A, M, H=64):
        super().__init__()
        f64 = torch.float64
        self.a0, self.a2 = nn.Linear(O, H, dtype=f64), nn.Linear(H, H, dtype=f64)
        self.c0, self.c2 = nn.Linear(O, H, dtype=f64), nn.Linear(H, H, dtype=f64)
        self.cl = nn.Linear(H, M, dtype=f64)
        self.fm = nn.Linear(H, A, dtype=f64)
        self.logstd = nn.Parameter(torch.zeros(A, 1, dtype=f64))
        self.load_flat(flat)

    def ordered(self):
        return [self.a0.weight, self.a0.bias, self.a2.weight, self.a2.bias, self.c0.weight, self.c0.bias,
                self.c2.weight, self.c2.bias, self.cl.weight, self.cl.bias, self.fm.weight, self.fm.bias,
                self.logstd]

    def load_flat(self, flat):
        flat = torch.as_tensor(flat, dtype=torch.float64)
        off = 0
        with torch.no_grad():
            for p in self.ordered():
                n = p.numel()
                p.copy_(flat[off:off + n].view_as(p))
                off += n

    def flat(self):
        return torch.cat([p.detach().reshape(-1) for p in self.ordered()]).numpy().copy()

    def base(self, x):
        hc = torch.tanh(self.c2(torch.tanh(self.c0(x))))
        ha = torch.tanh(self.a2(torch.tanh(self.a0(x))))
        return self.cl(hc), ha

    def dist(self, ha):
        mean = self.fm(ha)
        logstd = torch.zeros(mean.size(), dtype=mean.dtype) + self.logstd.t().view(1, -1)
        return torch.distributions.Normal(mean, logstd.exp())

    def act(self, x):
        value, ha = self.base(x)
        dist = self.dist(ha)
        action = dist.sample()
        return value, action, dist.log_prob(action).sum(-1, keepdim=True)

    def evaluate_actions(self, x, action):
        value, ha = self.base(x)
        dist = self.dist(ha)
        return value, dist.log_prob(action).sum(-1, keepdim=True), dist.entropy().sum(-1).mean()


def mopg_iteration_port(policy, optimizer, traj, j, lr, weights, obj_var, gamma=0.995, lam=0.95,
                        ppo_epoch=10, num_mini_batch=32, clip=0.2, vcoef=0.5, ecoef=0.0, max_grad_norm=0.5):
    """One iteration of the loop body of morl/mopg.py:96-144 (env replaced by `traj`)."""
    obs_all = torch.as_tensor(traj["obs"])            # float32, as the reference's env wrapper yields
    T, N = obs_all.shape[0] - 1, obs_all.shape[1]
    M = traj["rewards"].shape[-1]
    A = policy.fm.out_features
    f64 = torch.float64
    obs = torch.zeros(T + 1, N, obs_all.shape[2], dtype=f64)
    rewards = torch.zeros(T, N, M, dtype=f64)
    value_preds = torch.zeros(T + 1, N, M, dtype=f64)
    returns = torch.zeros(T + 1, N, M, dtype=f64)
    logps = torch.zeros(T, N, 1, dtype=f64)
    actions = torch.zeros(T, N, A, dtype=f64)
    masks = torch.ones(T + 1, N, 1, dtype=f64)
    bad_masks = torch.ones(T + 1, N, 1, dtype=f64)
    obs[0].copy_(obs_all[0])

    torch.manual_seed(j)
    for g in optimizer.param_groups:
        g["lr"] = lr
    tm, tb, tr = torch.as_tensor(traj["masks"]), torch.as_tensor(traj["bad_masks"]), torch.as_tensor(traj["rewards"])
    for step in range(T):
        with torch.no_grad():
            value, action, logp = policy.act(obs[step])
        obs[step + 1].copy_(obs_all[step + 1])
        actions[step].copy_(action); logps[step].copy_(logp); value_preds[step].copy_(value)
        rewards[step].copy_(tr[step])
        masks[step + 1].copy_(tm[step + 1].unsqueeze(-1)); bad_masks[step + 1].copy_(tb[step + 1].unsqueeze(-1))
    with torch.no_grad():
        value_preds[-1] = policy.base(obs[-1])[0]
    gae = 0
    for step in reversed(range(T)):
        delta = rewards[step] + gamma * value_preds[step + 1] * masks[step + 1] - value_preds[step]
        gae = delta + gamma * lam * masks[step + 1] * gae
        gae = gae * bad_masks[step + 1]
        returns[step] = gae + value_preds[step]

    w = torch.as_tensor(weights, dtype=f64)
    sc = torch.as_tensor(np.sqrt(obj_var + 1e-8), dtype=f64)
    adv = ((returns * sc)[:-1] * w).sum(-1) - ((value_preds * sc)[:-1] * w).sum(-1)
    adv = (adv - adv.mean()) / (adv.std() + 1e-5)

    S = T * N
    mb = S // num_mini_batch
    sums = np.zeros(3)
    X = obs[:-1].view(S, -1); ACT = actions.view(S, -1); VO = value_preds[:-1].view(S, -1)
    RET = returns[:-1].view(S, -1); LP = logps.view(S, 1); ADV = adv.view(S, 1)
    params = list(policy.parameters())
    for _ in range(ppo_epoch):
        perm = torch.randperm(S)
        for b in range(num_mini_batch):
            idx = perm[b * mb:(b + 1) * mb]
            values, logp, ent = policy.evaluate_actions(X[idx], ACT[idx])
            ratio = torch.exp(logp - LP[idx])
            surr1 = ratio * ADV[idx]
            surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * ADV[idx]
            action_loss = -torch.min(surr1, surr2).mean()
            vclip = VO[idx] + (values - VO[idx]).clamp(-clip, clip)
            value_loss = 0.5 * torch.max((values - RET[idx]).pow(2), (vclip - RET[idx]).pow(2)).mean()
            optimizer.zero_grad()
            (value_loss * vcoef + action_loss - ent * ecoef).backward()
            nn.utils.clip_grad_norm_(params, max_grad_norm)
            optimizer.step()
            sums += (value_loss.item(), action_loss.item(), ent.item())
    return {"returns": returns[:-1].numpy(), "adv": adv.numpy(), "losses": sums / (ppo_epoch * num_mini_batch)}


def _worker(args):
    """Process-per-task like morl/morl.py:84-88, one torch thread each (morl/morl.py:34)."""
    import time
    flat, dims, traj, j, lr, w, ov, kw = args
    torch.set_num_threads(1)
    pol = PortPolicy(flat, *dims)
    opt = torch.optim.Adam(pol.ordered(), lr=3e-4, eps=1e-5)
    t0 = time.perf_counter()
    out = mopg_iteration_port(pol, opt, traj, j, lr, w, ov, **kw)
    dt = time.perf_counter() - t0
    return dt, pol.flat(), out["losses"]


_WARM = False


def _warm_parent(dims):
    """Pay torch's lazy imports (optimizer / distributions machinery, several seconds) once in the
    parent so forked workers inherit them -- as the reference's workers do, being forked from a
    parent that already built every policy and optimizer in warm-up (morl/morl.py:55,84-88)."""
    global _WARM
    if _WARM:
        return
    O, A, M = dims
    pol = PortPolicy(np.zeros(2 * (64 * O + 64 + 64 * 64 + 64) + M * 64 + M + A * 64 + 2 * A), O, A, M)
    opt = torch.optim.Adam(pol.ordered(), lr=3e-4, eps=1e-5)
    traj = {"obs": np.zeros((3, 2, O), np.float32), "rewards": np.zeros((2, 2, M), np.float32),
            "masks": np.ones((3, 2), np.float32), "bad_masks": np.ones((3, 2), np.float32)}
    mopg_iteration_port(pol, opt, traj, 0, 3e-4, np.ones(M) / M, np.ones(M), ppo_epoch=1, num_mini_batch=1)
    _WARM = True


def timed_population_iteration(flats, dims, trajs, j, lr, weights, obj_var, processes, **kw):
    """Run one MOPG iteration for every task with a pool of `processes` workers.
    Returns (wall_seconds, per-task seconds, final flats)."""
    import multiprocessing as mp
    import time
    _warm_parent(dims)
    jobs = [(flats[p], dims, trajs[p], j, lr, weights[p], obj_var[p], kw) for p in range(len(flats))]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return wall, [r[0] for r in res], [r[1] for r in res]
