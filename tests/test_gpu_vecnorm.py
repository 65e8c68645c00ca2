"""K6 (pgm_vecnorm_step_f64) against the golden vectors of the unmodified reference VecNormalize /
RunningMeanStd (tests/golden/vecnorm.npz) and against the oracle on a multi-task shard: normalised observations
(float32), objective vectors, masks and every FP64 running moment BIT FOR BIT; then the fused per-step loop
(raw simulator output -> K6 -> K1 per-step inference) against the host-normalised loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vecnorm.npz"))


@pytest.mark.parametrize("case", ["n4", "n9", "n1"])
def test_matches_reference_bit_for_bit(case):
    from pgmorl_b200.vec_normalize import DeviceVecNormalize
    N, O, M, steps = (int(v) for v in G[case + "_dims"])
    P = 3                                  # the golden stream on task 1, other streams on its neighbours
    env = DeviceVecNormalize(P, N, O, M, ob=True, ret=True, obj_rms=True, gamma=float(G[case + "_gamma"]))
    dev = env.device
    obs_buf = torch.zeros(P, (steps + 1) * N, O, device=dev)           # rollout-buffer shaped: [P, (T+1) N, O]
    rew_buf, mask_buf = torch.zeros(P, steps, N, M, device=dev), torch.zeros(P, steps + 1, N, device=dev)
    rng = np.random.RandomState(3)
    rep = lambda x: np.stack([x * 1.7 + 0.1, x, -x])
    env.reset(rep(G[case + "_reset_obs"]), obs_buf[:, 0:N])
    assert np.array_equal(obs_buf[1, 0:N].cpu().numpy(), G[case + "_reset_out"])
    for t in range(steps):
        done = G[case + "_raw_done"][t]
        env.step(rep(G[case + "_raw_obs"][t]), rep(G[case + "_raw_rew"][t]), rep(G[case + "_raw_obj"][t]),
                 np.stack([rng.uniform(size=N) < 0.5, done, ~done]),
                 obs_buf[:, (t + 1) * N:(t + 2) * N], rew_buf[:, t], mask_buf[:, t + 1])
    torch.cuda.synchronize()
    got_obs = obs_buf[1, N:].reshape(steps, N, O).cpu().numpy()
    assert np.array_equal(got_obs, G[case + "_obs_out"])
    assert np.array_equal(rew_buf[1].cpu().numpy(), G[case + "_obj_out"].astype(np.float32))
    assert np.array_equal(mask_buf[1, 1:].cpu().numpy(), 1.0 - G[case + "_raw_done"].astype(np.float32))
    snap = env.snapshot(1)
    for name, key in (("ob", "ob_rms"), ("ret", "ret_rms"), ("obj", "obj_rms")):
        assert np.array_equal(snap[key].mean, G[case + "_" + name + "_mean"]), name
        assert np.array_equal(snap[key].var, G[case + "_" + name + "_var"]), name
        assert snap[key].count == float(G[case + "_" + name + "_count"])
    assert np.array_equal(env.ret_acc[1].cpu().numpy(), G[case + "_ret_acc"])
    assert np.array_equal(env.obj_acc[1].cpu().numpy(), G[case + "_obj_acc"])


def test_shard_against_oracle_with_frozen_moments_and_snapshots():
    """every task of a shard vs its own oracle instance, incl. eval() (frozen observation moments), disabled
    objective normalisation, and a snapshot -> load_task round trip into another slot"""
    from oracle.vecnorm_oracle import VecNormalizeOracle
    from pgmorl_b200.vec_normalize import DeviceVecNormalize
    P, N, O, M, steps = 5, 8, 376, 2, 12
    rng = np.random.RandomState(11)
    for obj_rms in (True, False):
        env = DeviceVecNormalize(P, N, O, M, ob=True, ret=True, obj_rms=obj_rms, gamma=0.99)
        orc = [VecNormalizeOracle(N, O, ob=True, ret=True, obj_rms=obj_rms, gamma=0.99) for _ in range(P)]
        obs_out, obj_out, mask = (torch.zeros(P, N, O, device=env.device), torch.zeros(P, N, M, device=env.device),
                                  torch.zeros(P, N, device=env.device))
        for t in range(steps):
            if t == 8:
                env.eval()
            raw_obs, raw_rew = rng.standard_normal((P, N, O)) * 50 + 3, rng.standard_normal((P, N))
            raw_obj, done = rng.uniform(0, 9, (P, N, M)), rng.uniform(size=(P, N)) < 0.2
            env.step(raw_obs, raw_rew, raw_obj, done, obs_out, obj_out, mask)
            for p in range(P):
                o, _, j = orc[p].step(raw_obs[p], raw_rew[p], raw_obj[p], done[p], update=t < 8)
                assert np.array_equal(obs_out[p].cpu().numpy(), o.astype(np.float32)), (t, p)
                assert np.array_equal(obj_out[p].cpu().numpy(), j.astype(np.float32)), (t, p)
        snap = env.snapshot(2)
        assert np.array_equal(snap['ob_rms'].var, orc[2].ob_rms.var) and snap['ob_rms'].count == orc[2].ob_rms.count
        assert (snap['obj_rms'] is None) == (not obj_rms)
        env.load_task(4, snap)
        assert torch.equal(env.ob_mean[4], env.ob_mean[2]) and float(env.ob_count[4]) == float(env.ob_count[2])


class _RawEnv:
    """Seeded raw simulator stand-in (independent of the actions): iteration j is read from torch's global seed, which
    both loops set with torch.manual_seed(j) before stepping (mopg.py:96)."""

    def __init__(self, N, O, M):
        self.N, self.O, self.M, self.t, self.j = N, O, M, 0, None
        self.scale = np.linspace(0.2, 25.0, O)

    def _rng(self, j, t):
        return np.random.RandomState(100000 + 1000 * j + t)

    def reset(self):
        self.t, self.j = 0, None
        return self._rng(-1, 0).standard_normal((self.N, self.O)) * self.scale + 1.0

    def step(self, action):
        j = int(torch.initial_seed())
        if j != self.j:
            self.j, self.t = j, 0
        r = self._rng(j, self.t)
        self.t += 1
        obs = r.standard_normal((self.N, self.O)) * self.scale + 1.0
        rew, obj = r.standard_normal(self.N), r.uniform(0.0, 6.0, (self.N, self.M)) * np.array([1.0, 30.0][:self.M])
        done = r.uniform(size=self.N) < 0.05
        infos = [dict(obj=obj[n].copy(), **({"bad_transition": True} if done[n] and r.uniform() < 0.5 else {})) for n in range(self.N)]
        return obs, rew, done, infos

    def close(self):
        pass


class _HostNormalised:
    """The reference arrangement: VecNormalize on the host around the raw environments (oracle restatement)."""

    def __init__(self, raw, dims, gamma):
        from oracle.vecnorm_oracle import VecNormalizeOracle
        from synth_envs import _Box
        self.raw, self.vn = raw, VecNormalizeOracle(raw.N, raw.O, ob=True, ret=True, obj_rms=True, gamma=gamma)
        self.observation_space, self.action_space, self.venv = _Box(dims.obs), _Box(dims.act), self

    ob_rms = property(lambda s: s.vn.ob_rms, lambda s, v: setattr(s.vn, "ob_rms", v))
    ret_rms = property(lambda s: s.vn.ret_rms, lambda s, v: setattr(s.vn, "ret_rms", v))
    obj_rms = property(lambda s: s.vn.obj_rms, lambda s, v: setattr(s.vn, "obj_rms", v))

    def reset(self):
        return torch.as_tensor(self.vn.reset(self.raw.reset()).astype(np.float32))

    def step(self, action):
        obs, rew, done, infos = self.raw.step(action)
        o, r, j = self.vn.step(obs, rew, np.stack([i["obj"] for i in infos]), done)
        for n, i in enumerate(infos):
            i["obj_raw"], i["obj"] = i["obj"], j[n]
        return torch.as_tensor(o.astype(np.float32)), r, done, infos

    def close(self):
        pass


def test_fused_rollout_loop_equals_host_normalised_loop(tmp_path):
    """mopg_population_update with the normalisation on the device (raw simulator output -> K6 -> K1 per step) produces
    the SAME offspring -- parameters, Adam state and running moments bit for bit -- as the loop that normalises on the
    host the way the reference does"""
    import synth_envs
    from pgmorl_b200 import mopg, synthetic, warm_up
    from pgmorl_b200.layout import NetDims
    from pgmorl_b200.sample import Task
    d = NetDims(17, 6, 2)
    args = synth_envs.run_args_2d(str(tmp_path), T=32, N=4)
    mopg.set_env_hooks(make_vec_envs=lambda **kw: _HostNormalised(_RawEnv(args.num_processes, d.obs, d.obj), d, args.gamma),
                       gym_make=lambda name: synth_envs.ToyEvalEnv(d), make_raw_vec_envs=False)
    torch.manual_seed(0)
    elites, scals = warm_up.initialize_warm_up_batch(args, "cuda")
    tasks = [Task(e, s) for e, s in zip(elites[:3], scals[:3])]
    host = mopg.mopg_population_update(args, tasks, torch.device("cuda"), 0, 3)
    mopg.set_env_hooks(make_raw_vec_envs=lambda **kw: _RawEnv(args.num_processes, d.obs, d.obj))
    try:
        tasks = [Task(e, s) for e, s in zip(elites[:3], scals[:3])]
        fused = mopg.mopg_population_update(args, tasks, torch.device("cuda"), 0, 3)
    finally:
        mopg.set_env_hooks(make_raw_vec_envs=False)
    for hs, fs in zip(host, fused):
        assert len(hs) == len(fs) == 3
        for a, b in zip(hs, fs):
            assert torch.equal(a.actor_critic.flat, b.actor_critic.flat)
            assert torch.equal(a.agent.optimizer.exp_avg_sq, b.agent.optimizer.exp_avg_sq)
            assert np.array_equal(a.objs, b.objs)
            for key in ("ob_rms", "ret_rms", "obj_rms"):
                assert np.array_equal(a.env_params[key].mean, b.env_params[key].mean), key
                assert np.array_equal(a.env_params[key].var, b.env_params[key].var), key
                assert a.env_params[key].count == b.env_params[key].count
