"""The reference-facing API on the GPU: Policy / RolloutStorage / PPO mirrors, Sample / Task lifecycle, the
MOPG_worker contract and the population-batched update, checked against the goldens of the unmodified
reference (one MOPG iteration on replayed synthetic trajectories). Tolerance: FP32 device arithmetic vs the
float64 reference, norm-wise max|a-b| <= 1e-4 * max|b| per tensor (north_star)."""
import queue
import threading
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import synth_envs
from pgmorl_b200 import synthetic
from tests.helpers import load_mopg_case, rel_err

pytestmark = pytest.mark.gpu


def make_args(meta, j):
    d = meta["dims"]
    T, N = meta["T"], meta["N"]
    return SimpleNamespace(env_name="replay", seed=0, num_processes=N, gamma=meta["gamma"], obj_rms=True, ob_rms=True,
                           num_steps=T, obj_num=d.obj, num_env_steps=meta["total_num_updates"] * T * N,
                           use_linear_lr_decay=True, lr_decay_ratio=1.0, lr=3e-4, use_gae=True, gae_lambda=meta["lam"],
                           use_proper_time_limits=True, rl_log_interval=0, update_iter=1, eval_num=1, raw=True,
                           ppo_epoch=meta["E"], num_mini_batch=meta["B"])


def build_task(z, meta, task):
    from pgmorl_b200.a2c_ppo_acktr.algo import PPO
    from pgmorl_b200.a2c_ppo_acktr.model import Policy
    from pgmorl_b200.sample import Sample, Task
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    d = meta["dims"]
    torch.manual_seed(1000 + task)
    pol = Policy((d.obs,), synth_envs._Box(d.act), base_kwargs={"layernorm": False}, obj_num=d.obj)
    assert rel_err(pol.flat.cpu().numpy(), z[f"t{task}_init"]) < 1e-6        # same init as the reference's Policy
    agent = PPO(pol, 0.2, meta["E"], meta["B"], 0.5, 0.0, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    env_params = {"ob_rms": None, "ret_rms": None, "obj_rms": None}
    sample = Sample(env_params, pol, agent, objs=np.zeros(d.obj), optgraph_id=-1)
    return Task(sample, WeightedSumScalarization(num_objs=d.obj, weights=z[f"t{task}_weights"]))


def install_hooks(meta, z, j, tasks):
    from pgmorl_b200 import mopg
    d = meta["dims"]
    trajs = synthetic.make_trajectories(meta["n_tasks"], meta["T"], meta["N"], d, seed=meta["traj_seed"] + j)
    it = iter(tasks)

    def factory(**kw):
        task = next(it)
        return synth_envs.ReplayVecEnv({k: v[task].numpy() for k, v in trajs.items()}, d, z[f"t{task}_obj_var"])

    mopg.set_env_hooks(make_vec_envs=factory, gym_make=lambda name: synth_envs.ToyEvalEnv(d))


@pytest.mark.parametrize("name", ["mopg_walker_small.npz", "mopg_hopper3_small.npz"])
def test_mopg_worker_contract_and_parity(name):
    from pgmorl_b200 import mopg
    z, meta = load_mopg_case(name)
    j = meta["iters"][0]
    args = make_args(meta, j)
    task_id = 1
    task = build_task(z, meta, task_id)
    install_hooks(meta, z, j, [task_id])
    q, ev = queue.Queue(), threading.Event()
    ev.set()
    mopg.MOPG_worker(args, task_id, task, torch.device("cuda"), j, 1, 0.0, q, ev)
    msg = q.get_nowait()
    assert msg["task_id"] == task_id and msg["done"] is True and len(msg["offspring_batch"]) == 1
    s = msg["offspring_batch"][0]
    sd = s.actor_critic.state_dict()
    flat = np.concatenate([v.numpy().reshape(-1) for v in sd.values()])
    pre = f"t{task_id}_i0_"
    assert rel_err(flat, z[pre + "params"]) < 1e-4
    osd = s.agent.optimizer.state_dict()
    m = np.concatenate([osd["state"][i]["exp_avg"].numpy().reshape(-1) for i in range(13)])
    v = np.concatenate([osd["state"][i]["exp_avg_sq"].numpy().reshape(-1) for i in range(13)])
    assert rel_err(m, z[pre + "adam_m"]) < 1e-4 and rel_err(v, z[pre + "adam_v"]) < 1e-4
    assert float(osd["state"][0]["step"]) == float(z[pre + "adam_step"])
    assert abs(osd["param_groups"][0]["lr"] - float(z[pre + "lr"])) < 1e-15
    assert s.objs.shape == (meta["dims"].obj,) and np.isfinite(s.objs).all()
    # the worker trains the task's own copy in place and hands back deep-copied snapshots (mopg.py:146-155)
    assert s.actor_critic is not task.sample.actor_critic
    assert s.actor_critic.flat.data_ptr() != task.sample.actor_critic.flat.data_ptr()


def test_population_update_matches_per_task_goldens():
    from pgmorl_b200 import mopg
    z, meta = load_mopg_case("mopg_walker_small.npz")
    j = meta["iters"][0]
    args = make_args(meta, j)
    tasks = [build_task(z, meta, t) for t in range(meta["n_tasks"])]
    install_hooks(meta, z, j, list(range(meta["n_tasks"])))
    offspring = mopg.mopg_population_update(args, tasks, torch.device("cuda"), j, 1)
    for t in range(meta["n_tasks"]):
        s = offspring[t][0]
        flat = np.concatenate([v.numpy().reshape(-1) for v in s.actor_critic.state_dict().values()])
        assert rel_err(flat, z[f"t{t}_i0_params"]) < 1e-4
        assert s.agent.optimizer.step_count == int(z[f"t{t}_i0_adam_step"])


def test_sample_task_lifecycle_and_state_dict_roundtrip():
    from pgmorl_b200.a2c_ppo_acktr.model import Policy
    from pgmorl_b200.sample import Sample
    z, meta = load_mopg_case("mopg_walker_small.npz")
    task = build_task(z, meta, 0)
    s = task.sample
    c = Sample.copy_from(s)
    assert c.actor_critic is not s.actor_critic and c.actor_critic.flat.data_ptr() != s.actor_critic.flat.data_ptr()
    assert torch.equal(c.actor_critic.flat, s.actor_critic.flat) and c.agent.actor_critic is c.actor_critic
    sd = s.actor_critic.state_dict()
    assert list(sd)[0] == "base.actor.0.weight" and list(sd)[-1] == "dist.logstd._bias" and sd["dist.logstd._bias"].shape == (6, 1)
    d = meta["dims"]
    p2 = Policy((d.obs,), synth_envs._Box(d.act), obj_num=d.obj)
    p2.load_state_dict(sd)
    assert torch.equal(p2.flat, s.actor_critic.flat)
    # act / get_value / evaluate_actions are consistent with each other
    x = torch.randn(4, d.obs)
    torch.manual_seed(3)
    value, action, logp, _ = s.actor_critic.act(x, None, None)
    v2, logp2, ent, _ = s.actor_critic.evaluate_actions(x, None, None, action)
    assert torch.allclose(value, v2) and torch.allclose(logp, logp2, atol=1e-5) and value.shape == (4, d.obj)
    assert torch.allclose(s.actor_critic.get_value(x, None, None), value)


def test_pipelined_uploads_match_blocking_path():
    """upload_staged_async + swap_inputs (next iteration's H2D under this one's kernels) gives the same iteration as upload()."""
    from pgmorl_b200 import synthetic
    from pgmorl_b200.layout import NetDims
    from pgmorl_b200.population_state import PopulationMOPG
    d = NetDims(17, 6, 2)
    P, T, N, E, B = 2, 64, 4, 2, 4
    outs = []
    for pipelined in (False, True):
        pop = PopulationMOPG(d, P, T, N, ppo_epoch=E, num_mini_batch=B)
        for p in range(P):
            pop.load_task(p, synthetic.init_policy_flat(d, seed=7 + p).numpy(), weights=[0.4, 0.6], obj_var=[1.0, 2.0])
        pop.set_lr(3e-4)
        for j in range(3):
            traj = synthetic.make_trajectories(P, T, N, d, seed=100 + j)
            eps, perm = synthetic.host_rng_streams(j, T, N, d.act, E)
            src = dict(obs=traj["obs"], rewards=traj["rewards"], masks=traj["masks"], bad_masks=traj["bad_masks"],
                       eps=eps.to(torch.float32), perm=perm.to(torch.int32))
            if pipelined:
                for k, t in src.items():
                    pop._h[k].copy_(torch.as_tensor(t).reshape(pop._h[k].shape))
                pop.upload_staged_async()
                pop.swap_inputs()
            else:
                pop.upload(**src)
            pop.step()
            torch.cuda.synchronize()
        outs.append(pop.params.clone())
    assert torch.equal(outs[0], outs[1])


def test_batched_evaluation_equals_per_sample_evaluation():
    """evaluation_batch steps the evaluation environments of all offspring in lockstep (one K1 launch per step): same
    objective vectors, bit for bit, as evaluation() sample by sample -- with episodes of different lengths, observation
    normalisation, several evaluation episodes and discounting."""
    from pgmorl_b200 import mopg
    from pgmorl_b200.a2c_ppo_acktr.model import Policy
    z, meta = load_mopg_case("mopg_walker_small.npz")
    d = meta["dims"]

    class RaggedEnv(synth_envs.ToyEvalEnv):
        def step(self, action):
            ob, r, done, info = super().step(action)
            a = np.asarray(action, dtype=np.float64).reshape(-1)
            return ob, r, self.k >= 3 + int(abs(a[0]) * 50) % 9, info       # length depends on the policy's own actions

    mopg.set_env_hooks(gym_make=lambda name: RaggedEnv(d))
    rng = np.random.RandomState(0)
    samples = []
    for t in range(5):
        torch.manual_seed(50 + t)
        pol = Policy((d.obs,), synth_envs._Box(d.act), base_kwargs={"layernorm": False}, obj_num=d.obj)
        pol.flat.mul_(3.0)                                                     # spread the actions (and episode lengths)
        rms = synth_envs._Rms(rng.normal(0, 0.3, d.obs), rng.uniform(0.5, 2.0, d.obs))
        samples.append(SimpleNamespace(actor_critic=pol, env_params={"ob_rms": rms, "ret_rms": None, "obj_rms": None}))
    for raw, ob_rms in ((True, True), (False, True), (False, False)):
        args = make_args(meta, 0)
        args.eval_num, args.raw, args.ob_rms, args.seed = 3, raw, ob_rms, 11
        single = [mopg.evaluation(args, s) for s in samples]
        batch = mopg.evaluation_batch(args, samples)
        assert len(batch) == len(samples)
        for a, b in zip(single, batch):
            assert np.array_equal(a, b), (raw, ob_rms, a, b)
        assert len({tuple(np.round(a, 6)) for a in single}) > 1              # the samples really differ
    assert mopg.evaluation_batch(args, []) == []
