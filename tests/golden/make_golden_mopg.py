"""Generate tests/golden/mopg_*.npz by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_mopg.py

What runs is the reference's own code: ``Policy`` (a2c_ppo_acktr/model.py),
``RolloutStorage`` (storage.py), ``algo.PPO.update`` (algo/ppo.py),
``WeightedSumScalarization`` (morl/scalarization_methods.py) and
``update_linear_schedule`` (a2c_ppo_acktr/utils.py), driven by a transcription of
the loop body of morl/mopg.py:96-144 in which ``envs.step`` is replaced by reading
the next synthetic observation / reward / mask (MuJoCo is not installed and env
stepping is out of scope). Inputs come from pgmorl_b200.synthetic with fixed seeds.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import ref_import  # noqa: E402

ref_import.install()

from a2c_ppo_acktr import algo, utils as a2c_utils  # noqa: E402
from a2c_ppo_acktr.model import Policy  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402
from scalarization_methods import WeightedSumScalarization  # noqa: E402

from pgmorl_b200.layout import NetDims  # noqa: E402
from pgmorl_b200 import synthetic  # noqa: E402


def flat_of(policy):
    return torch.cat([p.detach().reshape(-1) for p in policy.parameters()]).numpy().copy()


def adam_flat(agent, key):
    st = agent.optimizer.state
    return torch.cat([st[p][key].reshape(-1) for p in agent.actor_critic.parameters()]).numpy().copy()


def reference_iterations(dims, T, N, E, B, n_tasks, iters, traj_seed, total_num_updates,
                         gamma=0.995, lam=0.95, lr0=3e-4, lr_decay_ratio=1.0):
    """Returns per-task outputs of the reference for iterations j in ``iters`` (sequential,
    state carried across iterations like a generation)."""
    O, A, M = dims.obs, dims.act, dims.obj
    weights_grid = synthetic.simplex_weights(M, 0.2 if M == 2 else 0.25)
    obj_var = np.array([1.3, 0.7, 0.9][:M])
    out = []
    for task in range(n_tasks):
        torch.manual_seed(1000 + task)
        policy = Policy((O,), ref_import.Box(A), base_kwargs={"layernorm": False}, obj_num=M)
        policy.double()
        init_flat = flat_of(policy)
        # the product's own init must reproduce the reference's (checked in tests)
        mine = synthetic.init_policy_flat(dims, seed=1000 + task).numpy()
        assert np.array_equal(mine, init_flat), "init_policy_flat diverges from reference Policy init"
        agent = algo.PPO(policy, 0.2, E, B, 0.5, 0.0, lr=lr0, eps=1e-5, max_grad_norm=0.5)
        w = weights_grid[task % len(weights_grid)]
        scal = WeightedSumScalarization(num_objs=M, weights=w)
        rollouts = RolloutStorage(num_steps=T, num_processes=N, obs_shape=(O,),
                                  action_space=ref_import.Box(A),
                                  recurrent_hidden_state_size=policy.recurrent_hidden_state_size,
                                  obj_num=M)
        per_iter = []
        for j in iters:
            traj = synthetic.make_trajectories(n_tasks, T, N, dims, seed=traj_seed + j)
            obs_all = traj["obs"][task]            # f32 [T+1,N,O]
            rollouts.obs[0].copy_(obs_all[0])
            rollouts.masks[0].copy_(traj["masks"][task, 0].unsqueeze(-1))
            rollouts.bad_masks[0].copy_(traj["bad_masks"][task, 0].unsqueeze(-1))
            # ---- morl/mopg.py:96-144 ----
            torch.manual_seed(j)
            a2c_utils.update_linear_schedule(agent.optimizer, j * lr_decay_ratio, total_num_updates, lr0)
            for step in range(T):
                with torch.no_grad():
                    value, action, action_log_prob, rhs = policy.act(
                        rollouts.obs[step], rollouts.recurrent_hidden_states[step], rollouts.masks[step])
                obs = obs_all[step + 1]                                   # envs.step(action)
                obj_tensor = torch.zeros([N, M])
                obj_tensor.copy_(traj["rewards"][task, step])
                masks = torch.FloatTensor(traj["masks"][task, step + 1].unsqueeze(-1))
                bad_masks = torch.FloatTensor(traj["bad_masks"][task, step + 1].unsqueeze(-1))
                rollouts.insert(obs, rhs, action, action_log_prob, value, obj_tensor, masks, bad_masks)
            with torch.no_grad():
                next_value = policy.get_value(rollouts.obs[-1], rollouts.recurrent_hidden_states[-1],
                                              rollouts.masks[-1]).detach()
            rollouts.compute_returns(next_value, True, gamma, lam, True)
            # snapshot what PPO.update will consume
            snap = {
                "value": rollouts.value_preds.numpy().copy(),
                "action": rollouts.actions.numpy().copy(),
                "logp": rollouts.action_log_probs.numpy().copy()[..., 0],
                "returns": rollouts.returns.numpy().copy()[:-1],
            }
            # advantage exactly as algo/ppo.py:43-56 computes it (recomputed here for the fixture)
            sc = torch.Tensor(np.sqrt(obj_var + 1e-8))
            adv = scal.evaluate((rollouts.returns * sc)[:-1]) - scal.evaluate((rollouts.value_preds * sc)[:-1])
            adv = (adv - adv.mean()) / (adv.std() + 1e-5)
            snap["adv"] = adv.numpy().copy()
            vl, al, ent = agent.update(rollouts, scal, obj_var)
            rollouts.after_update()
            snap.update({
                "losses": np.array([vl, al, ent]),
                "params": flat_of(policy),
                "adam_m": adam_flat(agent, "exp_avg"),
                "adam_v": adam_flat(agent, "exp_avg_sq"),
                "adam_step": float(agent.optimizer.state[next(policy.parameters())]["step"]),
                "lr": agent.optimizer.param_groups[0]["lr"],
            })
            per_iter.append(snap)
        out.append({"init": init_flat, "weights": w, "obj_var": obj_var, "iters": per_iter})
    return out


def save_case(name, dims, T, N, E, B, n_tasks, iters, traj_seed, total_num_updates, keep, **kw):
    res = reference_iterations(dims, T, N, E, B, n_tasks, iters, traj_seed, total_num_updates, **kw)
    blob = {"meta": np.array([dims.obs, dims.act, dims.obj, T, N, E, B, n_tasks, traj_seed,
                              total_num_updates], dtype=np.int64),
            "iters": np.array(list(iters), dtype=np.int64),
            "gamma_lam": np.array([kw.get("gamma", 0.995), kw.get("lam", 0.95)])}
    for p, r in enumerate(res):
        blob[f"t{p}_init"] = r["init"]
        blob[f"t{p}_weights"] = r["weights"]
        blob[f"t{p}_obj_var"] = r["obj_var"]
        for k, snap in enumerate(r["iters"]):
            for key in keep:
                blob[f"t{p}_i{k}_{key}"] = np.asarray(snap[key])
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **blob)
    print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    full = ["value", "action", "logp", "returns", "adv", "losses", "params", "adam_m", "adam_v",
            "adam_step", "lr"]
    light = ["losses", "params", "adam_m", "adam_v", "adam_step", "lr"]
    # small cases: every intermediate kept
    save_case("mopg_walker_small.npz", NetDims(17, 6, 2), T=64, N=4, E=3, B=4, n_tasks=3,
              iters=[0, 1], traj_seed=1, total_num_updates=610, keep=full)
    save_case("mopg_hopper3_small.npz", NetDims(11, 3, 3), T=48, N=2, E=2, B=3, n_tasks=2,
              iters=[5], traj_seed=7, total_num_updates=976, keep=full)
    save_case("mopg_humanoid_small.npz", NetDims(376, 17, 2), T=32, N=8, E=2, B=2, n_tasks=1,
              iters=[3], traj_seed=11, total_num_updates=1220, keep=full, gamma=0.99)
    # full-size C2 (HalfCheetah shape): 2 of the 6 tasks, one iteration, final state only
    save_case("mopg_halfcheetah_full.npz", NetDims(17, 6, 2), T=2048, N=4, E=10, B=32, n_tasks=2,
              iters=[0], traj_seed=1, total_num_updates=610, keep=light + ["returns", "adv"])
