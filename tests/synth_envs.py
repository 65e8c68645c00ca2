"""Test infrastructure (not part of the product package): objective-only samples, synthetic optimisation histories and
full-size selection states for the selection path, and the replay / toy environments that
stand in for MuJoCo (out of scope) in the API and driver-loop tests. Used by tests/, tests/golden/make_golden_*.py,
bench.py's selection / generation legs and profiles/*.py."""
import numpy as np
import torch

from pgmorl_b200.synthetic import make_trajectories, simplex_weights

# ---------------------------------------------------------------------------------------------
# Synthetic optimisation histories for the selection path
# ---------------------------------------------------------------------------------------------
class ObjSample:
    """Objective-only stand-in for a Sample: selection reads nothing but .objs / .optgraph_id."""

    def __init__(self, objs, optgraph_id=None):
        self.env_params = None
        self.actor_critic = None
        self.agent = None
        self.objs = objs
        self.optgraph_id = optgraph_id


class SelectionArgs:
    """The argparse fields selection reads, with the reference's launch-script values
    (scripts/walker2d-v2.py:38-49 for 2 objectives, scripts/hopper-v3.py:36-48 for 3)."""

    def __init__(self, obj_num, **kw):
        self.obj_num = obj_num
        self.min_weight, self.max_weight = 0.0, 1.0
        self.num_processes = 4
        self.pbuffer_size = 2
        self.num_weight_candidates = 7
        if obj_num == 2:
            self.num_tasks, self.delta_weight, self.pbuffer_num, self.sparsity = 6, 0.2, 100, 1.0
        else:
            self.num_tasks, self.delta_weight, self.pbuffer_num, self.sparsity = 15, 0.25, 20, 1e6
        self.__dict__.update(kw)


def _response(rng, objs, w):
    """Toy training response: gain grows with the weight on that objective and saturates."""
    gain = 3.0 * (0.1 + 1.5 * w) * np.exp(-np.linalg.norm(objs) / 150.0)   # always positive: objectives only grow
    return gain + rng.normal(0.0, 0.1, size=objs.shape)


def run_selection_history(classes, args, generations, seed, update_iter=3, warmup_gens=3, on_generation=None):
    """Drive EP / Population / OptGraph classes (the product's or the reference's) through a
    synthetic run shaped like morl/morl.py:55-177; `on_generation(g, state)` sees every call's
    inputs and outputs. classes: dict(EP, Population, OptGraph, Scalarization)."""
    import torch
    rng = np.random.RandomState(seed)
    M = args.obj_num
    ep, population, graph = classes["EP"](), classes["Population"](args), classes["OptGraph"]()
    template = classes["Scalarization"](num_objs=M, weights=np.ones(M) / M)
    elite_batch, scal_batch = [], []
    for w in simplex_weights(M, args.delta_weight):
        s = ObjSample(rng.uniform(5.0, 10.0, M))
        sc = classes["Scalarization"](num_objs=M, weights=w)
        s.optgraph_id = graph.insert(sc.weights.clone(), s.objs.copy(), -1)   # roots carry torch weights
        elite_batch.append(s); scal_batch.append(sc)
    for g in range(generations):
        iters = update_iter * (warmup_gens if g == 0 else 1)
        all_samples, offspring = [], []
        for elite, sc in zip(elite_batch, scal_batch):
            w = sc.weights.detach().numpy().astype(np.float64)
            prev, objs = elite.optgraph_id, np.array(elite.objs, dtype=np.float64)
            for it in range(iters):
                objs = objs + _response(rng, objs, w)
                s = ObjSample(objs.copy())
                all_samples.append(s)
                if (it + 1) % update_iter == 0:
                    prev = graph.insert(w.copy(), objs.copy(), prev)
                    s.optgraph_id = prev
                    offspring.append(s)
        ep.update(all_samples)
        population.update(offspring)
        np.random.seed(1000 + g)          # the 3-objective candidate order consumes numpy's global RNG
        elite_batch, scal_batch, predicted = population.prediction_guided_selection(args, g, ep, graph, template)
        if on_generation is not None:
            on_generation(g, dict(ep=ep, population=population, graph=graph, elites=elite_batch,
                                  scalarizations=scal_batch, predicted=predicted))
        if len(elite_batch) == 0:
            break
    return ep, population, graph


def make_selection_state(obj_num, n_pop, n_ep, seed=0, siblings=6, **arg_overrides):
    """A full-size selection problem built directly (2 objectives at n_pop 200 / 1 400
    candidates / archive 300; 3 objectives at n_pop 420 / 2 940 candidates / archive 500), without running the
    hundreds of generations a real history would need to fill every performance buffer.

    Opt-graph: families of one root and `siblings` children trained with distinct simplex-grid weights (so every
    neighbourhood holds more than 3 distinct weights, population_2d.py:39-50); the population is made of children,
    `n_pop` of them. Archive: `n_ep` mutually non-dominated points on a sphere that cuts through the population's
    objective range, so some candidates extend the front and others are dominated.
    Returns (args, opt_graph, population, ep) of the product's classes."""
    from pgmorl_b200 import population_2d, population_3d
    from pgmorl_b200.ep import EP
    from pgmorl_b200.opt_graph import OptGraph
    rng = np.random.RandomState(seed)
    M = obj_num
    args = SelectionArgs(M, **arg_overrides)
    grid = [w for w in simplex_weights(M, 0.125 if M == 3 else 0.1) if np.min(w) > 0]
    graph = OptGraph()
    members = []
    while len(members) < n_pop:
        direction = rng.dirichlet(np.ones(M) * 4.0)
        root_objs = direction / np.linalg.norm(direction) * rng.uniform(60.0, 100.0)
        root = graph.insert(np.ones(M) / M, root_objs.copy(), -1)
        for j in rng.choice(len(grid), size=siblings, replace=False):
            w = np.asarray(grid[j], dtype=np.float64)
            child_objs = root_objs + _response(rng, root_objs, w)
            members.append(ObjSample(child_objs.copy(), graph.insert(w.copy(), child_objs.copy(), root)))
    members = members[:n_pop]
    pop = (population_2d if M == 2 else population_3d).Population(args)
    pop.sample_batch = members
    # points of equal norm in the positive orthant are mutually non-dominated (a >= b with a != b implies |a| > |b|)
    v = np.abs(rng.normal(size=(n_ep, M))) + 0.05
    shell = v / np.linalg.norm(v, axis=1, keepdims=True) * 80.0
    shell = shell[np.argsort(shell[:, 0], kind="stable")]                 # the archive is kept in objective-0 order
    ep = EP()
    ep.obj_batch = np.array(shell)
    ep.sample_batch = np.array([ObjSample(o.copy()) for o in ep.obj_batch], dtype=object)
    return args, graph, pop, ep


# ---------------------------------------------------------------------------------------------
# Replay environments: the VecEnv / gym surface MOPG_worker touches, fed from synthetic trajectories
# ---------------------------------------------------------------------------------------------
class _Box:
    def __init__(self, n):
        self.shape = (n,)


_Box.__name__ = "Box"        # the reference dispatches on the class NAME of the action space


class _Rms:
    def __init__(self, mean, var):
        self.mean, self.var = np.asarray(mean, dtype=np.float64), np.asarray(var, dtype=np.float64)
        self.count = 1e-4


class ReplayVecEnv:
    """Observations / objective vectors / termination flags replayed from make_trajectories() output of one task
    (independent of the actions): the stand-in for make_vec_envs() where MuJoCo is not available."""

    def __init__(self, traj_task, dims, obj_var):
        self.traj, self.t = traj_task, 0
        self.observation_space = _Box(dims.obs)
        self.action_space = _Box(dims.act)
        self.ob_rms = _Rms(np.zeros(dims.obs), np.ones(dims.obs))
        self.ret_rms = None
        self.obj_rms = _Rms(np.zeros(dims.obj), obj_var)
        self.venv = self

    def reset(self):
        self.t = 0
        return torch.as_tensor(self.traj["obs"][0])

    def step(self, action):
        t = self.t
        self.t += 1
        obs = torch.as_tensor(self.traj["obs"][t + 1])
        done = np.asarray(self.traj["masks"][t + 1]) == 0
        infos = []
        for n in range(len(done)):
            info = {"obj": np.asarray(self.traj["rewards"][t, n], dtype=np.float64),
                    "obj_raw": np.asarray(self.traj["rewards"][t, n], dtype=np.float64)}
            if self.traj["bad_masks"][t + 1, n] == 0:
                info["bad_transition"] = True
            infos.append(info)
        return obs, None, done, infos

    def close(self):
        pass


class RawReplayVecEnv:
    """Raw-simulator stand-in (no VecNormalize wrapper): un-normalised float64 observations, scalar rewards,
    raw objective vectors and termination flags replayed from make_trajectories() output of one task, scaled away from
    zero mean / unit variance so the running moments have something to do. Independent of the actions."""

    def __init__(self, traj_task, dims):
        self.traj, self.t, self.dims = traj_task, 0, dims
        self.obs = np.asarray(traj_task["obs"], dtype=np.float64) * 3.0 + 1.0
        self.obj = np.asarray(traj_task["rewards"], dtype=np.float64) * np.array([1.0, 30.0, 5.0][:dims.obj])
        self.done = np.asarray(traj_task["masks"]) == 0
        self.bad = np.asarray(traj_task["bad_masks"]) == 0
        self.N = self.obs.shape[1]

    def reset(self):
        self.t = 0
        return self.obs[0]

    def step(self, action):
        t = self.t
        self.t += 1
        infos = [dict(obj=self.obj[t, n], **({"bad_transition": True} if self.bad[t + 1, n] else {})) for n in range(self.N)]
        return self.obs[t + 1], self.obj[t].sum(axis=1), self.done[t + 1], infos

    def close(self):
        pass


class SeededReplayVecEnv:
    """Replay environment for whole runs (driver-loop tests): the trajectory of MOPG iteration j is
    make_trajectories(seed = base_seed + j), where j is read from torch's global seed -- both the reference's worker
    (mopg.py:96) and the batched update call torch.manual_seed(j) before stepping iteration j. reset() returns a fixed
    observation because it is called before the first manual_seed."""

    def __init__(self, dims, T, N, obj_var, base_seed=0):
        self.dims, self.T, self.N, self.base_seed = dims, T, N, base_seed
        self.observation_space, self.action_space = _Box(dims.obs), _Box(dims.act)
        self.ob_rms = _Rms(np.zeros(dims.obs), np.ones(dims.obs))
        self.ret_rms = None
        self.obj_rms = _Rms(np.zeros(dims.obj), obj_var)
        self.venv = self
        self.j, self.t, self.traj = None, 0, None

    def reset(self):
        self.j, self.t = None, 0
        return torch.full((self.N, self.dims.obs), 0.25, dtype=torch.float32)

    def step(self, action):
        j = int(torch.initial_seed())
        if j != self.j:
            self.j, self.t = j, 0
            self.traj = {k: v[0].numpy() for k, v in make_trajectories(1, self.T, self.N, self.dims, seed=self.base_seed + j).items()}
        t = self.t
        self.t += 1
        obs = torch.as_tensor(self.traj["obs"][t + 1])
        done = np.asarray(self.traj["masks"][t + 1]) == 0
        infos = []
        for n in range(len(done)):
            info = {"obj": np.asarray(self.traj["rewards"][t, n], dtype=np.float64),
                    "obj_raw": np.asarray(self.traj["rewards"][t, n], dtype=np.float64)}
            if self.traj["bad_masks"][t + 1, n] == 0:
                info["bad_transition"] = True
            infos.append(info)
        return obs, None, done, infos

    def close(self):
        pass


class ToyEvalEnv:
    """Deterministic gym-like evaluation env: objective = smooth function of the policy's mean action."""

    def __init__(self, dims, horizon=5):
        self.dims, self.horizon = dims, horizon
        self.observation_space, self.action_space = _Box(dims.obs), _Box(dims.act)

    def seed(self, s):
        self.rng = np.random.RandomState(s)

    def reset(self):
        self.k = 0
        return self.rng.uniform(-1, 1, self.dims.obs)

    def step(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        self.k += 1
        obj = np.array([1.0 + np.tanh(a[:self.dims.act // 2].sum()), 1.0 + np.tanh(a[self.dims.act // 2:].sum()), 1.0][:self.dims.obj])
        return self.rng.uniform(-1, 1, self.dims.obs), 0.0, self.k >= self.horizon, {"obj": obj}

    def close(self):
        pass


def run_args_2d(save_dir, T=64, N=4):
    """Namespace of a short 2-objective PG-MORL run at Walker2d dims (driver-loop golden): 6 warm-up tasks
    (delta_weight 0.2), warm-up of 2 iterations, 2 generations of 2 iterations, prediction-guided selection."""
    from types import SimpleNamespace
    return SimpleNamespace(
        env_name="replay", seed=0, obj_num=2, num_steps=T, num_processes=N, num_env_steps=6 * T * N,
        warmup_iter=2, update_iter=2, selection_method="prediction-guided", min_weight=0.0, max_weight=1.0,
        delta_weight=0.2, save_dir=save_dir, layernorm=False, algo="ppo", clip_param=0.2, ppo_epoch=2, num_mini_batch=4,
        value_loss_coef=0.5, entropy_coef=0.0, lr=3e-4, max_grad_norm=0.5, gamma=0.995, obj_rms=True, ob_rms=True,
        eval_num=1, raw=True, use_linear_lr_decay=True, lr_decay_ratio=1.0, use_gae=True, gae_lambda=0.95,
        use_proper_time_limits=True, rl_log_interval=0, pbuffer_num=100, pbuffer_size=2, num_tasks=6,
        num_weight_candidates=7, sparsity=1.0)


def run_args_2d_long(save_dir):
    """A longer cut of BASELINE.json configs[0] (C1): 6 tasks, T = 512 x 4 envs, minibatches of 256 rows, warm-up of 16
    iterations + 2 generations of 4 iterations (24 MOPG iterations = 768 Adam steps per task), prediction-guided selection."""
    args = run_args_2d(save_dir, T=512, N=4)
    args.warmup_iter, args.update_iter = 16, 4
    args.num_env_steps = (16 + 2 * 4) * 512 * 4
    args.ppo_epoch, args.num_mini_batch = 4, 8
    return args
