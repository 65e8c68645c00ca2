"""Golden vectors for K6 (running normalisation): the UNMODIFIED reference files
externals/baselines/baselines/common/vec_env/vec_normalize.py and .../common/running_mean_std.py are loaded from
where they lie (with a two-line stand-in for the VecEnvWrapper base class, whose real module drags in gym) and driven
with seeded raw simulator output. Recorded per step: normalised observations (float32, as a2c/envs.py:192 casts
them), normalised objective vectors, done flags; at the end: every running moment.

    python tests/golden/make_golden_vecnorm.py          (build container only: needs /root/reference)
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/externals/baselines/baselines/common"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_reference_vecnormalize():
    sys.dont_write_bytecode = True
    for name in ("baselines", "baselines.common"):
        sys.modules.setdefault(name, types.ModuleType(name))
    pkg = types.ModuleType("baselines.common.vec_env")
    pkg.__path__ = []

    class VecEnvWrapper:                       # vec_env.py:139-174 minus gym: holds the wrapped env, forwards attributes
        def __init__(self, venv, observation_space=None, action_space=None):
            self.venv, self.num_envs = venv, venv.num_envs
            self.observation_space = observation_space or venv.observation_space

    pkg.VecEnvWrapper = VecEnvWrapper
    sys.modules["baselines.common.vec_env"] = pkg
    for mod, path in (("baselines.common.running_mean_std", REF + "/running_mean_std.py"),
                      ("baselines.common.vec_env.vec_normalize", REF + "/vec_env/vec_normalize.py")):
        spec = importlib.util.spec_from_file_location(mod, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod] = m
        spec.loader.exec_module(m)
    return sys.modules["baselines.common.vec_env.vec_normalize"].VecNormalize


class RawEnv:
    """Seeded stand-in for the simulator side: raw observations, scalar rewards, objective vectors, dones."""

    def __init__(self, N, O, M, seed):
        self.num_envs, self.O, self.M = N, O, M
        self.observation_space = types.SimpleNamespace(shape=(O,))
        self.rng = np.random.RandomState(seed)
        self.scale = self.rng.uniform(0.1, 30.0, O)
        self.log = {"obs": [], "rew": [], "obj": [], "done": []}

    def _obs(self):
        return self.rng.standard_normal((self.num_envs, self.O)) * self.scale + 0.3 * self.scale

    def reset(self):
        o = self._obs()
        self.log["reset_obs"] = o.copy()
        return o

    def step_wait(self):
        N = self.num_envs
        obs, rew = self._obs(), self.rng.standard_normal(N) * 3.0
        obj = self.rng.uniform(-1.0, 5.0, (N, self.M)) * np.array([1.0, 40.0, 0.01][:self.M])
        done = self.rng.uniform(size=N) < 0.15
        for k, v in (("obs", obs), ("rew", rew), ("obj", obj), ("done", done)):
            self.log[k].append(v.copy())
        return obs, rew, done, [{"obj": obj[n].copy()} for n in range(N)]


def main():
    VecNormalize = load_reference_vecnormalize()
    out = {}
    for case, (N, O, M, steps, gamma) in {"n4": (4, 17, 2, 30, 0.995), "n9": (9, 11, 3, 20, 0.99), "n1": (1, 5, 2, 12, 0.99)}.items():
        raw = RawEnv(N, O, M, seed=7 + N)
        env = VecNormalize(raw, ob=True, ret=True, gamma=gamma, obj_rms=True)
        obs_n, obj_n = [], []
        out[case + "_reset_out"] = env.reset().astype(np.float32)
        for t in range(steps):
            obs, rews, news, infos = env.step_wait()
            obs_n.append(obs.astype(np.float32))
            obj_n.append(np.stack([i["obj"] for i in infos]))
            assert all((i["obj_raw"] == raw.log["obj"][-1][n]).all() for n, i in enumerate(infos))
        out[case + "_dims"] = np.array([N, O, M, steps])
        out[case + "_gamma"] = np.array(gamma)
        out[case + "_reset_obs"] = raw.log["reset_obs"]
        for k in ("obs", "rew", "obj", "done"):
            out[case + "_raw_" + k] = np.stack(raw.log[k])
        out[case + "_obs_out"], out[case + "_obj_out"] = np.stack(obs_n), np.stack(obj_n)
        for name, rms in (("ob", env.ob_rms), ("ret", env.ret_rms), ("obj", env.obj_rms)):
            out[case + "_" + name + "_mean"], out[case + "_" + name + "_var"] = np.asarray(rms.mean), np.asarray(rms.var)
            out[case + "_" + name + "_count"] = np.asarray(rms.count)
        out[case + "_ret_acc"], out[case + "_obj_acc"] = env.ret.copy(), np.stack(list(env.obj)).astype(np.float64)
    path = os.path.join(ROOT, "tests", "golden", "vecnorm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
