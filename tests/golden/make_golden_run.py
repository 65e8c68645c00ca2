"""Golden output of a WHOLE short PG-MORL run of the unmodified reference (`morl.run`, morl/morl.py:28-239: warm-up +
evolutionary generations with prediction-guided selection, forked MOPG workers) on replay environments, for the driver
loop of pgmorl_b200/morl.py. The reference's own result files (objs / optgraph / elites / weights / predictions /
offsprings per generation) are copied to tests/golden/run_2d/.

    python tests/golden/make_golden_run.py          (build container only: needs /root/reference)
"""
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.install()
sys.modules.setdefault("environments", types.ModuleType("environments"))     # morl.py:3 only registers gym envs

sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_envs  # noqa: E402
from pgmorl_b200.layout import NetDims  # noqa: E402


METHODS = ("prediction-guided", "moead", "ra", "pfa", "random")     # morl/morl.py:128-169


def run_args(save_dir, method="prediction-guided"):
    """The configuration of the golden run; tests/test_gpu_run.py builds the same namespace. The pseudo-method "long" is
    the longer prediction-guided cut (synth_envs.run_args_2d_long: warm-up 16 + 2 generations of 4, T = 512)."""
    if method == "long":
        return synth_envs.run_args_2d_long(save_dir)
    args = synth_envs.run_args_2d(save_dir)
    args.selection_method = method
    return args


def out_dir(method):
    return os.path.join(ROOT, "tests", "golden", "run_2d" if method == "prediction-guided" else "run_2d_" + method)


def one_run(method):
    import a2c_ppo_acktr.envs as envs_mod
    import gym
    d = NetDims(17, 6, 2)
    save_dir = tempfile.mkdtemp()
    args = run_args(save_dir, method)
    envs_mod.make_vec_envs = lambda **kw: synth_envs.SeededReplayVecEnv(d, args.num_steps, args.num_processes, [1.3, 0.7], base_seed=500)
    gym.make = lambda name: synth_envs.ToyEvalEnv(d)
    import morl  # the reference's driver
    t0 = time.time()
    morl.run(args)
    print("reference run: %.1f s" % (time.time() - t0))
    out = out_dir(method)
    shutil.rmtree(out, ignore_errors=True)
    for gen in sorted(os.listdir(save_dir)):
        if gen == "final":
            os.makedirs(os.path.join(out, gen), exist_ok=True)
            shutil.copy(os.path.join(save_dir, gen, "objs.txt"), os.path.join(out, gen, "objs.txt"))
            continue
        for sub in ("ep", "population", "elites"):
            os.makedirs(os.path.join(out, gen, sub), exist_ok=True)
            for f in os.listdir(os.path.join(save_dir, gen, sub)):
                shutil.copy(os.path.join(save_dir, gen, sub, f), os.path.join(out, gen, sub, f))
    print("golden written to", out, sorted(os.listdir(out)))


def main():
    for method in (sys.argv[1:] or METHODS):
        one_run(method)


if __name__ == "__main__":
    main()
