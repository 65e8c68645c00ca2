"""Import the UNMODIFIED reference (albo437/PGMORL at /root/reference) in-process.

Only used by make_golden.py in the build container; /root/reference does not exist
on the GPU box, so nothing under tests/ imports this at test time.

Recipe follows SURVEY.md section 8(c): stub `gym` and `a2c_ppo_acktr.envs`
(their real versions need gym/mujoco/baselines, absent here), then import the
reference's own modules from where they lie.
"""
import sys
import types

REF = "/root/reference"


def install():
    sys.dont_write_bytecode = True
    for p in (REF + "/morl", REF + "/externals/pytorch-a2c-ppo-acktr-gail"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "a2c_ppo_acktr.envs" not in sys.modules:
        envs = types.ModuleType("a2c_ppo_acktr.envs")

        class VecNormalize:  # only isinstance-checked by a2c_ppo_acktr.utils
            pass

        envs.VecNormalize = VecNormalize
        envs.make_env = None
        envs.make_vec_envs = None
        sys.modules["a2c_ppo_acktr.envs"] = envs
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.make = None
        sys.modules["gym"] = gym
    import torch
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)


class Box:
    """Stand-in for gym.spaces.Box: the reference only reads the class name and .shape."""

    def __init__(self, n):
        self.shape = (n,)
