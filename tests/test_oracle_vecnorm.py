"""oracle/vecnorm_oracle.py against the golden vectors of the unmodified reference VecNormalize / RunningMeanStd
(tests/golden/vecnorm.npz): every normalised observation, objective vector and running moment bit for bit."""
import os

import numpy as np
import pytest

from oracle.vecnorm_oracle import VecNormalizeOracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vecnorm.npz"))


@pytest.mark.parametrize("case", ["n4", "n9", "n1"])
def test_oracle_matches_reference_bit_for_bit(case):
    N, O, M, steps = (int(v) for v in G[case + "_dims"])
    env = VecNormalizeOracle(N, O, ob=True, ret=True, obj_rms=True, gamma=float(G[case + "_gamma"]))
    assert np.array_equal(env.reset(G[case + "_reset_obs"]).astype(np.float32), G[case + "_reset_out"])
    for t in range(steps):
        obs, _, obj = env.step(G[case + "_raw_obs"][t], G[case + "_raw_rew"][t], G[case + "_raw_obj"][t], G[case + "_raw_done"][t])
        assert np.array_equal(obs.astype(np.float32), G[case + "_obs_out"][t]), t
        assert np.array_equal(obj, G[case + "_obj_out"][t]), t
    for name, rms in (("ob", env.ob_rms), ("ret", env.ret_rms), ("obj", env.obj_rms)):
        assert np.array_equal(rms.mean, G[case + "_" + name + "_mean"]), name
        assert np.array_equal(rms.var, G[case + "_" + name + "_var"]), name
        assert rms.count == float(G[case + "_" + name + "_count"])
    assert np.array_equal(env.ret, G[case + "_ret_acc"]) and np.array_equal(env.obj, G[case + "_obj_acc"])
