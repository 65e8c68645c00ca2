"""Device-backed mirror of the part of `a2c_ppo_acktr` (externals/pytorch-a2c-ppo-acktr-gail) that PG-MORL's
hot path uses: model.Policy, storage.RolloutStorage, algo.PPO, utils.update_linear_schedule."""
