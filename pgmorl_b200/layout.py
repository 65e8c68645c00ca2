"""Flat parameter layout of one (policy, value) network of the population.

The order is the reference's ``Policy.named_parameters()`` order
(a2c_ppo_acktr/model.py:201-256 MLPBase/MOMLPBase with layernorm off,
distributions.py:71-79 DiagGaussian, utils.py:32-35 AddBias), so that a flat
vector converts to / from a reference ``state_dict`` by plain slicing:

    base.actor.0.{weight[H,O],bias[H]}, base.actor.2.{weight[H,H],bias[H]},
    base.critic.0.{...}, base.critic.2.{...},
    base.critic_linear.{weight[M,H],bias[M]},
    dist.fc_mean.{weight[A,H],bias[A]}, dist.logstd._bias[A,1]

The same offsets are compiled into the CUDA side (csrc/common.cuh,
``struct NetLayout``, exported as ``pgm_param_offsets``); tests/test_cabi.py checks
that both agree for every named environment shape.
"""
from collections import OrderedDict
from dataclasses import dataclass


@dataclass(frozen=True)
class NetDims:
    obs: int      # O
    act: int      # A
    obj: int      # M
    hidden: int = 64

    @property
    def n_par(self):
        return param_layout(self)[1]


def param_layout(d):
    """-> (OrderedDict name -> (offset, shape), n_par)."""
    O, A, M, H = d.obs, d.act, d.obj, d.hidden
    spec = [
        ("base.actor.0.weight", (H, O)), ("base.actor.0.bias", (H,)),
        ("base.actor.2.weight", (H, H)), ("base.actor.2.bias", (H,)),
        ("base.critic.0.weight", (H, O)), ("base.critic.0.bias", (H,)),
        ("base.critic.2.weight", (H, H)), ("base.critic.2.bias", (H,)),
        ("base.critic_linear.weight", (M, H)), ("base.critic_linear.bias", (M,)),
        ("dist.fc_mean.weight", (A, H)), ("dist.fc_mean.bias", (A,)),
        ("dist.logstd._bias", (A, 1)),
    ]
    out, off = OrderedDict(), 0
    for name, shape in spec:
        n = 1
        for s in shape:
            n *= s
        out[name] = (off, shape)
        off += n
    return out, off


# Named environment shapes from the reference's launch scripts.
ENV_SHAPES = {
    "walker2d": NetDims(17, 6, 2),
    "halfcheetah": NetDims(17, 6, 2),
    "hopper3": NetDims(11, 3, 3),
    "humanoid": NetDims(376, 17, 2),
}
