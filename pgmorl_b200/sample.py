"""Sample / Task: a policy with its optimiser state and running-normalisation snapshot
(mirror of morl/sample.py:10-32 and morl/task.py:7-10).

A Sample carries everything needed to resume training under another weight: the policy, the PPO
agent including Adam's ``step / exp_avg / exp_avg_sq`` and the decayed learning rate, and the
ob/ret/obj running moments. ``link_policy_agent`` re-binds a fresh Adam to the (copied) policy
and reloads the optimiser state, exactly the reference's contract (sample.py:28-32).
"""
from copy import deepcopy


class Sample:
    def __init__(self, env_params, actor_critic, agent, objs=None, optgraph_id=None, owner=None):
        self.env_params = env_params
        self.actor_critic = actor_critic
        self.agent = agent
        self.link_policy_agent()
        self.objs = objs
        self.optgraph_id = optgraph_id
        # sharded runs (morl.run under torchrun, dist.py): rank whose HBM holds this sample's policy / Adam / moments.
        # On every other rank the sample is a stub (actor_critic / agent / env_params None) that carries only the
        # metadata the selection reads (objs, optgraph_id). None = single-process run.
        self.owner = owner

    @classmethod
    def copy_from(cls, sample):
        return cls(deepcopy(sample.env_params), deepcopy(sample.actor_critic), deepcopy(sample.agent),
                   deepcopy(sample.objs), sample.optgraph_id, getattr(sample, "owner", None))

    @classmethod
    def stub(cls, objs, optgraph_id=None, owner=None):
        """Metadata-only sample: the state lives on rank `owner`."""
        return cls(None, None, None, objs, optgraph_id, owner)

    @property
    def is_stub(self):
        return self.actor_critic is None

    def link_policy_agent(self):
        if self.agent is None:        # objective-only samples (selection tests / replay)
            return
        self.agent.actor_critic = self.actor_critic
        state = deepcopy(self.agent.optimizer.state_dict())
        self.agent.optimizer = self.agent.make_optimizer(self.actor_critic, lr=3e-4, eps=1e-5)
        self.agent.optimizer.load_state_dict(state)


class Task:
    """A MOPG task = (policy sample, scalarisation weight); both deep-copied (task.py:9-10)."""

    def __init__(self, sample, scalarization):
        self.sample = Sample.copy_from(sample)
        self.scalarization = deepcopy(scalarization)
