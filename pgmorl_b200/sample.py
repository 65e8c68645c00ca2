"""Sample / Task: a policy with its optimiser state and running-normalisation snapshot
(mirror of morl/sample.py:10-32 and morl/task.py:7-10).

A Sample carries everything needed to resume training under another weight: the policy, the PPO
agent including Adam's ``step / exp_avg / exp_avg_sq`` and the decayed learning rate, and the
ob/ret/obj running moments. ``link_policy_agent`` re-binds a fresh Adam to the (copied) policy
and reloads the optimiser state, exactly the reference's contract (sample.py:28-32).
"""
from copy import deepcopy


class Sample:
    def __init__(self, env_params, actor_critic, agent, objs=None, optgraph_id=None):
        self.env_params = env_params
        self.actor_critic = actor_critic
        self.agent = agent
        self.link_policy_agent()
        self.objs = objs
        self.optgraph_id = optgraph_id

    @classmethod
    def copy_from(cls, sample):
        return cls(deepcopy(sample.env_params), deepcopy(sample.actor_critic), deepcopy(sample.agent),
                   deepcopy(sample.objs), sample.optgraph_id)

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
