"""External Pareto archive (mirror of morl/ep.py:10-31): all policies on the current front."""
from copy import copy, deepcopy

import numpy as np

from .utils import get_ep_indices


def _copy_sample(s):
    """Deep copy of a sample (ep.py:24 of the reference deep-copies every incoming sample). A metadata-only sample -- the stub
    of a policy whose state lives on another rank (sharded runs: all but 1/W of the offspring), or an objective-only
    stand-in -- holds nothing but its objective vector and ids, so a shallow copy plus a copy of the vector IS its deep
    copy, at a tenth of the cost of walking it with copy.deepcopy (1 000 offspring per generation at 48 tasks)."""
    if getattr(s, "actor_critic", 1) is None and getattr(s, "agent", 1) is None and getattr(s, "env_params", 1) is None:
        c = copy(s)
        if s.objs is not None:
            c.objs = np.array(s.objs, copy=True)
        return c
    return deepcopy(s)


class EP:
    def __init__(self):
        self.obj_batch = np.array([])
        self.sample_batch = np.array([])

    def index(self, indices, inplace=True):
        idx = np.array(indices, dtype=int)
        if inplace:
            self.obj_batch, self.sample_batch = self.obj_batch[idx], self.sample_batch[idx]
        else:
            return deepcopy(self.obj_batch[idx]), deepcopy(self.sample_batch[idx])

    def update(self, sample_batch):
        """Append `sample_batch`, keep the non-dominated set sorted by objective 0 (ep.py:23-31); the dominance filter
        runs on the GPU (K5 ep_filter). The reference deep-copies every incoming sample and then drops the dominated
        ones; here the filter runs first on the objective vectors and only the SURVIVING newcomers are copied -- the same
        archive (same members, same order, independent copies), without 100+ discarded policy copies per generation."""
        n_old = len(self.sample_batch)
        objs = [np.asarray(s.objs, dtype=np.float64) for s in sample_batch]
        if objs:
            self.obj_batch = np.vstack([self.obj_batch] + objs) if len(self.obj_batch) > 0 else np.vstack(objs)
        if len(self.obj_batch) == 0:
            return
        idx = np.array(get_ep_indices(self.obj_batch), dtype=int)
        kept = np.empty(len(idx), dtype=object)
        for j, i in enumerate(idx.tolist()):
            kept[j] = self.sample_batch[i] if i < n_old else _copy_sample(sample_batch[i - n_old])
        self.obj_batch, self.sample_batch = self.obj_batch[idx], kept
