"""External Pareto archive (mirror of morl/ep.py:10-31): all policies on the current front."""
from copy import deepcopy

import numpy as np

from .utils import get_ep_indices


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
        """Append deep copies of `sample_batch`, keep the non-dominated set sorted by objective 0
        (ep.py:23-31); the dominance filter runs on the GPU (K5 ep_filter)."""
        new = np.empty(len(sample_batch), dtype=object)
        for i, s in enumerate(sample_batch):
            new[i] = deepcopy(s)
        self.sample_batch = np.append(self.sample_batch, new)
        objs = [np.asarray(s.objs, dtype=np.float64) for s in sample_batch]
        if objs:
            self.obj_batch = np.vstack([self.obj_batch] + objs) if len(self.obj_batch) > 0 else np.vstack(objs)
        if len(self.obj_batch) == 0:
            return
        self.index(get_ep_indices(self.obj_batch))
