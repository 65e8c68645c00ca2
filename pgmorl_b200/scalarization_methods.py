"""Scalarisation functions (mirror of morl/scalarization_methods.py:5-29). Weights are float64 tensors: the
reference runs under torch.set_default_dtype(torch.float64) (morl/morl.py:33), this package does not touch the
process-wide default."""
from copy import deepcopy

import numpy as np
import torch


def _f64(weights):
    if isinstance(weights, torch.Tensor):
        return weights.detach().to(torch.float64).clone()
    return torch.as_tensor(np.asarray(weights, dtype=np.float64))


class ScalarizationFunction:
    def __init__(self, num_objs, weights=None):
        self.num_objs = num_objs
        self.weights = None if weights is None else _f64(weights)

    def update_weights(self, weights):
        if weights is not None:
            self.weights = _f64(weights)

    def __deepcopy__(self, memo):
        """An independent copy, as copy.deepcopy gives, without torch's generic tensor deep copy (~50 us per object; the
        selection copies the template once per task and generation, morl/population_2d.py:296)."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = v.detach().clone() if isinstance(v, torch.Tensor) else deepcopy(v, memo)
        return new

    def evaluate(self, objs):
        raise NotImplementedError


class WeightedSumScalarization(ScalarizationFunction):
    def update_z(self, z):
        pass

    def evaluate(self, objs):
        """(objs * weights).sum(-1)  (scalarization_methods.py:28-29)."""
        return (objs * self.weights).sum(axis=-1)
