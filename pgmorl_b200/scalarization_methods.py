"""Scalarisation functions (mirror of morl/scalarization_methods.py:5-29)."""
import torch


class ScalarizationFunction:
    def __init__(self, num_objs, weights=None):
        self.num_objs = num_objs
        self.weights = None if weights is None else torch.Tensor(weights)

    def update_weights(self, weights):
        if weights is not None:
            self.weights = torch.Tensor(weights)

    def evaluate(self, objs):
        raise NotImplementedError


class WeightedSumScalarization(ScalarizationFunction):
    def update_z(self, z):
        pass

    def evaluate(self, objs):
        """(objs * weights).sum(-1)  (scalarization_methods.py:28-29)."""
        return (objs * self.weights).sum(axis=-1)
