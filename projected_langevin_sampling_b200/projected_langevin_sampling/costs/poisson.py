"""Poisson cost (reference: src/projected_langevin_sampling/costs/poisson.py:10-104)."""
import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


class PoissonCost(PLSCost):
    """c = -2 y log|F| + link(F); with the square link d c/d F = -2 y / F + 2 F (poisson.py:63,76-82; no guard at F = 0)."""

    native_cost_id = nat.COST_POISSON
    closed_form_link = nat.LINK_SQUARE

    def __init__(self, y_train: torch.Tensor, link_function: PLSLinkFunction):
        super().__init__(link_function=link_function, observation_noise=None)
        self.y_train = y_train

    def predict(self, prediction_samples: torch.Tensor) -> torch.distributions.Poisson:
        return torch.distributions.Poisson(rate=prediction_samples.mean(dim=1))
