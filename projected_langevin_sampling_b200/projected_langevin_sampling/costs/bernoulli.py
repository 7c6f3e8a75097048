"""Bernoulli cost (reference: src/projected_langevin_sampling/costs/bernoulli.py:10-99)."""
import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


class BernoulliCost(PLSCost):
    """Cross entropy; with the sigmoid link d c/d F = -y (1 - p) + (1 - y) p on the CLIPPED p (bernoulli.py:72-77)."""

    native_cost_id = nat.COST_BERNOULLI
    closed_form_link = nat.LINK_SIGMOID

    def __init__(self, y_train: torch.Tensor, link_function: PLSLinkFunction):
        super().__init__(link_function=link_function, observation_noise=None)
        self.y_train = y_train.type(torch.double)  # bernoulli.py:32

    def predict(self, prediction_samples: torch.Tensor) -> torch.distributions.Bernoulli:
        return torch.distributions.Bernoulli(probs=prediction_samples.mean(dim=1))
