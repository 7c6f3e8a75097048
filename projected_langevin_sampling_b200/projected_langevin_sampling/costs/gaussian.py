"""Gaussian cost (reference: src/projected_langevin_sampling/costs/gaussian.py:11-110)."""
import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


class GaussianCost(PLSCost):
    """c = (mu - y)^2 / (2 s), d c/d F = (F - y) / s with the identity link; `observation_noise` s is a VARIANCE here
    (gaussian.py:71,86)."""

    native_cost_id = nat.COST_GAUSSIAN
    closed_form_link = nat.LINK_IDENTITY

    def __init__(self, observation_noise: float, y_train: torch.Tensor, link_function: PLSLinkFunction):
        super().__init__(link_function=link_function, observation_noise=observation_noise)
        self.y_train = y_train

    def predict(self, prediction_samples: torch.Tensor) -> torch.distributions.MultivariateNormal:
        # the reference returns gpytorch's MultivariateNormal (gaussian.py:40-52); torch's has the same mean / covariance
        return torch.distributions.MultivariateNormal(
            loc=prediction_samples.mean(dim=1),
            covariance_matrix=torch.diag(prediction_samples.var(axis=1)),
            validate_args=False,
        )
