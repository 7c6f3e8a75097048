"""Student-t cost (reference: src/projected_langevin_sampling/costs/student_t.py:11-110)."""
from dataclasses import dataclass

import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


@dataclass
class StudentTMarginals:
    """Independent Student-t marginals with a shared df (reference: src/distributions.py:8-42)."""

    df: float
    loc: torch.Tensor
    scale: torch.Tensor

    def negative_log_likelihood(self, y: torch.Tensor) -> torch.Tensor:
        dist = torch.distributions.StudentT(df=self.df, loc=self.loc, scale=self.scale)
        return -dist.log_prob(y).mean()


class StudentTCost(PLSCost):
    """c = (nu + 1)/2 log(1 + e^2 / (nu s^2)); identity link: d c/d F = (nu + 1) e / (nu s^2 + e^2) (student_t.py:66-88)."""

    native_cost_id = nat.COST_STUDENT_T
    closed_form_link = nat.LINK_IDENTITY

    def __init__(self, degrees_of_freedom: float, y_train: torch.Tensor, link_function: PLSLinkFunction, scale: float = 1.0):
        super().__init__(link_function=link_function, observation_noise=None)
        self.y_train = y_train
        self.degrees_of_freedom = degrees_of_freedom
        self.scale = scale

    def _extra_native_fields(self, c: nat.PlsCost) -> None:
        c.degrees_of_freedom = float(self.degrees_of_freedom)
        c.scale = float(self.scale)

    def predict(self, prediction_samples: torch.Tensor) -> StudentTMarginals:
        return StudentTMarginals(
            df=self.degrees_of_freedom,
            loc=self.link_function(prediction_samples).mean(dim=1),
            scale=self.scale * torch.ones((prediction_samples.shape[0]), device=prediction_samples.device),
        )
