"""Two-mode Gaussian mixture cost (reference: src/projected_langevin_sampling/costs/multimodal.py:7-91)."""
import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


class MultiModalCost(PLSCost):
    """c = -logsumexp(log a + l1, log(1 - a) + l2), mode-1 error y - F + shift (multimodal.py:51), noise squared (:56).
    The reference differentiates it with autograd only (:79-91); the CUDA functor uses the closed form of that
    derivative, -(w1 e1 + w2 e2) / s^2 with w = softmax of the two log terms."""

    native_cost_id = nat.COST_MULTIMODAL
    closed_form_link = -1

    def __init__(self, observation_noise: float, shift: float, bernoulli_noise: float, y_train: torch.Tensor,
                 link_function: PLSLinkFunction):
        super().__init__(link_function=link_function, observation_noise=observation_noise)
        self.shift = shift
        self.bernoulli_noise = bernoulli_noise
        self.y_train = y_train

    def _extra_native_fields(self, c: nat.PlsCost) -> None:
        c.shift = float(self.shift)
        c.bernoulli_noise = float(self.bernoulli_noise)
        # the reference builds these constants as torch tensors in the DEFAULT dtype at call time (multimodal.py:55-72);
        # under a float32 default they are float32-rounded, and that rounding is part of its result
        c.log_weight_1 = float(torch.log(torch.tensor(self.bernoulli_noise)))
        c.log_weight_2 = float(torch.log(torch.tensor(1 - self.bernoulli_noise)))
        c.log_normaliser = float(torch.log(torch.sqrt(2 * torch.tensor([torch.pi]) * (self.observation_noise**2))))

    def predict(self, prediction_samples: torch.Tensor) -> None:
        return None

    def calculate_cost_derivative(self, untransformed_train_prediction_samples: torch.Tensor, force_autograd: bool = True):
        return super().calculate_cost_derivative(untransformed_train_prediction_samples, force_autograd=True)
