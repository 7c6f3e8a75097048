"""Cost base class (reference: src/projected_langevin_sampling/costs/base.py:8-133).

The per-point arithmetic of every cost -- value and derivative w.r.t. the untransformed prediction, including the link
function -- lives in the CUDA library (csrc/pls_cost.cuh) and is selected with a `pls_cost` struct; this class keeps the
reference's Python surface and hands tensors to it.  Where the reference falls back to autograd
(costs/base.py:68-84: O(N^2 J) vmap(jacfwd)), the library evaluates the same derivative in closed form by the chain rule.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch

from ... import _native as nat
from ... import ops
from ..link_functions import PLSLinkFunction, ProbitLinkFunction


class PLSCost(ABC):
    native_cost_id: int = -1
    #: link (native id) for which the reference has a hand-written derivative
    closed_form_link: int = -1

    def __init__(self, link_function: PLSLinkFunction, observation_noise: Optional[float] = None):
        self.observation_noise = observation_noise
        self.link_function = link_function
        self.y_train: Optional[torch.Tensor] = None
        self._y_dev = None

    # ---- native plumbing -----------------------------------------------------------------------------------------
    def _extra_native_fields(self, c: nat.PlsCost) -> None:
        pass

    def native(self, force_autograd: bool = False) -> nat.PlsCost:
        """The `pls_cost` struct the CUDA functors read (include/pls_b200.h)."""
        link = self.link_function
        if getattr(link, "native_id", -1) < 0:
            raise TypeError(f"{type(link).__name__} has no CUDA implementation (no CPU fallback)")
        c = nat.PlsCost()
        c.cost_id = self.native_cost_id
        c.link_id = link.native_id
        c.closed_form = int(link.native_id == self.closed_form_link and not force_autograd)
        c.observation_noise = float(self.observation_noise) if self.observation_noise is not None else 1.0
        c.shift, c.bernoulli_noise, c.degrees_of_freedom, c.scale = 0.0, 0.5, 1.0, 1.0
        c.link_jitter = float(getattr(link, "jitter", 0.0))
        c.probit_divisor = ProbitLinkFunction.divisor()
        c.log_weight_1 = c.log_weight_2 = c.log_normaliser = 0.0
        self._extra_native_fields(c)
        return c

    def y_device(self, device=None) -> torch.Tensor:
        """Training labels as a float64 CUDA vector (cached)."""
        dev = device or torch.device("cuda", torch.cuda.current_device())
        version = getattr(self.y_train, "_version", 0)  # an in-place edit of y_train bumps it: the device copy is refreshed
        if self._y_dev is None or self._y_dev.device != dev or self._y_src is not self.y_train or self._y_version != version:
            self._y_dev = ops.as_device_f64(self.y_train.reshape(-1), dev)
            self._y_src, self._y_version = self.y_train, version
        return self._y_dev

    # ---- reference API -------------------------------------------------------------------------------------------------
    @abstractmethod
    def predict(self, prediction_samples: torch.Tensor):
        raise NotImplementedError()

    def calculate_cost(self, untransformed_train_prediction_samples: torch.Tensor) -> torch.Tensor:
        """c_j = sum_n c(y_n, F[n, j])  -> (J,)  (costs/*.py `calculate_cost`)."""
        f = ops.as_device_f64(untransformed_train_prediction_samples)
        ctx = nat.context(f.device)
        return ops.cost_value(ctx, self.native(), self.y_device(f.device), f)

    def calculate_cost_derivative(self, untransformed_train_prediction_samples: torch.Tensor,
                                  force_autograd: bool = False) -> torch.Tensor:
        """d_2 c(y, F) -> (N, J)  (costs/*.py `calculate_cost_derivative`; `force_autograd` selects the chain-rule form
        that equals the reference's autograd result)."""
        f = ops.as_device_f64(untransformed_train_prediction_samples)
        ctx = nat.context(f.device)
        return ops.cost_derivative(ctx, self.native(force_autograd), self.y_device(f.device), f)

    def sample_observation_noise(self, number_of_particles: int, seed: Optional[int] = None) -> torch.Tensor:
        """costs/base.py:86-115 (note: observation_noise is used as a std here)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.observation_noise is None:
            return torch.zeros(number_of_particles, device=dev)
        generator = torch.Generator().manual_seed(seed) if seed is not None else None
        noise = torch.normal(mean=0.0, std=self.observation_noise, size=(number_of_particles,), generator=generator).flatten()
        return noise.to(dev)

    def predict_samples(self, untransformed_samples: torch.Tensor, observation_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """link(untransformed + observation noise)  (costs/base.py:117-133)."""
        if observation_noise is None:
            observation_noise = self.sample_observation_noise(number_of_particles=untransformed_samples.shape[1])
        observation_noise = observation_noise.to(untransformed_samples.device, untransformed_samples.dtype)
        return self.link_function(untransformed_samples + observation_noise[None, :])
