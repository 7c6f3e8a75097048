"""Link functions (reference: src/projected_langevin_sampling/link_functions.py:6-80).

`transform` works on any torch tensor (they are applied to prediction samples at predict time); inside the Langevin
step the link and its derivative are evaluated in registers by the CUDA cost functors (csrc/pls_cost.cuh), selected
through `native_id`.
"""
from abc import ABC, abstractmethod

import torch

from .. import _native as nat


class PLSLinkFunction(ABC):
    native_id: int = -1
    jitter: float = 0.0

    @abstractmethod
    def transform(self, y: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        return self.transform(*args, **kwargs)


class ProbitLinkFunction(PLSLinkFunction):
    """Phi(y) via erf, clipped to [jitter, 1 - jitter]  (:30-45)."""

    native_id = nat.LINK_PROBIT

    def __init__(self, jitter: float = 1e-10):
        self.jitter = jitter

    @staticmethod
    def divisor() -> float:
        # the reference divides by sqrt(tensor(2.0)) created in the *default dtype* (:42)
        return float(torch.sqrt(torch.tensor(2.0)))

    def transform(self, y: torch.Tensor) -> torch.Tensor:
        return torch.clip((1 + torch.erf(y / torch.sqrt(torch.tensor(2.0)))) / 2, self.jitter, 1 - self.jitter)


class IdentityLinkFunction(PLSLinkFunction):
    native_id = nat.LINK_IDENTITY

    def transform(self, y: torch.Tensor) -> torch.Tensor:
        return y


class SigmoidLinkFunction(PLSLinkFunction):
    """1 / (1 + exp(-y)), clipped to [jitter, 1 - jitter]  (:58-70)."""

    native_id = nat.LINK_SIGMOID

    def __init__(self, jitter: float = 1e-10):
        self.jitter = jitter

    def transform(self, y: torch.Tensor) -> torch.Tensor:
        return torch.clip(torch.reciprocal(1 + torch.exp(-y)), self.jitter, 1 - self.jitter)


class SquareLinkFunction(PLSLinkFunction):
    native_id = nat.LINK_SQUARE

    def transform(self, y: torch.Tensor) -> torch.Tensor:
        return torch.square(y)
