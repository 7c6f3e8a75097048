"""InducingPointSelector (reference: src/inducing_point_selectors/base.py:8-34)."""
from abc import ABC, abstractmethod
from typing import Tuple

import torch


class InducingPointSelector(ABC):
    @abstractmethod
    def compute_induce_data(self, x: torch.Tensor, m: int, kernel, **params) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def __call__(self, x: torch.Tensor, m: int, kernel, **params) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.compute_induce_data(x=x, m=m, kernel=kernel, **params)
