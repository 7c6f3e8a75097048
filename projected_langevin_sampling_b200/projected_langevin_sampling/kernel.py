"""PLSKernel: the r-kernel wrapper (reference: src/projected_langevin_sampling/kernel.py:5-79).

r(x, x') = (1/S) sum_s k(x, z_s) k(x', z_s) over the UNIQUE rows of the approximation samples.  In the Langevin loop
it is only the container of `base_kernel` (orthonormal.py:36-41); `forward` is predict-time and is built from the
library's dense Gram and DGEMM entry points."""
from typing import Optional

import torch

from .. import _native as nat
from .. import ops
from ..kernels import dense_gram


class PLSKernel:
    is_stationary: bool = False

    def __init__(self, base_kernel, approximation_samples: torch.Tensor, **kwargs):
        self.base_kernel = base_kernel
        self.approximation_samples = approximation_samples

    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size([])

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, additional_approximation_samples: Optional[torch.Tensor] = None,
                last_dim_is_batch: bool = False, diag: bool = False, **params) -> torch.Tensor:
        parts = [ops.as_device_f64(self.approximation_samples)]
        if additional_approximation_samples is not None:
            parts.append(ops.as_device_f64(additional_approximation_samples))
        samples = torch.cat(parts, dim=0).unique(dim=0)
        g1 = dense_gram(self.base_kernel, x1, samples)  # (n1, S)
        g2t = dense_gram(self.base_kernel, samples, x2)  # (S, n2) = k(x2, samples)^T
        ctx = nat.context(g1.device)
        out = torch.empty((g1.shape[0], g2t.shape[1]), dtype=torch.float64, device=g1.device)
        ops.gemm(ctx, g1, g2t, out)
        res = torch.mul(torch.div(1, samples.shape[0]), out)
        return res.diag() if diag else res

    def __call__(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        return self.forward(x1, x1 if x2 is None else x2, **kwargs)

    def num_outputs_per_input(self, x1: torch.Tensor, x2: torch.Tensor) -> int:
        return 1
