"""Base-kernel specifications.

The reference passes `gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=D))` objects around and
only ever *calls* them (README.md:144-146, experiments/uci/regression/main.py:169-213).  gpytorch is not a dependency of
this package, so light stand-ins with the same attribute names are provided; a real gpytorch kernel is accepted too
(duck-typed on `.base_kernel.lengthscale` / `.outputscale`).  Only RBF/ARD x Scale (and the reference's linear test
double, mockers/kernel.py:8-23) have a CUDA path; anything else raises TypeError -- there is no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _native as nat
from . import ops


class RBFKernel:
    """exp(-0.5 * sum_d ((x_d - x'_d) / lengthscale_d)^2).  `lengthscale`: float or tensor (1, D) / (D,)."""

    def __init__(self, ard_num_dims: Optional[int] = None, lengthscale=1.0):
        self.ard_num_dims = ard_num_dims
        self.lengthscale = lengthscale

    def __call__(self, x1, x2=None, diag: bool = False, **params):
        return dense_gram(ScaleKernel(self, outputscale=1.0), x1, x1 if x2 is None else x2, diag=diag)

    forward = __call__


class ScaleKernel:
    """outputscale * base_kernel(x, x')."""

    def __init__(self, base_kernel: RBFKernel, outputscale: float = 1.0):
        self.base_kernel = base_kernel
        self.outputscale = outputscale

    def __call__(self, x1, x2=None, diag: bool = False, **params):
        return dense_gram(self, x1, x1 if x2 is None else x2, diag=diag)

    forward = __call__


class LinearKernel:
    """x1 @ x2^T: the reference's MockKernel (mockers/kernel.py:8-23), kept so its golden vectors replay on the GPU."""

    def __call__(self, x1, x2=None, diag: bool = False, **params):
        return dense_gram(self, x1, x1 if x2 is None else x2, diag=diag)

    forward = __call__


@dataclass
class KernelSpec:
    kernel_id: int
    lengthscale: List[float]  # length D
    outputscale: float

    @property
    def inv_lengthscale(self) -> List[float]:
        return [1.0 / v for v in self.lengthscale]

    @property
    def log_outputscale(self) -> float:
        return math.log(self.outputscale) if self.kernel_id == nat.KERNEL_RBF else 0.0


def _to_float_list(value, d: int) -> List[float]:
    if isinstance(value, (int, float)):  # python scalars keep their full double precision
        return [float(value)] * d
    t = torch.as_tensor(value, dtype=None if isinstance(value, torch.Tensor) else torch.float64)
    t = t.detach().to("cpu", torch.float64).reshape(-1)
    if t.numel() == 1:
        return [float(t[0])] * d
    if t.numel() != d:
        raise ValueError(f"lengthscale has {t.numel()} entries but the inputs have D={d}")
    return [float(v) for v in t]


def kernel_spec(kernel, d: int) -> KernelSpec:
    """Read (kind, lengthscale, outputscale) back from a kernel object (ours or gpytorch's)."""
    name = type(kernel).__name__
    if isinstance(kernel, LinearKernel) or name in ("MockKernel", "LinearKernel"):
        return KernelSpec(nat.KERNEL_LINEAR, [1.0] * d, 1.0)
    base = getattr(kernel, "base_kernel", None)
    if base is not None and hasattr(kernel, "outputscale") and "RBF" in type(base).__name__ and hasattr(base, "lengthscale"):
        return KernelSpec(nat.KERNEL_RBF, _to_float_list(base.lengthscale, d), _to_float_list(kernel.outputscale, 1)[0])
    if "RBF" in name and hasattr(kernel, "lengthscale"):
        return KernelSpec(nat.KERNEL_RBF, _to_float_list(kernel.lengthscale, d), 1.0)
    raise TypeError(
        f"{name}: only ScaleKernel(RBFKernel) (RBF/ARD x scale) and the linear test kernel have a CUDA path; "
        "projected_langevin_sampling_b200 has no CPU fallback for other kernels"
    )


def _as_points(x: torch.Tensor) -> torch.Tensor:
    x = ops.as_device_f64(x)
    return x.unsqueeze(-1) if x.dim() == 1 else x


def prepare_pair(ctx, spec: KernelSpec, rows: torch.Tensor, cols: torch.Tensor):
    """Augmented layouts of a (row set, column set) pair, centred on the column set's mean; the outputscale goes to
    the column side only."""
    centre = cols.mean(dim=0).tolist() if spec.kernel_id == nat.KERNEL_RBF else [0.0] * cols.shape[1]
    ra = ops.prepare_points(ctx, spec.kernel_id, rows, spec.inv_lengthscale, centre, 0.0)
    ca = ops.prepare_points(ctx, spec.kernel_id, cols, spec.inv_lengthscale, centre, spec.log_outputscale)
    return ra, ca


def dense_gram(kernel, x1: torch.Tensor, x2: torch.Tensor, diag: bool = False) -> torch.Tensor:
    """k(x1, x2) as a dense CUDA matrix (setup / predict-time sizes only)."""
    x1, x2 = _as_points(x1), _as_points(x2)
    spec = kernel_spec(kernel, x1.shape[1])
    ctx = nat.context(x1.device)
    ra, ca = prepare_pair(ctx, spec, x1, x2)
    g = ops.gram(ctx, spec.kernel_id, ra, ca, x1.shape[1])
    return g.diagonal().clone() if diag else g
