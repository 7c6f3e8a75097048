"""Minimal stand-in for gpytorch, used ONLY by tests/golden/make_golden.py and the CPU reference arm of bench.py (oracle/reference_arm.py) to execute the unmodified reference
modules in the build container (gpytorch / linear_operator are not installed and there is no network).

It provides just the names the reference's hot-path modules touch: kernels.Kernel (dense evaluation instead of
lazy tensors -- numerically the same thing, lazy tensors only re-associate), ScaleKernel, RBFKernel (arithmetic
delegated to oracle.pls_oracle.RBFScaleKernel, the restatement of gpytorch 1.15.2's formula), distributions
.MultivariateNormal, solve.  Nothing in the product or in the GPU tests imports this."""
import torch

from . import distributions, kernels  # noqa: F401


def solve(input, rhs, lhs=None):
    """gpytorch.solve: input^{-1} rhs, or lhs input^{-1} rhs (InducingPointBasis only; a "next" row)."""
    res = torch.linalg.solve(input, rhs)
    return res if lhs is None else lhs @ res


class _Anything:
    """Placeholder for every gpytorch class the reference merely subclasses or names in an annotation while its module
    is imported (models.ExactGP, means.Mean, likelihoods.Likelihood, ...); never instantiated by the golden script."""

    def __init__(self, *args, **kwargs):
        pass


class _Namespace:
    def __getattr__(self, name):
        return _Anything


def __getattr__(name):  # PEP 562: gpytorch.models, gpytorch.means, gpytorch.likelihoods, gpytorch.variational, ...
    if name.startswith("__"):
        raise AttributeError(name)
    return _Namespace()
