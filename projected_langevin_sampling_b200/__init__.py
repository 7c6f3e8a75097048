"""projected_langevin_sampling_b200 -- the Projected Langevin Sampling gradient-flow hot path on NVIDIA B200.

Mirrors the import layout of the reference's `src` package for that path:

    reference                                              here
    src.projected_langevin_sampling (PLS, PLSKernel)       projected_langevin_sampling_b200.projected_langevin_sampling
    src.projected_langevin_sampling.basis                  ....projected_langevin_sampling.basis   (OrthonormalBasis)
    src.projected_langevin_sampling.costs                  ....projected_langevin_sampling.costs
    src.projected_langevin_sampling.link_functions         ....projected_langevin_sampling.link_functions
    src.inducing_point_selectors                           projected_langevin_sampling_b200.inducing_point_selectors
    src.samplers / src.utils                               projected_langevin_sampling_b200.samplers / .utils
    gpytorch.kernels.ScaleKernel / RBFKernel               projected_langevin_sampling_b200.kernels

All arithmetic of the path runs in libpls_b200.so (hand-written sm_100a CUDA, C ABI in include/pls_b200.h).  There is no
CPU or PyTorch fallback: importing works anywhere, computing needs the built library and a B200.
"""
from . import kernels  # noqa: F401
from .inducing_point_selectors import ConditionalVarianceInducingPointSelector  # noqa: F401
from .kernels import LinearKernel, RBFKernel, ScaleKernel  # noqa: F401
from .projected_langevin_sampling import PLS, PLSKernel  # noqa: F401
from .projected_langevin_sampling.basis import InducingPointBasis, OrthonormalBasis  # noqa: F401
from .utils import set_seed  # noqa: F401

__version__ = "0.1.0"
