from .kernel import PLSKernel
from .projected_langevin_sampling import PLS

__all__ = ["PLS", "PLSKernel"]
