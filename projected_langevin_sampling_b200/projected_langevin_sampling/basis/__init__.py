from .base import PLSBasis
from .inducing_point import InducingPointBasis
from .orthonormal import OrthonormalBasis

__all__ = ["InducingPointBasis", "OrthonormalBasis", "PLSBasis"]
