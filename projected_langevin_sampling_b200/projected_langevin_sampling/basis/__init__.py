from .base import PLSBasis
from .orthonormal import OrthonormalBasis

__all__ = ["OrthonormalBasis", "PLSBasis"]
