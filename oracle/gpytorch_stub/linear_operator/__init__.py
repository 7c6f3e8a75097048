"""Stand-in for linear_operator (only a type annotation in the reference's mockers/kernel.py uses it)."""
from . import operators  # noqa: F401
