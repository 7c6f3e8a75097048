from .base import InducingPointSelector
from .conditional_variance import ConditionalVarianceInducingPointSelector

__all__ = ["ConditionalVarianceInducingPointSelector", "InducingPointSelector"]
