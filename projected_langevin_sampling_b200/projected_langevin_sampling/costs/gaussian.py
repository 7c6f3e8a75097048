"""Gaussian cost (reference: src/projected_langevin_sampling/costs/gaussian.py:11-110)."""
import torch

from ... import _native as nat
from ..link_functions import PLSLinkFunction
from .base import PLSCost


class GaussianCost(PLSCost):
    """c = (mu - y)^2 / (2 s), d c/d F = (F - y) / s with the identity link; `observation_noise` s is a VARIANCE here
    (gaussian.py:71,86)."""

    native_cost_id = nat.COST_GAUSSIAN
    closed_form_link = nat.LINK_IDENTITY

    def __init__(self, observation_noise: float, y_train: torch.Tensor, link_function: PLSLinkFunction):
        super().__init__(link_function=link_function, observation_noise=observation_noise)
        self.y_train = y_train

    def predict(self, prediction_samples: torch.Tensor) -> "DiagonalNormal":
        # the reference returns gpytorch's MultivariateNormal with a diagonal covariance (gaussian.py:40-52)
        return DiagonalNormal(loc=prediction_samples.mean(dim=1), variance=prediction_samples.var(axis=1))


class DiagonalNormal:
    """What the reference's callers read from gpytorch's MultivariateNormal(mean, diag(var)) (gaussian.py:40-52; used as
    `.mean`, `.variance`, `.stddev`, `.covariance_matrix`, `.confidence_region()`, `.sample`, `.log_prob` by the experiment
    layer).  gpytorch keeps the covariance lazy; torch.distributions.MultivariateNormal would run a dense N* x N* Cholesky in
    its constructor and reject a zero variance (J = 1, identical samples) -- so the diagonal is kept as a vector here."""

    def __init__(self, loc: torch.Tensor, variance: torch.Tensor):
        self.loc = loc
        self._variance = variance

    @property
    def mean(self) -> torch.Tensor:
        return self.loc

    @property
    def variance(self) -> torch.Tensor:
        return self._variance

    @property
    def stddev(self) -> torch.Tensor:
        return self._variance.sqrt()

    @property
    def covariance_matrix(self) -> torch.Tensor:
        return torch.diag(self._variance)

    def confidence_region(self):
        """mean -/+ 2 standard deviations, as gpytorch's MultivariateNormal.confidence_region."""
        std2 = self.stddev.mul(2)
        return self.mean.sub(std2), self.mean.add(std2)

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        eps = torch.randn(tuple(sample_shape) + tuple(self.loc.shape), dtype=self.loc.dtype).to(self.loc.device)
        return self.loc + self.stddev * eps

    rsample = sample

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        var = self._variance
        return -0.5 * (((value - self.loc) ** 2 / var).sum(-1) + var.log().sum(-1) + self.loc.shape[-1] * torch.log(torch.tensor(2 * torch.pi, dtype=var.dtype)))
