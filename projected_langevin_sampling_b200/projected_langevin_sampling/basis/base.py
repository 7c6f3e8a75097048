"""PLSBasis (reference: src/projected_langevin_sampling/basis/base.py:7-193)."""
from abc import ABC, abstractmethod
from typing import Optional

import torch


class PLSBasis(ABC):
    def __init__(self, additional_predictive_noise_distribution: Optional[torch.distributions.Distribution] = None):
        self.additional_predictive_noise_distribution = additional_predictive_noise_distribution

    @property
    def approximation_dimension(self) -> int:
        raise NotImplementedError

    def _initialise_particles_noise(self, number_of_particles: int, seed: Optional[int] = None, mean: float = 0.0,
                                    stdev: float = 1.0) -> torch.Tensor:
        """torch.normal(mean, stdev, (M, J)) on the CPU generator (fresh one when `seed` is given)  (base.py:39-63)."""
        generator = torch.Generator().manual_seed(seed) if seed is not None else None
        return torch.normal(mean=mean, std=stdev, size=(self.approximation_dimension, number_of_particles), generator=generator)

    @abstractmethod
    def _initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        raise NotImplementedError

    def initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        """(M, J) particles on the GPU, float64 (the reference returns them on cuda when available, base.py:81-102)."""
        particles = self._initialise_particles(number_of_particles=number_of_particles, noise_only=noise_only, seed=seed)
        return particles.to(device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.float64)

    @abstractmethod
    def calculate_untransformed_train_prediction_samples(self, particles: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def calculate_energy_potential(self, particles: torch.Tensor, cost: torch.Tensor) -> float:
        raise NotImplementedError

    @abstractmethod
    def _calculate_particle_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float, **kwargs) -> torch.Tensor:
        raise NotImplementedError

    def calculate_particle_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float, **kwargs) -> torch.Tensor:
        assert (
            particles.shape[0] == self.approximation_dimension
        ), f"Particles have shape {particles.shape} but requires ({self.approximation_dimension}, J) dimension."
        return self._calculate_particle_update(particles=particles, cost_derivative=cost_derivative, step_size=step_size, **kwargs)

    @abstractmethod
    def sample_predictive_noise(self, particles: torch.Tensor, x: torch.Tensor):
        raise NotImplementedError

    @abstractmethod
    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        raise NotImplementedError
