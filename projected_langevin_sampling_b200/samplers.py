"""Samplers (reference: src/samplers.py:6-62).

`sample_multivariate_normal` keeps the reference's noise stream exactly: eigh(cov), clip, ONE torch.normal((dim, *size))
call on the CPU generator (global, or a fresh one seeded with `seed`), in the default dtype.  `langevin_noise` is the
draw OrthonormalBasis._calculate_particle_update makes every step (orthonormal.py:141-145) with cov = I, for which
eigh(I) = (1, I) exactly, so it is the raw torch.normal((M_k, J)) tensor -- the ~70 ms eigh(eye(M_k)) per step the
reference spends there is skipped, the numbers are identical."""
from typing import Optional, Tuple

import torch


def sample_multivariate_normal(mean: torch.Tensor, cov: torch.Tensor, size: Optional[Tuple[int, ...]] = None,
                               seed: Optional[int] = None) -> torch.Tensor:
    generator = torch.Generator().manual_seed(seed) if seed is not None else None
    size = (1,) if not size else size
    # The decomposition runs on the HOST (LAPACK) whatever the device of `cov`: eigenvector signs (and bases of repeated
    # eigenvalues) differ between LAPACK and cuSOLVER, and the draw V sqrt(lambda) z depends on them -- the reference's runs
    # and golden vectors are CPU ones.  These matrices are (M_k + N*) x (M_k + N*): predict-time sizes.
    eigenvalues, eigenvectors = torch.linalg.eigh(cov.detach().cpu())
    eigenvalues, eigenvectors = eigenvalues.to(cov.device), eigenvectors.to(cov.device)
    eigenvalues = torch.clip(eigenvalues, 0, None)
    normal_sample = torch.normal(mean=0.0, std=1.0, size=(eigenvalues.shape[0], *size), generator=generator)
    dev = eigenvectors.device
    normal_sample = normal_sample.to(device=dev, dtype=eigenvectors.dtype)
    return torch.real(mean.to(dev)[:, None] + eigenvectors @ torch.diag(torch.sqrt(eigenvalues)) @ normal_sample).T


def langevin_noise(approximation_dimension: int, number_of_particles: int) -> torch.Tensor:
    """xi (M_k, J) from torch's GLOBAL CPU generator, default dtype -- the reference's per-step draw."""
    return torch.normal(mean=0.0, std=1.0, size=(approximation_dimension, number_of_particles))


def sample_point(x: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
    generator = torch.Generator().manual_seed(seed) if seed is not None else None
    random_idx = torch.randperm(x.shape[0], generator=generator)[0]
    return x[random_idx : random_idx + 1, ...]
