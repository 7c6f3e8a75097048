"""CPU oracle for the Projected Langevin Sampling (PLS) hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`projected_langevin_sampling_b200`) never does: its path is the CUDA library and it fails loudly without it.

What it is: a float64 torch-CPU / numpy restatement of the reference algorithm
(jswu18/projected-langevin-sampling), written from the reference's behaviour, each function citing the
reference file:line it follows (paths relative to the reference root).  gpytorch / linear_operator are not
installed in this image, so the RBF/ARD x Scale kernel arithmetic is restated from gpytorch 1.15.2's published
formula (`gpytorch/kernels/rbf_kernel.py`, `gpytorch/kernels/kernel.py: sq_dist`, `scale_kernel.py`).

Parity pinning status
  * pinned by the reference's own golden vectors (tests/test_basis.py, test_costs.py, test_pls_kernel.py,
    test_inducing_point_selectors.py, test_samplers.py, test_set_seed.py): ONB dimension / init / forward /
    energy / predictive noise / predict, all cost values and derivatives, the r-kernel, the selector, the sampler.
  * pinned by executing the UNMODIFIED reference modules in the build container behind a gpytorch stub
    (tests/golden/make_golden.py -> tests/golden/*.npz): the Langevin update `_calculate_particle_update`, whole
    Langevin trajectories, the autograd multimodal derivative, selector runs with m > 2.
  * RBF/ARD kernel VALUES: parity unpinned against real gpytorch (absent here); restated from its formula.
"""
from __future__ import annotations

import math
import os
import random
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch

F64 = torch.float64


# ----------------------------------------------------------------------------------------------------------------
# seeding -- src/utils.py:8-22
# ----------------------------------------------------------------------------------------------------------------
def set_seed(seed: int = 42) -> None:
    """Seeds numpy, python `random` and torch global generators, as src/utils.py:8-22 does."""
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)


# ----------------------------------------------------------------------------------------------------------------
# base kernels k(x, x')
# ----------------------------------------------------------------------------------------------------------------
def _gpytorch_sq_dist(a: torch.Tensor, b: torch.Tensor, same: bool) -> torch.Tensor:
    """Squared distances the way gpytorch 1.15.2 `Kernel.covar_dist(square_dist=True)` forms them:
    centre both inputs on a.mean(0), expand |a|^2 + |b|^2 - 2ab as ONE matmul of [-2a, |a|^2, 1] with
    [b, 1, |b|^2]^T, zero the diagonal when the two inputs are the same tensor, clamp at 0."""
    shift = a.mean(dim=-2, keepdim=True)
    a = a - shift
    a_sq = a.pow(2).sum(dim=-1, keepdim=True)
    ones_a = torch.ones_like(a_sq)
    if same:
        b, b_sq, ones_b = a, a_sq, ones_a
    else:
        b = b - shift
        b_sq = b.pow(2).sum(dim=-1, keepdim=True)
        ones_b = torch.ones_like(b_sq)
    left = torch.cat([-2.0 * a, a_sq, ones_a], dim=-1)
    right = torch.cat([b, ones_b, b_sq], dim=-1)
    d2 = left.matmul(right.transpose(-2, -1))
    if same:
        d2.diagonal(dim1=-2, dim2=-1).fill_(0)
    return d2.clamp_min_(0)


@dataclass
class RBFScaleKernel:
    """sigma^2 * exp(-0.5 * sum_d ((x_d - x'_d) / l_d)^2): gpytorch ScaleKernel(RBFKernel(ard_num_dims=D)).

    Call sites in the reference: basis/orthonormal.py:36-41, inducing_point_selectors/conditional_variance.py:66-69,81-89,
    kernel.py:48-68.  `lengthscale` is a scalar or a (D,) / (1, D) tensor, `outputscale` a scalar."""

    lengthscale: object = 1.0
    outputscale: float = 1.0

    def _ls(self, like: torch.Tensor) -> torch.Tensor:
        return torch.as_tensor(self.lengthscale, dtype=like.dtype).reshape(1, -1)

    def __call__(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, diag: bool = False) -> torch.Tensor:
        x2 = x1 if x2 is None else x2
        if x1.ndim == 1:
            x1 = x1.unsqueeze(-1)
        if x2.ndim == 1:
            x2 = x2.unsqueeze(-1)
        same = x1.shape == x2.shape and bool(torch.equal(x1, x2))
        a = x1 / self._ls(x1)
        b = x2 / self._ls(x2)
        if diag:
            if same:  # gpytorch returns exact zeros distance -> exactly outputscale
                return torch.full((x1.shape[0],), float(self.outputscale), dtype=x1.dtype)
            d2 = (a - b).pow(2).sum(-1)
            return torch.exp(d2 / -2.0) * float(self.outputscale)
        d2 = _gpytorch_sq_dist(a, b, same)
        return torch.exp(d2 / -2.0) * float(self.outputscale)

    forward = __call__


@dataclass
class LinearKernel:
    """x1 @ x2^T -- the reference's test double mockers/kernel.py:8-23 (`MockKernel`)."""

    def __call__(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, diag: bool = False) -> torch.Tensor:
        x2 = x1 if x2 is None else x2
        if diag:
            return (x1 * x2).sum(-1)
        return x1 @ x2.transpose(-1, -2)

    forward = __call__


def r_kernel(
    base_kernel: Callable,
    approximation_samples: torch.Tensor,
    x1: torch.Tensor,
    x2: torch.Tensor,
    additional_approximation_samples: Optional[torch.Tensor] = None,
    diag: bool = False,
) -> torch.Tensor:
    """PLSKernel.forward, kernel.py:31-76: r(x, x') = (1/S) sum_s k(x, z_s) k(x', z_s) over the UNIQUE rows of
    [approximation_samples; additional_approximation_samples]."""
    parts = [approximation_samples]
    if additional_approximation_samples is not None:
        parts.append(additional_approximation_samples)
    samples = torch.cat(parts, dim=0).unique(dim=0)
    g1 = base_kernel(x1, samples)
    g2 = base_kernel(x2, samples)
    res = torch.mul(torch.div(1, samples.shape[0]), g1 @ g2.T)
    return res.diag() if diag else res


# ----------------------------------------------------------------------------------------------------------------
# sampler -- src/samplers.py:6-44
# ----------------------------------------------------------------------------------------------------------------
def sample_multivariate_normal(
    mean: torch.Tensor, cov: torch.Tensor, size: Optional[Tuple[int, ...]] = None, seed: Optional[int] = None
) -> torch.Tensor:
    """eigh(cov), clip eigenvalues at 0, z = torch.normal(0, 1, (dim, *size)) from the global CPU generator (or a
    fresh one seeded with `seed`), returns (mean + V sqrt(L) z)^T  (samplers.py:22-44)."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    size = (1,) if not size else size
    lam, vec = torch.linalg.eigh(cov)
    lam = torch.clip(lam, 0, None)
    z = torch.normal(mean=0.0, std=1.0, size=(lam.shape[0], *size), generator=gen)
    return torch.real(mean[:, None] + vec @ torch.diag(torch.sqrt(lam)) @ z).T


def langevin_noise(m_k: int, j: int) -> torch.Tensor:
    """The xi of one Langevin step, exactly as basis/orthonormal.py:141-145 draws it:
    sample_multivariate_normal(zeros(M_k), eye(M_k), size=(J,)).T  -> (M_k, J), consuming the GLOBAL torch CPU
    generator with one torch.normal((M_k, J)) call in the default dtype."""
    return sample_multivariate_normal(mean=torch.zeros(m_k), cov=torch.eye(m_k), size=(j,)).T


# ----------------------------------------------------------------------------------------------------------------
# link functions -- src/projected_langevin_sampling/link_functions.py:30-80
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Link:
    name: str  # "identity" | "sigmoid" | "probit" | "square"
    jitter: float = 1e-10

    def __call__(self, f: torch.Tensor) -> torch.Tensor:
        if self.name == "identity":  # :48-55
            return f
        if self.name == "square":  # :73-80
            return torch.square(f)
        if self.name == "sigmoid":  # :58-70 (clipped to [jitter, 1-jitter])
            return torch.clip(torch.reciprocal(1 + torch.exp(-f)), self.jitter, 1 - self.jitter)
        if self.name == "probit":  # :30-45 (erf form, sqrt(2) as a float32-default tensor in the reference)
            return torch.clip((1 + torch.erf(f / torch.sqrt(torch.tensor(2.0)))) / 2, self.jitter, 1 - self.jitter)
        raise ValueError(self.name)

    def derivative(self, f: torch.Tensor) -> torch.Tensor:
        """d link / d f, what torch autograd gives through the clip (0 where clipped)."""
        if self.name == "identity":
            return torch.ones_like(f)
        if self.name == "square":
            return 2 * f
        if self.name == "sigmoid":
            s = torch.reciprocal(1 + torch.exp(-f))
            inside = (s >= self.jitter) & (s <= 1 - self.jitter)
            return torch.where(inside, s * (1 - s), torch.zeros_like(f))
        if self.name == "probit":
            root2 = torch.sqrt(torch.tensor(2.0)).to(f.dtype)
            p = (1 + torch.erf(f / root2)) / 2
            inside = (p >= self.jitter) & (p <= 1 - self.jitter)
            dens = torch.exp(-torch.square(f / root2)) / (root2 * math.sqrt(math.pi))
            return torch.where(inside, dens, torch.zeros_like(f))
        raise ValueError(self.name)


# ----------------------------------------------------------------------------------------------------------------
# costs -- src/projected_langevin_sampling/costs/*.py
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Cost:
    """One of the reference's likelihood costs.

    kind:   "gaussian" | "bernoulli" | "poisson" | "multimodal" | "student_t"
    y:      training labels (N,)
    link:   Link
    observation_noise (gaussian: used as a VARIANCE, gaussian.py:71,86; multimodal: squared, multimodal.py:56),
    shift / bernoulli_noise (multimodal.py:21-33), degrees_of_freedom / scale (student_t.py:23-38)."""

    kind: str
    y: torch.Tensor
    link: Link
    observation_noise: Optional[float] = None
    shift: float = 0.0
    bernoulli_noise: float = 0.5
    degrees_of_freedom: float = 3.0
    scale: float = 1.0

    def __post_init__(self):
        if self.kind == "bernoulli":  # bernoulli.py:32 force-casts labels to double
            self.y = self.y.type(torch.double)

    # ---- c(y, F) summed over n -> (J,) ---------------------------------------------------------------------
    def value(self, f: torch.Tensor) -> torch.Tensor:
        y = self.y[:, None]
        mu = self.link(f)
        if self.kind == "gaussian":  # gaussian.py:54-73
            err = mu - y
            return (1 / (2 * self.observation_noise)) * (err * err).sum(dim=0)
        if self.kind == "bernoulli":  # bernoulli.py:48-62
            return -torch.log(mu).T @ self.y - torch.log(1 - mu).T @ (1 - self.y)
        if self.kind == "poisson":  # poisson.py:47-66  (log|F| of the UNtransformed samples)
            return (-2 * torch.multiply(y, torch.log(torch.abs(f))) + mu).sum(dim=0)
        if self.kind == "student_t":  # student_t.py:55-72
            err = mu - y
            nu = self.degrees_of_freedom
            return 0.5 * (nu + 1) * torch.log(1 + torch.square(err) / (nu * (self.scale**2))).sum(dim=0)
        if self.kind == "multimodal":  # multimodal.py:37-77
            s2 = self.observation_noise**2
            e1 = y - mu + self.shift
            e2 = y - mu
            lognorm = torch.log(torch.sqrt(2 * torch.tensor([torch.pi]) * s2))
            l1 = -0.5 * (torch.square(e1) / s2) - lognorm
            l2 = -0.5 * (torch.square(e2) / s2) - lognorm
            stacked = torch.stack(
                [torch.log(torch.tensor(self.bernoulli_noise)) + l1, torch.log(torch.tensor(1 - self.bernoulli_noise)) + l2]
            )
            return -torch.logsumexp(stacked, dim=0).sum(axis=0)
        raise ValueError(self.kind)

    # ---- d c / d F  (N, J) ----------------------------------------------------------------------------------
    def _has_closed_form(self) -> bool:
        return (self.kind, self.link.name) in {
            ("gaussian", "identity"),  # gaussian.py:75-88
            ("bernoulli", "sigmoid"),  # bernoulli.py:64-77
            ("poisson", "square"),  # poisson.py:68-82
            ("student_t", "identity"),  # student_t.py:74-88
        }

    def derivative(self, f: torch.Tensor, force_autograd: bool = False) -> torch.Tensor:
        """calculate_cost_derivative: the reference's closed form when the link matches, else its autograd fallback
        (costs/base.py:68-84).  MultiModalCost is autograd-only (multimodal.py:79-91)."""
        if self._has_closed_form() and not force_autograd:
            y = self.y[:, None]
            if self.kind == "gaussian":
                return (1 / self.observation_noise) * (self.link(f) - y)
            if self.kind == "bernoulli":
                p = self.link(f)  # the CLIPPED probability
                return -torch.mul(y, 1 - p) + torch.mul(1 - y, p)
            if self.kind == "poisson":
                return -2 * torch.divide(y, f) + 2 * f
            if self.kind == "student_t":
                err = self.link(f) - y
                nu = self.degrees_of_freedom
                return (nu + 1) * torch.divide(err, (nu * (self.scale**2) + torch.square(err)))
        return self.derivative_autograd(f)

    def derivative_autograd(self, f: torch.Tensor) -> torch.Tensor:
        """costs/base.py:68-84: vmap(jacfwd(cost)) over particle columns.  O(N^2 J): small inputs only."""
        jac = torch.vmap(torch.func.jacfwd(self.value), in_dims=2)(f[:, None, :])
        return jac.permute(2, 0, 1, 3).reshape(f.shape)

    def derivative_chain_rule(self, f: torch.Tensor) -> torch.Tensor:
        """Closed form of what autograd computes: (d cost / d mu) * link'(F) (+ the explicit F term for Poisson).
        This is the formula the CUDA cost functors implement for (cost, link) pairs with no reference closed form;
        tests check it against `derivative_autograd` on small inputs."""
        y = self.y[:, None]
        mu = self.link(f)
        dmu = self.link.derivative(f)
        if self.kind == "gaussian":
            return (mu - y) / self.observation_noise * dmu
        if self.kind == "bernoulli":
            return (-y / mu + (1 - y) / (1 - mu)) * dmu
        if self.kind == "poisson":
            return -2 * y / f + dmu
        if self.kind == "student_t":
            err = mu - y
            nu = self.degrees_of_freedom
            return (nu + 1) * err / (nu * self.scale**2 + err * err) * dmu
        if self.kind == "multimodal":
            s2 = self.observation_noise**2
            e1 = y - mu + self.shift
            e2 = y - mu
            # the reference's log-weights are torch tensors in the default dtype (multimodal.py:66-72)
            a1 = float(torch.log(torch.tensor(self.bernoulli_noise))) - 0.5 * e1 * e1 / s2
            a2 = float(torch.log(torch.tensor(1 - self.bernoulli_noise))) - 0.5 * e2 * e2 / s2
            w = torch.softmax(torch.stack([a1, a2]), dim=0)
            return -(w[0] * e1 + w[1] * e2) / s2 * dmu
        raise ValueError(self.kind)

    # ---- prediction-side helpers (costs/base.py:86-133) -------------------------------------------------------
    def sample_observation_noise(self, number_of_particles: int, seed: Optional[int] = None) -> torch.Tensor:
        if self.observation_noise is None:
            return torch.zeros(number_of_particles)
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        return torch.normal(mean=0.0, std=self.observation_noise, size=(number_of_particles,), generator=gen).flatten()

    def predict_samples(self, untransformed: torch.Tensor, observation_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if observation_noise is None:
            observation_noise = self.sample_observation_noise(untransformed.shape[1])
        return self.link(untransformed + observation_noise[None, :])


# ----------------------------------------------------------------------------------------------------------------
# orthonormal basis -- src/projected_langevin_sampling/basis/orthonormal.py, basis/base.py
# ----------------------------------------------------------------------------------------------------------------
class OrthonormalBasisOracle:
    """Eigen-basis of (1/M) k(Z, Z); particles are coordinates in it (orthonormal.py:22-68).

    `eig=(eigenvalues, eigenvectors)` injects a precomputed decomposition of (1/M) K_zz (ascending, as
    torch.linalg.eigh returns) so that oracle and CUDA path share eigenvector signs (hard part 2)."""

    def __init__(
        self,
        base_kernel: Callable,
        x_induce: torch.Tensor,
        x_train: torch.Tensor,
        eigenvalue_threshold: float = 0.0,
        eig: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        r_kernel_samples: Optional[torch.Tensor] = None,
        additional_predictive_noise_distribution=None,
    ):
        self.base_kernel = base_kernel
        self.x_induce = x_induce
        self.x_train = x_train
        self.r_kernel_samples = x_induce if r_kernel_samples is None else r_kernel_samples
        self.additional_predictive_noise_distribution = additional_predictive_noise_distribution
        self.k_zz = base_kernel(x_induce, x_induce)  # (M, M)  :36-38
        self.k_zx = base_kernel(x_induce, x_train)  # (M, N)  :39-41
        if eig is None:
            lam, vec = torch.linalg.eigh((1 / x_induce.shape[0]) * self.k_zz)  # :46-48
        else:
            lam, vec = eig
        keep = torch.where(lam > eigenvalue_threshold)[0]  # strict, ascending order kept  :52-56
        self.eigenvalues = lam[keep].real
        self.eigenvectors = vec[:, keep].real
        m_k = self.eigenvalues.shape[0]
        # V~ = V / sqrt(M_k * lambda)  -- uses M_k, not M  (:63-68)
        self.scaled_eigenvectors = torch.multiply(
            torch.reciprocal(torch.sqrt(m_k * self.eigenvalues))[None, :], self.eigenvectors
        )

    @property
    def approximation_dimension(self) -> int:  # :70-76
        return self.eigenvalues.shape[0]

    def initialise_particles(self, number_of_particles: int, seed: Optional[int] = None) -> torch.Tensor:
        """basis/base.py:39-63: torch.normal(0, 1, (M_k, J)) with a fresh generator seeded `seed` (global if None)."""
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        return torch.normal(mean=0.0, std=1.0, size=(self.approximation_dimension, number_of_particles), generator=gen)

    def forward(self, particles: torch.Tensor) -> torch.Tensor:
        """F = k(X, Z) @ V~ @ P evaluated LEFT TO RIGHT as orthonormal.py:106-108 does -> (N, J)."""
        return self.k_zx.T @ self.scaled_eigenvectors @ particles

    def energy_potential(self, particles: torch.Tensor, cost: torch.Tensor) -> float:
        """orthonormal.py:110-126: mean_j [ c_j + 1/2 sum_m P_mj (diag(1/lambda) P)_mj ]."""
        e = cost + 1 / 2 * torch.multiply(particles, torch.diag(torch.reciprocal(self.eigenvalues)) @ particles).sum(dim=0)
        return e.mean().item()

    def particle_update(
        self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float, noise: Optional[torch.Tensor] = None
    ) -> torch.Tensor:
        """orthonormal.py:128-159, the algebra kept literally (left-to-right products):
        delta = -eta * V~^T @ K_zx @ Dc  - eta * diag(1/lambda) @ P + sqrt(2 eta) * xi.
        `noise=None` draws xi from the global torch generator exactly as the reference does."""
        assert particles.shape[0] == self.approximation_dimension  # basis/base.py:156-158
        if noise is None:
            noise = langevin_noise(particles.shape[0], particles.shape[1])
        return (
            -step_size * self.scaled_eigenvectors.T @ self.k_zx @ cost_derivative
            - step_size * torch.diag(torch.reciprocal(self.eigenvalues)) @ particles
            + math.sqrt(2.0 * step_size) * noise
        )

    # ---- prediction side ("next" rows; pinned by tests/test_basis.py:522-635,754-862) ---------------------------
    def sample_predictive_noise(self, particles: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """orthonormal.py:161-214."""
        gram_x = r_kernel(self.base_kernel, self.r_kernel_samples, x, x, additional_approximation_samples=x)
        k_xz = self.base_kernel(x, self.x_induce)
        off = k_xz @ self.scaled_eigenvectors @ torch.diag(self.eigenvalues)
        cov = torch.concatenate(
            [
                torch.concatenate([torch.diag(self.eigenvalues), off.T], dim=1),
                torch.concatenate([off, gram_x], dim=1),
            ],
            dim=0,
        )
        noise = sample_multivariate_normal(mean=torch.zeros(cov.shape[0]), cov=cov, size=(particles.shape[1],)).T
        if self.additional_predictive_noise_distribution is not None:
            noise = noise + self.additional_predictive_noise_distribution.sample(noise.shape).reshape(noise.shape)
        return noise

    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """orthonormal.py:216-244."""
        k_xz = self.base_kernel(x, self.x_induce)
        if noise is None:
            noise = self.sample_predictive_noise(particles, x)
        m_k = self.approximation_dimension
        return noise[m_k:, :] + (k_xz @ self.scaled_eigenvectors @ (particles - noise[:m_k, :]))


class InducingPointBasisOracle:
    """Particles are function values at the inducing points (basis/inducing_point.py:23-240).  `gpytorch.solve(input, rhs)` /
    `solve(lhs=, input=, rhs=)` is restated as a dense solve with k(Z, Z) (gpytorch uses a Cholesky solve up to M = 800 and CG
    above, so parity is only well-defined for M <= 800)."""

    def __init__(self, base_kernel: Callable, x_induce: torch.Tensor, y_induce: torch.Tensor, x_train: torch.Tensor,
                 r_kernel_samples: Optional[torch.Tensor] = None, additional_predictive_noise_distribution=None,
                 r_kernel_is_base: bool = False):
        self.base_kernel, self.x_induce, self.y_induce, self.x_train = base_kernel, x_induce, y_induce, x_train
        self.r_kernel_samples = x_induce if r_kernel_samples is None else r_kernel_samples
        self.additional_predictive_noise_distribution = additional_predictive_noise_distribution
        self.r_kernel_is_base = r_kernel_is_base  # the reference tests' mock r-kernel is the plain base kernel (mockers/kernel.py:26-43)
        self.gram_induce = self._r(x_induce, x_induce)  # :38-40
        self.base_gram_induce = base_kernel(x_induce, x_induce)  # :41-43
        self.base_gram_induce_train = base_kernel(x_induce, x_train)  # :44-46

    def _r(self, x1, x2, extra=None):
        if self.r_kernel_is_base:
            return self.base_kernel(x1, x2)
        return r_kernel(self.base_kernel, self.r_kernel_samples, x1, x2, additional_approximation_samples=extra)

    @property
    def approximation_dimension(self) -> int:  # :52-58
        return self.x_induce.shape[0]

    def initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        noise = torch.normal(mean=0.0, std=1.0, size=(self.approximation_dimension, number_of_particles), generator=gen)
        return noise if noise_only else (self.y_induce[:, None] + noise)  # :60-80

    def forward(self, particles: torch.Tensor) -> torch.Tensor:
        """k(X, Z) k(Z, Z)^{-1} P  (:82-95)."""
        return self.base_gram_induce_train.T @ torch.linalg.solve(self.base_gram_induce, particles)

    def energy_potential(self, particles: torch.Tensor, cost: torch.Tensor) -> float:
        """mean_j [ c_j + M/2 sum_m (k(Z,Z)^{-1} P)_mj^2 ]  (:97-119)."""
        w = torch.linalg.solve(self.base_gram_induce, particles)
        return (cost + self.approximation_dimension / 2 * torch.square(w).sum(dim=0)).mean().item()

    def particle_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float,
                        noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-eta k(Z,X) Dc - eta M k(Z,Z)^{-1} P + sqrt(2 eta) e,  e ~ N(0, k(Z,Z)) drawn as samplers.py:6-44 does  (:121-150).
        `noise` injects the STANDARD normal draw z (M, J); e = V sqrt(clip(lambda, 0)) z."""
        w = torch.linalg.solve(self.base_gram_induce, particles)
        if noise is None:
            e = sample_multivariate_normal(mean=torch.zeros(particles.shape[0]), cov=self.base_gram_induce, size=(particles.shape[1],)).T
        else:
            lam, vec = torch.linalg.eigh(self.base_gram_induce)
            e = vec @ torch.diag(torch.sqrt(torch.clip(lam, 0, None))) @ noise
        return (-step_size * self.base_gram_induce_train @ cost_derivative - step_size * self.approximation_dimension * w
                + math.sqrt(2.0 * step_size) * e)

    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """G(x) + r(x, Z) r(Z, Z)^{-1} (P - G(Z)) with the r-kernel's extra approximation samples x  (:204-240)."""
        gram_x_induce = self._r(x, self.x_induce, extra=x)
        gram_induce = self._r(self.x_induce, self.x_induce, extra=x)
        m = self.approximation_dimension
        return noise[m:, :] + gram_x_induce @ torch.linalg.solve(gram_induce, particles - noise[:m, :])


# ----------------------------------------------------------------------------------------------------------------
# PLS facade -- src/projected_langevin_sampling/projected_langevin_sampling.py
# ----------------------------------------------------------------------------------------------------------------
class PLSOracle:
    def __init__(self, basis, cost: Cost):
        self.basis = basis
        self.cost = cost

    def calculate_cost(self, particles):  # :75-88
        return self.cost.value(self.basis.forward(particles))

    def calculate_cost_derivative(self, particles, force_autograd: bool = False):  # :90-105
        return self.cost.derivative(self.basis.forward(particles), force_autograd=force_autograd)

    def calculate_particle_update(self, particles, step_size, noise=None):  # :107-123
        dc = self.calculate_cost_derivative(particles)
        return self.basis.particle_update(particles, dc, step_size, noise=noise)

    def calculate_energy_potential(self, particles) -> float:  # :125-138
        assert particles.shape[0] == self.basis.approximation_dimension
        return self.basis.energy_potential(particles, self.calculate_cost(particles))

    def run(self, particles, step_size, number_of_steps, noises: Optional[Sequence[torch.Tensor]] = None, energies: bool = False):
        """The caller's loop (experiments/trainers.py:149-161, README.md:258-265): P += delta each step."""
        p = particles.clone()
        hist = []
        for s in range(number_of_steps):
            p = p + self.calculate_particle_update(p, step_size, noise=None if noises is None else noises[s])
            if energies:
                hist.append(self.calculate_energy_potential(p))
        return (p, hist) if energies else p


# ----------------------------------------------------------------------------------------------------------------
# the caller of the hot path -- experiments/early_stopper.py:4-24, experiments/trainers.py:139-162
# ----------------------------------------------------------------------------------------------------------------
class EarlyStopperOracle:
    """experiments/early_stopper.py:4-24: stop on a non-finite loss, or once the loss has failed to improve on its minimum
    for `patience` of accumulated SIMULATED time (sum of step sizes); an improvement resets the clock."""

    def __init__(self, patience: float = 1e-4):
        self.patience = patience
        self.simulation_time = 0
        self.min_loss = float("inf")

    def should_stop(self, loss: float, step_size: float) -> bool:
        if not np.isfinite(loss):
            return True
        if loss >= self.min_loss:
            self.simulation_time += step_size
            return self.simulation_time >= self.patience
        self.min_loss = loss
        self.simulation_time = 0
        return False


def train_pls_oracle(pls: "PLSOracle", particles: torch.Tensor, number_of_epochs: int, step_size: float,
                     early_stopper_patience: float) -> Tuple[torch.Tensor, list]:
    """experiments/trainers.py:139-162: per epoch P += update(P) (noise from torch's global CPU generator), E = energy(P);
    the early stopper is asked BEFORE the energy is appended, so a stopping epoch's update is kept but its energy is not."""
    energy_potentials = []
    stopper = EarlyStopperOracle(patience=early_stopper_patience)
    for _ in range(number_of_epochs):
        particles = particles + pls.calculate_particle_update(particles, step_size)
        energy = pls.calculate_energy_potential(particles)
        if stopper.should_stop(loss=energy, step_size=step_size):
            break
        energy_potentials.append(energy)
    return particles, energy_potentials


def train_pls_runner_oracle(pls: "PLSOracle", particles: torch.Tensor, simulation_duration: float, maximum_number_of_steps: int,
                            early_stopper_patience: float, number_of_step_searches: int, step_size_upper: float,
                            minimum_change_in_energy_potential: float, seed: int):
    """experiments/runners.py:331-446 with metric_to_optimise="loss" (the final energy potential): log-spaced step sizes from
    step_size_upper down to simulation_duration / maximum_number_of_steps, every run restarted from the same particles under
    set_seed(seed), best = lowest final energy among finite runs, early exit when two consecutive step sizes end within
    `minimum_change_in_energy_potential` (relative) of each other.  Parity unpinned against the reference's own runner
    (its module imports matplotlib / sklearn plotting code that is absent here); restated line by line."""
    best_metric_value, best_lr = float("inf"), None
    history = {}
    step_sizes = np.logspace(np.log10(step_size_upper), np.log10(simulation_duration / maximum_number_of_steps), number_of_step_searches)
    particles_out = particles.detach().clone()
    for i, step_size in enumerate(step_sizes):
        number_of_epochs = int(simulation_duration / step_size)
        set_seed(seed)
        particles_i, energies = train_pls_oracle(pls, particles.detach().clone(), number_of_epochs, step_size, early_stopper_patience)
        if energies and torch.isfinite(particles_i).all():
            history[step_size] = energies
            metric_value = energies[-1]
            if metric_value < best_metric_value:
                best_metric_value, best_lr, particles_out = metric_value, step_size, particles_i.detach().clone()
            if (i > 0 and step_sizes[i - 1] in history
                    and abs(history[step_sizes[i - 1]][-1] - energies[-1]) / history[step_sizes[i - 1]][-1] < minimum_change_in_energy_potential):
                break
    return particles_out, best_lr, len(history[best_lr]), history


# ----------------------------------------------------------------------------------------------------------------
# ConditionalVariance inducing-point selector -- src/inducing_point_selectors/conditional_variance.py:27-120
# ----------------------------------------------------------------------------------------------------------------
def conditional_variance_select(
    x: torch.Tensor,
    m: int,
    kernel: Callable,
    threshold: Optional[float] = 0.0,
    jitter: float = 1e-12,
    argsort_kind: Optional[str] = None,
    return_trace: bool = False,
):
    """Greedy pivoted Cholesky / DPP-MAP selection, restated step by step.

    * permutation from numpy's GLOBAL generator (:60); returned indices refer to the ORIGINAL x (:119)
    * di = diag k(x, x) + jitter (:66-71); first pivot = np.argmax (first occurrence on ties) (:72)
    * per iteration: column = round(k(x, x_j), 20) (:95; NOT a no-op: multiply-rint-divide in float64),
      column[j] += jitter (:96), e = (column - c_j . C[:i]) / sqrt(di[j]) (:97), di -= e^2, clip at 0 (:100-103)
    * next pivot = LAST entry of np.argsort(di) not chosen yet (:106-109) -- `argsort_kind=None` keeps numpy's default
      (unstable) sort exactly as the reference; "stable" makes ties resolve to the highest permuted index, the rule
      the CUDA selector implements ("selector ties")
    * early stop when sum(clip(di, 0)) < threshold (:111-116)."""
    assert m > 1, "Must have at least 2 inducing points"
    n = x.shape[0]
    perm = np.random.permutation(n)
    xp = x[perm, ...]
    indices = np.zeros(m, dtype=int) + n
    di = kernel(xp, xp, diag=True).detach().numpy().astype(np.float64) + jitter
    indices[0] = np.argmax(di)
    ci = np.zeros((m - 1, n))
    min_gap = np.inf
    for i in range(m - 1):
        j = int(indices[i])
        dj = np.sqrt(di[j])
        cj = ci[:i, j]
        col = np.round(np.squeeze(kernel(xp, xp[j : j + 1]).detach().numpy().astype(np.float64)), 20)
        col[j] += jitter
        ei = (col - np.dot(cj, ci[:i])) / dj
        ci[i, :] = ei
        di -= ei**2
        di = np.clip(di, 0, None)
        order = np.argsort(di) if argsort_kind is None else np.argsort(di, kind=argsort_kind)
        for nxt in reversed(order):
            if int(nxt) not in indices[: i + 1]:
                indices[i + 1] = int(nxt)
                break
        if return_trace:
            masked = di.copy()
            masked[indices[: i + 1]] = -np.inf
            top2 = np.sort(masked)[-2:]
            if top2[1] > 0:
                min_gap = min(min_gap, (top2[1] - top2[0]) / top2[1])
        if np.sum(np.clip(di, 0, None)) < threshold:
            break
    indices = indices.astype(int)
    induce = xp[indices]  # raises IndexError if the early stop left the sentinel N in `indices` (:111-118)
    out = (induce, torch.from_numpy(perm[indices]))
    if return_trace:
        return out + ({"perm": perm, "local_indices": indices, "min_top2_rel_gap": float(min_gap), "di": di},)
    return out


# ----------------------------------------------------------------------------------------------------------------
# the reference-faithful CPU step used as bench.py's cpu_baseline / --impl reference
# ----------------------------------------------------------------------------------------------------------------
def reference_style_cpu_step(k_zx: torch.Tensor, vt: torch.Tensor, lam: torch.Tensor, y: torch.Tensor,
                             particles: torch.Tensor, step_size: float, observation_noise: float) -> torch.Tensor:
    """One Langevin step the way the reference executes it on CPU with the dense N x M Gram cached (the favourable
    reading): left-to-right matmuls (orthonormal.py:106-108,151-158), Gaussian closed-form derivative
    (gaussian.py:75-88), eigh(eye(M_k)) + torch.normal per step (samplers.py:27-35), dense diag(1/lambda) @ P."""
    f = k_zx.T @ vt @ particles
    dc = (1 / observation_noise) * (f - y[:, None])
    xi = langevin_noise(particles.shape[0], particles.shape[1]).to(particles.dtype)
    return (
        -step_size * vt.T @ k_zx @ dc
        - step_size * torch.diag(torch.reciprocal(lam)) @ particles
        + math.sqrt(2.0 * step_size) * xi
    )
