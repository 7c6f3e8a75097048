"""InducingPointBasis (reference: src/projected_langevin_sampling/basis/inducing_point.py:23-240): the particles are function
values at the inducing points, F = k(X, Z) k(Z, Z)^{-1} P.

The two N-sized contractions are the same generated-operand kernels as the OrthonormalBasis uses (pls_forward_f64 with
W = k(Z, Z)^{-1} P, pls_backward_f64 for k(Z, X) Dc): k(X, Z) and the (N, J) prediction are never materialised in the fused
step.  The M x M pieces follow the reference: `gpytorch.solve(k(Z, Z), .)` is a Cholesky solve (gpytorch's choice up to
M = 800; above that it switches to CG with a loose tolerance, so parity is only well-defined for M <= 800) with the factor
computed once on the host in float64 (`psd_safe_cholesky`: the same jitter retries when k(Z, Z) is numerically indefinite;
the per-step solve stays `torch.cholesky_solve` = cuSOLVER potrs, 2 M^2 J flops of library TRSM against the step's 4 N M J --
for M <= 800 that is < 0.1 % of a step at any N >= 100 M, so a hand-written triangular solve would buy nothing measurable and
would have to reproduce potrs's rounding to keep the 1e-10 parity with the reference's own Cholesky solve); the N(0, k(Z, Z)) noise is V sqrt(clip(lambda, 0)) z with the eigendecomposition
computed once (the reference recomputes it every step, samplers.py:6-44) and z the reference's torch.normal((M, J)) draw.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from ... import _native as nat
from ... import ops
from ...engine import DEFAULT_DC_BUDGET, LangevinEngine, gram_mode, want_gram_cache
from ...kernels import dense_gram, kernel_spec
from ...samplers import langevin_noise, sample_multivariate_normal
from .base import PLSBasis


def psd_safe_cholesky(a: torch.Tensor, max_tries: int = 3) -> torch.Tensor:
    """What `gpytorch.solve` factorises with (linear_operator.utils.cholesky.psd_safe_cholesky, the path taken by
    basis/inducing_point.py:104-106,141-143 for M <= 800): plain Cholesky first; if the matrix is numerically indefinite, retry
    with a growing diagonal jitter -- 1e-8 in float64, times 10 per retry (gpytorch's cholesky_jitter default for double,
    cholesky_max_tries = 3) -- and raise only when the last retry fails.  A warning reports the jitter that was needed."""
    import warnings

    l, info = torch.linalg.cholesky_ex(a)
    if not bool(info.any()):
        return l
    if bool(torch.isnan(a).any()):
        raise ValueError(f"cholesky: {int(torch.isnan(a).sum())} of {a.numel()} elements of the {tuple(a.shape)} matrix are NaN")
    base, prev, a_j = 1e-8 if a.dtype == torch.float64 else 1e-6, 0.0, a.clone()
    for i in range(max_tries):
        jitter = base * (10**i)
        a_j.diagonal().add_(jitter - prev)
        prev = jitter
        l, info = torch.linalg.cholesky_ex(a_j)
        if not bool(info.any()):
            warnings.warn(f"k(Z, Z) not positive definite: added jitter of {jitter:.1e} to the diagonal", RuntimeWarning)
            return l
    raise torch.linalg.LinAlgError(f"k(Z, Z) not positive definite after adding {prev:.1e} to the diagonal")


class InducingPointBasis(PLSBasis):
    def __init__(self, kernel, x_induce: torch.Tensor, y_induce: torch.Tensor, x_train: torch.Tensor,
                 additional_predictive_noise_distribution: Optional[torch.distributions.Distribution] = None, *,
                 dc_budget_bytes: int = DEFAULT_DC_BUDGET, gradient_reduce=None, gram_cache=False):
        super().__init__(additional_predictive_noise_distribution=additional_predictive_noise_distribution)
        self.kernel = kernel
        self._gram_cache_mode, self._gram = gram_cache, None  # as OrthonormalBasis
        self.ctx = nat.context()
        dev = torch.device("cuda", self.ctx.device_index)
        self.x_induce = ops.as_device_f64(x_induce if x_induce.dim() > 1 else x_induce.unsqueeze(-1), dev)  # (M, D)
        self.y_induce = ops.as_device_f64(y_induce.reshape(-1), dev)  # (M,)
        self._x_train = ops.as_device_f64(x_train if x_train.dim() > 1 else x_train.unsqueeze(-1), dev)  # (N, D)
        m, d = self.x_induce.shape
        self._d = d
        self._spec = kernel_spec(kernel.base_kernel, d)
        self._dc_budget, self._gradient_reduce = dc_budget_bytes, gradient_reduce
        centre = self.x_induce.mean(dim=0).tolist() if self._spec.kernel_id == nat.KERNEL_RBF else [0.0] * d
        self._centre = centre
        inv_ls = self._spec.inv_lengthscale
        self._za = ops.prepare_points(self.ctx, self._spec.kernel_id, self.x_induce, inv_ls, centre, self._spec.log_outputscale)
        za_plain = ops.prepare_points(self.ctx, self._spec.kernel_id, self.x_induce, inv_ls, centre, 0.0)
        self._xa = ops.prepare_points(self.ctx, self._spec.kernel_id, self._x_train, inv_ls, centre, 0.0)
        self.gram_induce = kernel.forward(x1=self.x_induce, x2=self.x_induce)  # r(Z, Z)   (:38-40)
        self.base_gram_induce = ops.gram(self.ctx, self._spec.kernel_id, za_plain, self._za, d)  # k(Z, Z)   (:41-43)
        k_host = self.base_gram_induce.cpu()
        self._chol = psd_safe_cholesky(k_host).to(dev)  # gpytorch.solve's Cholesky path (with its jitter retries)
        lam, vec = torch.linalg.eigh(k_host)  # samplers.py:23-26, once
        self._noise_factor = (vec * torch.sqrt(torch.clip(lam, 0, None))[None, :]).to(dev).contiguous()  # V sqrt(lambda)
        self._m_over = torch.full((m,), float(m), dtype=torch.float64, device=dev)
        self._engines: Dict[int, LangevinEngine] = {}

    # ---- properties ------------------------------------------------------------------------------------------------------
    @property
    def approximation_dimension(self) -> int:
        return self.x_induce.shape[0]

    @property
    def x_train(self) -> torch.Tensor:
        return self._x_train

    @property
    def base_gram_induce_train(self) -> torch.Tensor:
        """k(Z, X) (M, N), materialised on request only (the reference holds it, :44-46; the CUDA path never needs it)."""
        return dense_gram(self.kernel.base_kernel, self.x_induce, self._x_train)

    def _solve(self, rhs: torch.Tensor) -> torch.Tensor:
        return torch.cholesky_solve(rhs, self._chol)

    def engine(self, number_of_particles: int) -> LangevinEngine:
        eng = self._engines.get(number_of_particles)
        if eng is None:
            self._engines.clear()

            def weights(particles: torch.Tensor, w: torch.Tensor) -> None:
                w[:, : particles.shape[1]].copy_(self._solve(particles))  # W = k(Z, Z)^{-1} P

            if self._gram is None and want_gram_cache(self._gram_cache_mode, self.ctx, self._xa.shape[0], self._za.shape[0], self._xa.device):
                self._gram = ops.gram_cache(self.ctx, self._spec.kernel_id, self._xa, self._za, self._d)
            eye = torch.empty((self.approximation_dimension, 0), dtype=torch.float64, device=self.x_induce.device)
            eng = LangevinEngine(self.ctx, self._spec.kernel_id, self._d, self._xa, self._za, eye, self._m_over, number_of_particles,
                                 dc_budget_bytes=self._dc_budget, gradient_reduce=self._gradient_reduce, weights_fn=weights,
                                 gram=self._gram, gram_staged=gram_mode(self._gram_cache_mode) == "staged")
            self._engines[number_of_particles] = eng
        return eng

    def _particles(self, particles: torch.Tensor) -> torch.Tensor:
        p = particles if (particles.is_cuda and particles.dtype == torch.float64) else ops.as_device_f64(particles, self.x_induce.device)
        return p if p.stride(1) == 1 else p.contiguous()

    # ---- reference API ----------------------------------------------------------------------------------------------------
    def _initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        noise = self._initialise_particles_noise(number_of_particles=number_of_particles, seed=seed)
        return noise if noise_only else (self.y_induce.cpu().to(noise.dtype)[:, None] + noise)  # :60-80

    def calculate_untransformed_train_prediction_samples(self, particles: torch.Tensor) -> torch.Tensor:
        """k(X, Z) k(Z, Z)^{-1} P  (N, J)  (:82-95)."""
        p = self._particles(particles)
        return self.engine(p.shape[1]).prediction(p)

    def calculate_energy_potential(self, particles: torch.Tensor, cost: torch.Tensor) -> float:
        """mean_j [ c_j + M/2 sum_m (k(Z,Z)^{-1} P)_mj^2 ]  (:97-119)."""
        p = self._particles(particles)
        j = p.shape[1]
        w = self._solve(p).contiguous()
        partial = ops.as_device_f64(cost, p.device).reshape(1, j)
        return ops.energy_terms(self.ctx, partial, j, w, self._m_over).mean().item()  # 1/2 sum w^2 * M

    def _noise(self, particles: torch.Tensor, noise, philox: Optional[Tuple[int, int, int]] = None) -> Optional[torch.Tensor]:
        """e ~ N(0, k(Z, Z)) (M, J).  noise=None: the reference's draw (one torch.normal((M, J)) on the global CPU generator,
        samplers.py:27-35); a tensor: the STANDARD normal z to colour; False: no noise; philox=(seed, step, j_global_offset):
        z from the device-side Philox4x32-10 stream keyed on the global (row, particle), as the OrthonormalBasis uses it."""
        if noise is False:
            return None
        if philox is not None:
            seed, step, j_off = philox
            z = ops.philox_normal(self.ctx, seed, step, particles.shape[0], particles.shape[1], j_off, device=particles.device)
        else:
            z = langevin_noise(particles.shape[0], particles.shape[1]) if noise is None else noise
        z = ops.as_device_f64(z, particles.device)
        e, _ = ops.alloc_matrix(particles.shape[0], particles.shape[1], particles.device)
        ops.gemm(self.ctx, self._noise_factor, z, e)  # V sqrt(lambda) z
        return e[:, : particles.shape[1]]

    def _combine(self, particles: torch.Tensor, gm: torch.Tensor, step_size: float, noise, in_place: bool,
                 philox: Optional[Tuple[int, int, int]] = None, w: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-eta G' - eta M k(Z,Z)^{-1} P + sqrt(2 eta) e   (:140-149).  w: k(Z,Z)^{-1} P when the caller already holds it (the
        gradient leaves it in the engine's W workspace)."""
        j = particles.shape[1]
        if w is None:
            w = self._solve(particles).contiguous()
        e = self._noise(particles, noise, philox)
        out = particles if in_place else torch.empty_like(particles, memory_format=torch.contiguous_format)
        zero = e if e is not None else w
        return ops.lincomb3(self.ctx, -step_size, gm, -step_size * self.approximation_dimension, w,
                            math.sqrt(2.0 * step_size) if e is not None else 0.0, zero, j, out, base=particles if in_place else None)

    def _calculate_particle_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float, noise=None) -> torch.Tensor:
        p = self._particles(particles)
        eng = self.engine(p.shape[1])
        dc = cost_derivative if (cost_derivative.is_cuda and cost_derivative.dtype == torch.float64) else ops.as_device_f64(cost_derivative, p.device)
        if dc.stride(1) != 1 or (dc.stride(0) & 1) or (dc.data_ptr() & 15):
            buf, _ = ops.alloc_matrix(dc.shape[0], dc.shape[1], dc.device)
            buf[:, : dc.shape[1]].copy_(dc)
            dc = buf[:, : dc.shape[1]]
        splits = ops.backward_splits(self.ctx, eng.n, eng.m, eng.j)
        gp = torch.empty((splits, eng.m, eng.ldj), dtype=torch.float64, device=dc.device)
        ops.backward(self.ctx, self._spec.kernel_id, self._za, self._xa, self._d, dc, eng.j, gp, splits, accumulate=False)
        ops.reduce_splits(self.ctx, gp, eng.j, eng.gm)
        if self._gradient_reduce is not None:
            self._gradient_reduce(eng.gm)
        return self._combine(p, eng.gm, float(step_size), noise, in_place=False)

    def fused_particle_update(self, particles: torch.Tensor, cost, step_size: float, noise=None, in_place: bool = False,
                              philox: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
        """One Langevin step without materialising k(X, Z) or the (N, J) prediction."""
        p = self._particles(particles)
        if in_place and p is not particles:
            raise ValueError("in_place needs float64 CUDA particles with unit column stride")
        assert (
            p.shape[0] == self.approximation_dimension
        ), f"Particles have shape {p.shape} but requires ({self.approximation_dimension}, J) dimension."
        eng = self.engine(p.shape[1])
        gm = eng.gradient(p, cost.native(), cost.y_device(p.device))  # enqueued first: the host noise draw overlaps it
        return self._combine(p, gm, float(step_size), noise, in_place=in_place, philox=philox, w=eng.w[:, : p.shape[1]])

    def apply_langevin_update(self, eng: LangevinEngine, p: torch.Tensor, step_size: float,
                              philox: Optional[Tuple[int, int, int]] = None) -> None:
        """In-place update from the gradient and W = k(Z, Z)^{-1} P that eng.gradient / eng.energy_and_gradient left behind."""
        self._combine(p, eng.gm, step_size, None, in_place=True, philox=philox, w=eng.w[:, : p.shape[1]])

    # ---- prediction side (:152-240) -----------------------------------------------------------------------------------------
    def sample_predictive_noise(self, particles: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        x = ops.as_device_f64(x if x.dim() > 1 else x.unsqueeze(-1), self.x_induce.device)
        gram_x = self.kernel.forward(x1=x, x2=x, additional_approximation_samples=x)
        gram_induce_x = self.kernel.forward(x1=self.x_induce, x2=x, additional_approximation_samples=x)
        noise_covariance = torch.concatenate(
            [torch.concatenate([self.gram_induce, gram_induce_x], dim=1), torch.concatenate([gram_induce_x.T, gram_x], dim=1)], dim=0)
        predictive_noise = sample_multivariate_normal(
            mean=torch.zeros(noise_covariance.shape[0], dtype=torch.float64, device=x.device), cov=noise_covariance,
            size=(particles.shape[1],)).T
        if self.additional_predictive_noise_distribution is not None:
            extra = self.additional_predictive_noise_distribution.sample(predictive_noise.shape).reshape(predictive_noise.shape)
            predictive_noise = predictive_noise + extra.to(predictive_noise.device, predictive_noise.dtype)
        return predictive_noise

    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """G(x) + r(x, Z) r(Z, Z)^{-1} (P - G(Z))  (:204-240); r = the PLSKernel with x as extra approximation samples."""
        p = self._particles(particles)
        x = ops.as_device_f64(x if x.dim() > 1 else x.unsqueeze(-1), p.device)
        gram_x_induce = self.kernel.forward(x1=x, x2=self.x_induce, additional_approximation_samples=x)
        gram_induce = self.kernel.forward(x1=self.x_induce, x2=self.x_induce, additional_approximation_samples=x)
        if noise is None:
            noise = self.sample_predictive_noise(particles=p, x=x)
        noise = ops.as_device_f64(noise, p.device)
        m = self.approximation_dimension
        return noise[m:, :] + gram_x_induce @ torch.linalg.solve(gram_induce, p - noise[:m, :])
