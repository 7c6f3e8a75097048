"""OrthonormalBasis on the B200 (reference: src/projected_langevin_sampling/basis/orthonormal.py:9-244).

Particles are coordinates in the eigen-basis of (1/M) k(Z, Z).  Unlike the reference, the M x N cross-Gram is never
held: the training and inducing points are stored once in the library's augmented layout and every product with
k(X, Z) regenerates its tiles inside the CUDA kernels (csrc/pls_gen_gemm.cuh).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from ... import _native as nat
from ... import ops
from ...engine import DEFAULT_DC_BUDGET, LangevinEngine, gram_mode, want_gram_cache
from ...kernels import dense_gram, kernel_spec
from ...samplers import langevin_noise, sample_multivariate_normal
from .base import PLSBasis


class OrthonormalBasis(PLSBasis):
    """
    N training points, M inducing points, M_k kept eigen-directions, J particles, D input dimension.

    Extra keyword arguments (not in the reference, all optional):
      eigendecomposition  (eigenvalues, eigenvectors) of (1/M) k(Z, Z), ascending as torch.linalg.eigh returns them;
                          lets a caller share one decomposition between implementations (eigenvector signs are
                          library-dependent)
      eigh_device         "cpu" (default; LAPACK, deterministic) or "cuda"
      dc_budget_bytes     size cap of the d_2 c row-chunk workspace of the fused step
      gradient_reduce     callable applied in place to the (M, J) gradient before the update (NCCL all-reduce when
                          the training rows are sharded across GPUs)
      gaussian_normal_equations   opt-in: with GaussianCost + identity link, form k(Z,X)k(X,Z)/s and k(Z,X)y/s once and run every
                          step in M x M algebra (2 M^2 J flops instead of 4 N M J); see LangevinEngine._normal_equations
      gram_cache          False (default: Gram tiles regenerated inside the kernels, nothing N x M in memory) / True / "auto":
                          keep k(X, Z) resident in HBM, as the reference does (orthonormal.py:36-41), and stream it;
                          "auto" does so when it fits comfortably (engine.want_gram_cache); "staged": re-form k(X_c, Z) every
                          step for the row chunk in flight into one chunk-sized buffer shared by that chunk's launches
    """

    def __init__(self, kernel, x_induce: torch.Tensor, x_train: torch.Tensor, eigenvalue_threshold: float = 0.0,
                 additional_predictive_noise_distribution: Optional[torch.distributions.Distribution] = None, *,
                 eigendecomposition: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eigh_device: str = "cpu",
                 dc_budget_bytes: int = DEFAULT_DC_BUDGET, gradient_reduce=None, verbose: bool = True,
                 gram_cache=False, gaussian_normal_equations: bool = False):
        super().__init__(additional_predictive_noise_distribution=additional_predictive_noise_distribution)
        self.kernel = kernel
        self._gram_cache_mode, self._gram = gram_cache, None
        self._gaussian_normal_equations = gaussian_normal_equations
        self.ctx = nat.context()
        dev = torch.device("cuda", self.ctx.device_index)
        self.x_induce = ops.as_device_f64(x_induce if x_induce.dim() > 1 else x_induce.unsqueeze(-1), dev)  # (M, D)
        self._x_train = ops.as_device_f64(x_train if x_train.dim() > 1 else x_train.unsqueeze(-1), dev)  # (N, D)
        m, d = self.x_induce.shape
        self._d = d
        self._spec = kernel_spec(kernel.base_kernel, d)
        self._dc_budget = dc_budget_bytes
        self._gradient_reduce = gradient_reduce

        # augmented point sets, centred on the inducing points' mean (orthonormal.py:36-41: the two kernel calls)
        centre = self.x_induce.mean(dim=0).tolist() if self._spec.kernel_id == nat.KERNEL_RBF else [0.0] * d
        inv_ls = self._spec.inv_lengthscale
        self._centre = centre
        self._za = ops.prepare_points(self.ctx, self._spec.kernel_id, self.x_induce, inv_ls, centre, self._spec.log_outputscale)
        za_plain = ops.prepare_points(self.ctx, self._spec.kernel_id, self.x_induce, inv_ls, centre, 0.0)
        self._xa = ops.prepare_points(self.ctx, self._spec.kernel_id, self._x_train, inv_ls, centre, 0.0)
        self.base_gram_induce = ops.gram(self.ctx, self._spec.kernel_id, za_plain, self._za, d)  # k(Z, Z), (M, M)

        if eigendecomposition is None:
            scaled = (1 / m) * self.base_gram_induce  # orthonormal.py:46-48
            eigenvalues, eigenvectors = torch.linalg.eigh(scaled if eigh_device == "cuda" else scaled.cpu())
        else:
            eigenvalues, eigenvectors = eigendecomposition
        eigenvalues = ops.as_device_f64(eigenvalues, dev)
        eigenvectors = ops.as_device_f64(eigenvectors, dev)
        keep = torch.where(eigenvalues > eigenvalue_threshold)[0]  # strict; ascending order kept (orthonormal.py:52-56)
        self.eigenvalues = eigenvalues[keep].contiguous()  # (M_k,)
        self.eigenvectors = eigenvectors[:, keep].contiguous()  # (M, M_k)
        if verbose:
            print(f"Number of eigenvalues kept: {self.eigenvalues.shape[0]} out of {m}")
        # V~ = V / sqrt(M_k lambda): scaled with M_k, not M (orthonormal.py:63-68)
        self.scaled_eigenvectors = torch.multiply(
            torch.reciprocal(torch.sqrt(self.approximation_dimension * self.eigenvalues))[None, :], self.eigenvectors
        ).contiguous()
        self._inv_lambda = torch.reciprocal(self.eigenvalues).contiguous()
        self._engines: Dict[int, LangevinEngine] = {}

    # ---- properties --------------------------------------------------------------------------------------------------
    @property
    def approximation_dimension(self) -> int:
        return self.eigenvalues.shape[0]

    @property
    def x_train(self) -> torch.Tensor:
        return self._x_train

    @property
    def base_gram_induce_train(self) -> torch.Tensor:
        """k(Z, X) as a dense (M, N) matrix.  The fused step never forms it; this is for inspection at small N only."""
        return dense_gram(self.kernel.base_kernel, self.x_induce, self._x_train)

    def engine(self, number_of_particles: int) -> LangevinEngine:
        eng = self._engines.get(number_of_particles)
        if eng is None:
            self._engines.clear()  # one set of workspaces at a time
            if self._gram is None and want_gram_cache(self._gram_cache_mode, self.ctx, self._xa.shape[0], self._za.shape[0], self._xa.device):
                self._gram = ops.gram_cache(self.ctx, self._spec.kernel_id, self._xa, self._za, self._d)  # k(X, Z), once
            eng = LangevinEngine(self.ctx, self._spec.kernel_id, self._d, self._xa, self._za, self.scaled_eigenvectors,
                                 self._inv_lambda, number_of_particles, dc_budget_bytes=self._dc_budget,
                                 gradient_reduce=self._gradient_reduce, gram=self._gram,
                                 gaussian_normal_equations=self._gaussian_normal_equations,
                                 gram_staged=gram_mode(self._gram_cache_mode) == "staged")
            self._engines[number_of_particles] = eng
        return eng

    def _particles(self, particles: torch.Tensor) -> torch.Tensor:
        p = particles if (particles.is_cuda and particles.dtype == torch.float64) else ops.as_device_f64(particles, self.x_induce.device)
        return p if p.stride(-1) == 1 else p.contiguous()

    # ---- reference API -------------------------------------------------------------------------------------------------
    def _initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        if not noise_only:
            raise ValueError("For ONB base, noise_only must be True.")
        return self._initialise_particles_noise(number_of_particles=number_of_particles, seed=seed)

    def calculate_untransformed_train_prediction_samples(self, particles: torch.Tensor) -> torch.Tensor:
        """F = k(X, Z) V~ P  (N, J)  (orthonormal.py:98-108)."""
        p = self._particles(particles)
        return self.engine(p.shape[1]).prediction(p)

    def calculate_energy_potential(self, particles: torch.Tensor, cost: torch.Tensor) -> float:
        """mean_j [ c_j + 1/2 sum_m P_mj^2 / lambda_m ]  (orthonormal.py:110-126; host sync like its `.item()`)."""
        p = self._particles(particles)
        j = p.shape[1]
        partial = ops.as_device_f64(cost, p.device).reshape(1, j)
        per_particle = ops.energy_terms(self.ctx, partial, j, p, self._inv_lambda)
        return per_particle.mean().item()

    def _noise(self, particles: torch.Tensor, noise) -> Tuple[int, Optional[torch.Tensor]]:
        if noise is None:  # the reference's draw: torch.normal((M_k, J)) on the global CPU generator (orthonormal.py:141-145)
            noise = langevin_noise(particles.shape[0], particles.shape[1])
        if isinstance(noise, torch.Tensor):
            xi = ops.as_device_f64(noise, particles.device)
            assert xi.shape == particles.shape, f"noise has shape {tuple(xi.shape)}, particles {tuple(particles.shape)}"
            return nat.NOISE_GIVEN, xi
        if noise is False:
            return nat.NOISE_NONE, None
        raise TypeError("noise must be None, False or a tensor of the particles' shape")

    def _calculate_particle_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, step_size: float,
                                   noise=None) -> torch.Tensor:
        """delta = -eta V~^T k(Z,X) Dc - eta Lambda^{-1} P + sqrt(2 eta) xi  (orthonormal.py:128-159)."""
        p = self._particles(particles)
        dc = cost_derivative if (cost_derivative.is_cuda and cost_derivative.dtype == torch.float64) else ops.as_device_f64(cost_derivative, p.device)
        mode, xi = self._noise(p, noise)
        out = torch.empty_like(p, memory_format=torch.contiguous_format)
        return self.engine(p.shape[1]).backproject_update(p, dc, float(step_size), out, mode, xi)

    def fused_particle_update(self, particles: torch.Tensor, cost, step_size: float, noise=None, in_place: bool = False,
                              philox: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
        """One Langevin step without materialising F or k(X, Z).  `philox=(seed, step, j_global_offset)` switches the
        noise to the on-device Philox stream; otherwise `noise` follows `_calculate_particle_update`."""
        p = self._particles(particles)
        if in_place and p is not particles:
            raise ValueError("in_place needs float64 CUDA particles with unit column stride")
        assert (
            p.shape[0] == self.approximation_dimension
        ), f"Particles have shape {p.shape} but requires ({self.approximation_dimension}, J) dimension."
        eng = self.engine(p.shape[1])
        out = p if in_place else torch.empty_like(p, memory_format=torch.contiguous_format)
        if philox is not None or noise is not None:
            # the noise needs no host work: the whole step is ONE library call (pls_step_f64; pls_grad_f64 + the group
            # all-reduce + pls_project_update_f64 when the training rows are sharded)
            if philox is not None:
                mode, xi = nat.NOISE_PHILOX, None
                seed, step_index, j_off = philox
            else:
                mode, xi = self._noise(p, noise)
                seed = step_index = j_off = 0
            return eng.step(p, float(step_size), cost.native(), cost.y_device(p.device), out, mode, xi=xi, seed=seed,
                            step_index=step_index, j_global_offset=j_off, in_place=in_place)
        # The reference's noise is a host draw (~40 ms for 1024 x 4096 normals).  The gradient launches are asynchronous: enqueue
        # them FIRST, draw on the host generator while the GPU works, then upload the draw and enqueue the update.
        eng.gradient(p, cost.native(), cost.y_device(p.device))
        mode, xi = self._noise(p, None)
        return eng.apply_update(p, float(step_size), out, mode, xi=xi, in_place=in_place)

    def apply_langevin_update(self, eng: LangevinEngine, p: torch.Tensor, step_size: float,
                              philox: Optional[Tuple[int, int, int]] = None) -> None:
        """In-place update from the gradient eng.gradient / eng.energy_and_gradient left in the engine (the second half of a step)."""
        if philox is None:
            xi = ops.as_device_f64(langevin_noise(p.shape[0], p.shape[1]), p.device)  # samplers.py:27-35 via orthonormal.py:141-145
            eng.apply_update(p, step_size, p, nat.NOISE_GIVEN, xi=xi, in_place=True)
        else:
            seed, step, j_off = philox
            eng.apply_update(p, step_size, p, nat.NOISE_PHILOX, seed=seed, step_index=step, j_global_offset=j_off, in_place=True)

    # ---- prediction side (reference: orthonormal.py:161-244) -------------------------------------------------------------
    def sample_predictive_noise(self, particles: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        x = ops.as_device_f64(x if x.dim() > 1 else x.unsqueeze(-1), self.x_induce.device)
        gram_x = self.kernel.forward(x1=x, x2=x, additional_approximation_samples=x)
        base_gram_x_induce = dense_gram(self.kernel.base_kernel, x, self.x_induce)
        off_diagonal_block = base_gram_x_induce @ self.scaled_eigenvectors @ torch.diag(self.eigenvalues)
        noise_covariance = torch.concatenate(
            [
                torch.concatenate([torch.diag(self.eigenvalues), off_diagonal_block.T], dim=1),
                torch.concatenate([off_diagonal_block, gram_x], dim=1),
            ],
            dim=0,
        )
        predictive_noise = sample_multivariate_normal(
            mean=torch.zeros(noise_covariance.shape[0], dtype=torch.float64, device=x.device),
            cov=noise_covariance,
            size=(particles.shape[1],),
        ).T
        if self.additional_predictive_noise_distribution is not None:
            extra = self.additional_predictive_noise_distribution.sample(predictive_noise.shape).reshape(predictive_noise.shape)
            predictive_noise = predictive_noise + extra.to(predictive_noise.device, predictive_noise.dtype)
        return predictive_noise

    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """noise[M_k:] + k(x, Z) V~ (P - noise[:M_k])  (orthonormal.py:216-244).  The contraction runs through the same
        generated-operand kernel as the training forward (pls_gemm_f64 + pls_forward_f64): k(x, Z) is not materialised."""
        p = self._particles(particles)
        x = ops.as_device_f64(x if x.dim() > 1 else x.unsqueeze(-1), p.device)
        if noise is None:
            noise = self.sample_predictive_noise(particles=p, x=x)
        noise = ops.as_device_f64(noise, p.device)
        m_k = self.approximation_dimension
        j = p.shape[1]
        xa = ops.prepare_points(self.ctx, self._spec.kernel_id, x, self._spec.inv_lengthscale, self._centre, 0.0)
        w, _ = ops.alloc_matrix(self.x_induce.shape[0], j, p.device)
        ops.gemm(self.ctx, self.scaled_eigenvectors, (p - noise[:m_k, :]).contiguous(), w)  # W = V~ (P - noise_top)
        out, _ = ops.alloc_matrix(x.shape[0], j, p.device)
        ops.forward(self.ctx, self._spec.kernel_id, xa, self._za, self._d, w, j, nat.EPI_PREDICTION, out)
        return noise[m_k:, :] + out[:, :j]
