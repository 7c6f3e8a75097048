"""The CPU arm of bench.py: the reference's own Langevin step on the box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE (only bench.py's `cpu_baseline` / `--impl reference` legs import it).

Two implementations of the same step, chosen by what is on the machine:

* kind "reference+stub" -- the UNMODIFIED reference installed under oracle/_ref (oracle/install_reference.py) driven through
  its own public API: `OrthonormalBasis(kernel, x_induce, x_train)`, `GaussianCost / BernoulliCost / PoissonCost`, `PLS`,
  `particles += pls.calculate_particle_update(particles, step_size)` (README.md:197-215; src/projected_langevin_sampling/
  projected_langevin_sampling.py:107-123).  gpytorch is not installed, so the reference runs behind oracle/gpytorch_stub, whose
  kernel call returns the DENSE Gram: k(Z, X) is evaluated once in the constructor and kept (favourable to the reference:
  gpytorch's lazy tensor re-evaluates the kernel when it is transposed).
* kind "port" -- oracle.pls_oracle's restatement of that step (same algebra, same order), when oracle/_ref is absent.

Sizing: FULL N and M, a chunk of J_c particles per call (the N x J matrices of all J = 4096 particles
do not fit host memory next to their temporaries: 3 x 32.8 GB + the Gram).  One timed "step" = one call on J_c particles.
The step has a J-independent part (k(X,Z) V~ and V~^T k(Z,X), re-formed by the reference on every call, orthonormal.py:
106-108,151-154; eigh(eye(M_k)) of the noise sampler, samplers.py:27) and a part linear in J; both are MEASURED at full N
(calls with J_c and 2 J_c particles) and the whole-J step is composed as t_fixed + (J / J_c) t_linear -- i.e. the
J-independent work is counted once per step, as the reference would do it with all particles in one call.
"""
from __future__ import annotations

import math
import os
import statistics
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "src", "projected_langevin_sampling", "projected_langevin_sampling.py"))


def _build_reference_pls(workload, x, y, z, ls, outputscale):
    """The unmodified reference's objects behind the gpytorch stub."""
    for path in (os.path.join(HERE, "_ref"), os.path.join(HERE, "gpytorch_stub"), ROOT):
        if path not in sys.path:
            sys.path.insert(0, path)
    import gpytorch  # the stub
    from gpytorch import kernels as stub_kernels

    stub_kernels.LAZY_ABOVE = float("inf")  # dense Gram kept, never re-evaluated (favourable to the reference)
    from src.projected_langevin_sampling import PLS, PLSKernel
    from src.projected_langevin_sampling.basis import OrthonormalBasis
    from src.projected_langevin_sampling.costs import BernoulliCost, GaussianCost, PoissonCost
    from src.projected_langevin_sampling.link_functions import IdentityLinkFunction, SigmoidLinkFunction, SquareLinkFunction

    kernel = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=x.shape[1]))
    kernel.base_kernel.lengthscale = ls.reshape(1, -1).clone()
    kernel.outputscale = outputscale
    basis = OrthonormalBasis(kernel=PLSKernel(base_kernel=kernel, approximation_samples=z), x_induce=z, x_train=x)
    if workload["cost"] == "gaussian":
        cost = GaussianCost(observation_noise=0.01, y_train=y, link_function=IdentityLinkFunction())
    elif workload["cost"] == "bernoulli":
        cost = BernoulliCost(y_train=y, link_function=SigmoidLinkFunction())
    else:
        cost = PoissonCost(y_train=y, link_function=SquareLinkFunction())
    pls = PLS(basis=basis, cost=cost)
    return pls, basis.approximation_dimension


class _PortPLS:
    """The oracle's restatement of the same step (kind "port")."""

    def __init__(self, workload, x, y, z, ls, outputscale):
        from oracle.pls_oracle import OrthonormalBasisOracle, RBFScaleKernel

        basis = OrthonormalBasisOracle(RBFScaleKernel(ls, outputscale), z, x)
        self.k_zx, self.vt, self.lam = basis.k_zx.contiguous(), basis.scaled_eigenvectors, basis.eigenvalues
        self.y, self.cost = y, workload["cost"]

    def calculate_particle_update(self, particles, step_size):
        from oracle.pls_oracle import langevin_noise

        f = self.k_zx.T @ self.vt @ particles  # orthonormal.py:106-108, left to right
        if self.cost == "gaussian":
            dc = (1 / 0.01) * (f - self.y[:, None])
        elif self.cost == "bernoulli":
            pr = torch.clip(torch.reciprocal(1 + torch.exp(-f)), 1e-10, 1 - 1e-10)
            dc = -self.y[:, None] * (1 - pr) + (1 - self.y[:, None]) * pr
        else:
            dc = -2 * self.y[:, None] / f + 2 * f
        xi = langevin_noise(particles.shape[0], particles.shape[1])  # eigh(eye(M_k)) + torch.normal, samplers.py:27-35
        return (-step_size * self.vt.T @ self.k_zx @ dc - step_size * torch.diag(torch.reciprocal(self.lam)) @ particles
                + math.sqrt(2 * step_size) * xi)


def run(workload: dict, inputs, steps: int, warmup: int, eta: float, j_chunk: int = 256, force_port: bool = False) -> dict:
    """Times `steps` reference calls on J_c particles at full N and M (after `warmup` untimed ones) and two more on 2 J_c
    particles for the J-linear part.  Returns the composed whole-J figure and everything it was composed from."""
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_default_dtype(torch.float64)  # README.md:86-87
    x, y, z, ls, outputscale = inputs
    n, m, j_full = x.shape[0], z.shape[0], workload["j"]
    j_c = min(j_chunk, j_full)
    t0 = time.perf_counter()
    if reference_available() and not force_port:
        kind = "reference+stub"
        pls, m_k = _build_reference_pls(workload, x, y, z, ls, outputscale)
    else:
        kind = "port"
        pls = _PortPLS(workload, x, y, z, ls, outputscale)
        m_k = pls.vt.shape[1]
    setup_s = time.perf_counter() - t0

    def timed_calls(j, count, skip):
        p = torch.randn(m_k, j, generator=torch.Generator().manual_seed(1))
        out = []
        for it in range(skip + count):
            t = time.perf_counter()
            p += pls.calculate_particle_update(particles=p, step_size=eta)
            if it >= skip:
                out.append(time.perf_counter() - t)
        return out

    t_c = timed_calls(j_c, steps, warmup)
    two = min(2 * j_c, j_full)
    # (one untimed call first: the first call at a new particle count first-touches ~4 GB of fresh N x 2 J_c temporaries)
    t_2c = timed_calls(two, max(2, min(3, steps)), 1) if two > j_c else list(t_c)
    # the fastest call of each kind: the 2 J_c calls first-touch ~4 GB of fresh temporaries and scatter by +-15 % from run to run; the
    # minimum is the undisturbed time and the choice favourable to the reference (medians are reported beside it)
    t1, t2 = min(t_c), min(t_2c)
    t_linear = max(t2 - t1, 0.0) * (j_c / (two - j_c)) if two > j_c else t1  # seconds per J_c particles
    t_fixed = max(t1 - t_linear, 0.0)
    t_full = t_fixed + (j_full / j_c) * t_linear
    te = time.perf_counter()
    torch.linalg.eigh(torch.eye(m_k))
    eigh_s = time.perf_counter() - te
    cores = torch.get_num_threads()
    return {
        "value": j_full / t_full, "unit": "particle-updates/s", "cores": cores, "kind": kind,
        "ms_per_step_whole_j_composed": t_full * 1e3,
        "measured": {"rows": n, "m": m, "m_k": m_k, "j_chunk": j_c, "steps": len(t_c), "ms_per_call_median": statistics.median(t_c) * 1e3,
                     "ms_per_call_min": t1 * 1e3,
                     "ms_per_call_all": [round(v * 1e3, 1) for v in t_c], "particle_updates_per_s_of_the_chunk_alone": j_c / t1,
                     "j_chunk_2": two, "ms_per_call_2_median": statistics.median(t_2c) * 1e3, "ms_per_call_2_min": t2 * 1e3,
                     "ms_per_call_2_all": [round(v * 1e3, 1) for v in t_2c],
                     "ms_j_independent": t_fixed * 1e3, "ms_per_j_chunk_linear": t_linear * 1e3,
                     "ms_eigh_identity": eigh_s * 1e3, "setup_s": setup_s},
        "sample": (f"{kind}: the reference's PLS.calculate_particle_update on {cores} host threads at FULL N={n}, M={m} (M_k={m_k}), "
                   f"{j_c} particles per call ({len(t_c)} timed calls, fastest {t1:.3f} s; {two} particles: fastest of {len(t_2c)} {t2:.3f} s) -> J-independent "
                   f"{t_fixed:.3f} s (of which eigh(eye(M_k)) {eigh_s:.3f} s) + {t_linear:.3f} s per {j_c} particles; whole J={j_full} "
                   f"composed with the J-independent work counted once = {t_full:.2f} s/step; dense Gram kept in memory"),
        "sample_seconds": sum(t_c) + sum(t_2c),
    }
