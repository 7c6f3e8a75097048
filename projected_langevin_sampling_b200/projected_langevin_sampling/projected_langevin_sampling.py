"""PLS facade (reference: src/projected_langevin_sampling/projected_langevin_sampling.py:7-204).

Same methods and argument meaning as the reference.  `calculate_particle_update` takes the fused CUDA path when the
basis offers one (OrthonormalBasis): forward contraction, cost derivative and back-projection run without ever
materialising the N x J matrices the reference builds in between.  Additive extensions: `noise=` (inject the Langevin
noise for parity runs), `step_` (in-place step, optionally with the on-device Philox noise stream) and `run`."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .basis.base import PLSBasis
from .costs.base import PLSCost


class PLS:
    def __init__(self, basis: PLSBasis, cost: PLSCost, name: Optional[str] = None):
        self.basis = basis
        self.cost = cost
        self.name: str = name if name is not None else "pls"

    @property
    def observation_noise(self) -> Optional[float]:
        return self.cost.observation_noise

    @observation_noise.setter
    def observation_noise(self, value: float):
        self.cost.observation_noise = value

    def initialise_particles(self, number_of_particles: int, noise_only: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        return self.basis.initialise_particles(number_of_particles=number_of_particles, noise_only=noise_only, seed=seed)

    def sample_observation_noise(self, number_of_particles: int, seed: Optional[int] = None) -> torch.Tensor:
        return self.cost.sample_observation_noise(number_of_particles=number_of_particles, seed=seed)

    def sample_predictive_noise(self, particles: torch.Tensor, x: torch.Tensor):
        return self.basis.sample_predictive_noise(particles=particles, x=x)

    def _fused(self) -> bool:
        return hasattr(self.basis, "fused_particle_update") and getattr(self.cost, "native_cost_id", -1) >= 0

    def calculate_cost(self, particles: torch.Tensor) -> torch.Tensor:
        """(J,) cost of each particle (:75-88)."""
        if self._fused():
            p = self.basis._particles(particles)
            return self.basis.engine(p.shape[1]).cost(p, self.cost.native(), self.cost.y_device(p.device))
        f = self.basis.calculate_untransformed_train_prediction_samples(particles=particles)
        return self.cost.calculate_cost(untransformed_train_prediction_samples=f)

    def calculate_cost_derivative(self, particles: torch.Tensor) -> torch.Tensor:
        """(N, J) derivative of the cost w.r.t. the untransformed train predictions (:90-105)."""
        if self._fused():
            p = self.basis._particles(particles)
            return self.basis.engine(p.shape[1]).cost_derivative(p, self.cost.native(), self.cost.y_device(p.device))
        f = self.basis.calculate_untransformed_train_prediction_samples(particles=particles)
        return self.cost.calculate_cost_derivative(untransformed_train_prediction_samples=f)

    def calculate_particle_update(self, particles: torch.Tensor, step_size: float, noise=None) -> torch.Tensor:
        """delta (M_k, J) of one Langevin step (:107-123); the caller applies `particles += delta`.
        noise=None draws xi from torch's global CPU generator exactly as the reference does; a tensor injects it."""
        if self._fused():
            return self.basis.fused_particle_update(particles, self.cost, float(step_size), noise=noise)
        cost_derivative = self.calculate_cost_derivative(particles=particles)
        return self.basis.calculate_particle_update(particles=particles, cost_derivative=cost_derivative, step_size=step_size)

    def step_(self, particles: torch.Tensor, step_size: float, noise=None, philox: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
        """In-place Langevin step: particles += delta, one fused launch sequence, no (M_k, J) temporary.
        philox=(seed, step_index, j_global_offset) uses the device-side noise stream (no host RNG, no H2D copy)."""
        if not self._fused():  # a basis / cost without the fused path: the reference's two-call composition, applied in place
            if philox is not None:
                raise ValueError("philox noise needs the fused CUDA path (OrthonormalBasis / InducingPointBasis with a native cost)")
            particles += self.calculate_particle_update(particles, step_size, noise=noise)
            return particles
        return self.basis.fused_particle_update(particles, self.cost, float(step_size), noise=noise, in_place=True, philox=philox)

    def calculate_energy_potential(self, particles: torch.Tensor) -> float:
        """mean over particles of cost + 1/2 P^T Lambda^{-1} P (:125-138)."""
        assert (
            particles.shape[0] == self.basis.approximation_dimension
        ), f"Particles have shape {particles.shape} but requires ({self.basis.approximation_dimension}, J) dimension."
        cost = self.calculate_cost(particles=particles)
        return self.basis.calculate_energy_potential(particles=particles, cost=cost)

    def run(self, particles: torch.Tensor, step_size: float, number_of_steps: int, seed: Optional[int] = None,
            j_global_offset: int = 0, energy_every: int = 0, cuda_graph: bool = False) -> Tuple[torch.Tensor, List[float]]:
        """The caller's loop (experiments/trainers.py:149-161) run in place on the device.  With `seed` the noise is the
        Philox stream keyed on (seed, step, row, global particle); without it the reference's host stream is replayed.
        cuda_graph=True (Philox noise only) captures ONE step -- its five kernel launches plus the increment of a device
        step counter -- in a CUDA graph and replays it: for small problems (BASELINE configs 1-3) a step is tens of
        microseconds of GPU work and the eager loop is bound by launch + host overhead.  The particles are the same."""
        energies: List[float] = []
        if cuda_graph:
            if seed is None:
                raise ValueError("cuda_graph=True needs the device-side Philox noise (seed=...): the reference's host noise "
                                 "stream cannot be captured")
            return self._run_graph(particles, step_size, number_of_steps, seed, j_global_offset, energy_every)
        for s in range(number_of_steps):
            self.step_(particles, step_size, philox=None if seed is None else (seed, s, j_global_offset))
            if energy_every and (s + 1) % energy_every == 0:
                energies.append(self.calculate_energy_potential(particles))
        return particles, energies

    def _run_graph(self, particles: torch.Tensor, step_size: float, number_of_steps: int, seed: int, j_global_offset: int,
                   energy_every: int) -> Tuple[torch.Tensor, List[float]]:
        from .. import _native as nat

        ctx = nat.context(particles.device)
        energies: List[float] = []
        if number_of_steps <= 0:
            return particles, energies
        self.basis.engine(particles.shape[1])  # workspaces must exist before capture
        counter = torch.zeros(1, dtype=torch.int64, device=particles.device)
        # warm-up on a scratch copy: first-launch work (function attributes, lazy module loading) must not happen in capture
        self.step_(particles.clone(), step_size, philox=(seed, 0, j_global_offset))
        torch.cuda.synchronize(particles.device)
        graph = torch.cuda.CUDAGraph()
        ctx.lib.pls_set_step_counter(ctx.handle, counter.data_ptr())
        try:
            with torch.cuda.graph(graph):
                self.step_(particles, step_size, philox=(seed, 0, j_global_offset))  # step index = 0 + *counter
                ctx.check(ctx.lib.pls_advance_step_counter(ctx.handle, counter.data_ptr(), 1, ctx.stream()))
        finally:
            ctx.lib.pls_set_step_counter(ctx.handle, None)
        # capture does not execute: `particles` and the counter are untouched so far
        for s in range(number_of_steps):
            graph.replay()
            if energy_every and (s + 1) % energy_every == 0:
                energies.append(self.calculate_energy_potential(particles))
        torch.cuda.synchronize(particles.device)
        return particles, energies

    # ---- prediction (:140-204) -------------------------------------------------------------------------------------------
    def predict_samples(self, particles: torch.Tensor, x: torch.Tensor, predictive_noise: Optional[torch.Tensor] = None,
                        observation_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        untransformed_samples = self.predict_untransformed_samples(particles=particles, x=x, noise=predictive_noise)
        return self.cost.predict_samples(untransformed_samples=untransformed_samples, observation_noise=observation_noise)

    def predict_untransformed_samples(self, particles: torch.Tensor, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.basis.predict_untransformed_samples(particles=particles, x=x, noise=noise)

    def predict(self, x: torch.Tensor, particles: torch.Tensor, predictive_noise: Optional[torch.Tensor] = None,
                observation_noise: Optional[torch.Tensor] = None):
        prediction_samples = self.predict_samples(particles=particles, x=x, predictive_noise=predictive_noise,
                                                  observation_noise=observation_noise)
        return self.cost.predict(prediction_samples=prediction_samples)

    def __call__(self, x: torch.Tensor, particles: torch.Tensor, predictive_noise: Optional[torch.Tensor] = None,
                 observation_noise: Optional[torch.Tensor] = None):
        return self.predict(x=x, particles=particles, predictive_noise=predictive_noise, observation_noise=observation_noise)
