"""`train_pls`, the caller of the Langevin step (reference: experiments/trainers.py:139-162), on the device.

The reference runs, per epoch, `PLS.calculate_particle_update` and then `PLS.calculate_energy_potential` -- two forward
contractions F = k(X, Z) V~ P over the training set.  Here the energy of the particles entering an epoch comes out of
the SAME forward pass that produces that epoch's cost derivative (pls_forward_step_f64), so an epoch costs one forward +
one backward; only the energy after the last epoch needs a forward of its own.  The results are the reference's:
the energy list is shifted back by one epoch internally, the EarlyStopper sees the same sequence of (energy, step_size),
and an epoch whose energy stops the run keeps its update and drops its energy exactly as trainers.py:157-161 does.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .early_stopper import EarlyStopper
from .projected_langevin_sampling import PLS


def train_pls(pls: PLS, particles: torch.Tensor, number_of_epochs: int, step_size: float, early_stopper_patience: float,
              tqdm_desc: Optional[str] = None, philox_seed: Optional[int] = None,
              j_global_offset: int = 0) -> Tuple[torch.Tensor, List[float]]:
    """Returns (particles, energy_potentials) like the reference.  `particles` is updated in place when it is a float64
    CUDA tensor (the reference's `particles += particle_update`); otherwise the result is copied back into it.
    philox_seed=None replays the reference's noise (one torch.normal((M_k, J)) per epoch on the global CPU generator);
    an integer switches to the on-device Philox stream keyed on (seed, epoch, row, j_global_offset + column).
    Both bases run the fused epoch: for the InducingPointBasis the prior term of the energy is taken on the
    W = k(Z, Z)^{-1} P that the gradient has just formed (inducing_point.py:97-119).
    `tqdm_desc` is accepted for signature compatibility (no progress bar is drawn)."""
    del tqdm_desc
    basis, cost = pls.basis, pls.cost
    if not pls._fused():
        raise TypeError("train_pls needs a basis and a cost with a CUDA implementation (there is no CPU fallback)")
    p = basis._particles(particles)  # float64, on the device, unit column stride (a copy if `particles` is not)
    assert (
        p.shape[0] == basis.approximation_dimension
    ), f"Particles have shape {p.shape} but requires ({basis.approximation_dimension}, J) dimension."
    step_size = float(step_size)
    eng = basis.engine(p.shape[1])
    native_cost, y = cost.native(), cost.y_device(p.device)
    energy_potentials: List[float] = []
    early_stopper = EarlyStopper(patience=early_stopper_patience)
    stopped = False
    for epoch in range(number_of_epochs):
        # energy of the particles as they are now (= after the previous epoch's update) + gradient for this epoch
        per_particle = eng.energy_and_gradient(p, native_cost, y)
        if epoch > 0:
            energy = per_particle.mean().item()  # the reference's `.mean().item()` host sync (orthonormal.py:124-126)
            if early_stopper.should_stop(loss=energy, step_size=step_size):
                stopped = True  # the previous epoch's update stays, its energy is not recorded (trainers.py:159-161)
                break
            energy_potentials.append(energy)
        # both bases: the update from the gradient (and, InducingPointBasis, from W = k(Z, Z)^{-1} P) that pass left in the engine;
        # the reference's torch.normal((M, J)) draw per epoch, or the device-side Philox stream
        basis.apply_langevin_update(eng, p, step_size, None if philox_seed is None else (philox_seed, epoch, j_global_offset))
    if not stopped and number_of_epochs > 0:
        energy = pls.calculate_energy_potential(p)
        if not early_stopper.should_stop(loss=energy, step_size=step_size):
            energy_potentials.append(energy)
    if p is not particles:
        particles.copy_(p.to(device=particles.device, dtype=particles.dtype))
    return particles, energy_potentials
