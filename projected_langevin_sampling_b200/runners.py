"""Step-size search around `train_pls` and the checkpoint format of the reference's experiments
(reference: experiments/runners.py:331-446 `train_pls_runner`, experiments/uci/regression/main.py:300-308,
experiments/loaders.py:10-28 `load_pls`).

The search re-runs the Langevin loop over log-spaced step sizes from the SAME initial particles and seed, keeps the
particles of the best run and stops early once two consecutive step sizes end at (relatively) the same energy.  The
reference scores a run with metrics computed from `pls.predict` on the training data (nll / mse / mae / acc / auc / f1,
experiments/metrics.py -- out of scope here) or with the final energy potential ("loss"); this module takes the score as
"loss" or as a callable `metric(pls, particles) -> float` plus its direction, so the experiment layer can plug its own.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .projected_langevin_sampling import PLS
from .trainers import train_pls
from .utils import set_seed

Metric = Union[str, Callable[[PLS, torch.Tensor], float]]


def train_pls_runner(pls: PLS, particles: torch.Tensor, simulation_duration: float, maximum_number_of_steps: int,
                     early_stopper_patience: float, number_of_step_searches: int, step_size_upper: float,
                     minimum_change_in_energy_potential: float, seed: int, metric_to_optimise: Metric = "loss",
                     higher_is_better: bool = False, particle_name: str = "pls",
                     energy_potentials_history: Optional[Dict[float, List[float]]] = None) -> Tuple[torch.Tensor, float, int]:
    """Returns (best particles, best step size, number of accepted epochs of the best run), as runners.py:331-446.
    `energy_potentials_history` (optional dict) receives {step_size: energies} of every finite run (the reference plots it)."""
    if isinstance(metric_to_optimise, str) and metric_to_optimise != "loss":
        raise NotImplementedError(f"Unknown metric to optimise {metric_to_optimise}: pass 'loss' or a callable(pls, particles).")
    best_metric_value = 0 if higher_is_better else float("inf")  # runners.py:349-355
    best_lr = None
    history: Dict[float, List[float]] = {} if energy_potentials_history is None else energy_potentials_history
    step_sizes = np.logspace(np.log10(step_size_upper), np.log10(simulation_duration / maximum_number_of_steps),
                             number_of_step_searches)  # :358-362
    particles_out = particles.detach().clone()
    for i, step_size in enumerate(step_sizes):
        number_of_epochs = int(simulation_duration / step_size)
        set_seed(seed)  # identical noise stream for every step size (:366)
        particles_i, energy_potentials = train_pls(
            pls=pls, particles=particles.detach().clone(), number_of_epochs=number_of_epochs, step_size=step_size,
            early_stopper_patience=early_stopper_patience,
            tqdm_desc=f"PLS Step Size Search {i + 1} of {number_of_step_searches} for {particle_name} ({step_size=})")
        if energy_potentials and bool(torch.isfinite(particles_i).all()):
            history[step_size] = energy_potentials
            metric_value = energy_potentials[-1] if isinstance(metric_to_optimise, str) else float(metric_to_optimise(pls, particles_i))
            if (metric_value > best_metric_value) if higher_is_better else (metric_value < best_metric_value):
                best_metric_value = metric_value
                best_lr = step_size
                particles_out = particles_i.detach().clone()
            if (i > 0 and step_sizes[i - 1] in history
                    and abs(history[step_sizes[i - 1]][-1] - energy_potentials[-1]) / history[step_sizes[i - 1]][-1]
                    < minimum_change_in_energy_potential):
                break  # :423-433
    return particles_out, best_lr, len(history[best_lr])  # KeyError when no run was finite, like the reference


def save_pls(pls: PLS, particles: torch.Tensor, model_path: str, best_lr: Optional[float] = None,
             number_of_epochs: Optional[int] = None) -> None:
    """The `.pth` dictionary the reference's experiments write (experiments/uci/regression/main.py:300-308)."""
    torch.save({"particles": particles, "observation_noise": pls.observation_noise,
                "best_lr": None if best_lr is None else float(best_lr),  # the search yields numpy scalars
                "number_of_epochs": None if number_of_epochs is None else int(number_of_epochs)}, model_path)


def load_pls(pls: PLS, model_path: str) -> Tuple[PLS, torch.Tensor, Optional[float], Optional[int]]:
    """experiments/loaders.py:10-28: restores the particles and the observation noise into `pls`."""
    # weights_only=False: checkpoints written by the reference hold numpy scalars (best_lr), which the restricted unpickler of
    # torch >= 2.6 rejects; these are the user's own local files
    model_config = torch.load(model_path, map_location="cuda" if torch.cuda.is_available() else "cpu", weights_only=False)
    particles = model_config["particles"]
    pls.observation_noise = model_config["observation_noise"]
    print(f"Loaded particles and observation_noise from {model_path=}.")
    best_lr = model_config.get("best_lr")
    number_of_epochs = model_config.get("number_of_epochs")
    if torch.cuda.is_available():
        particles = particles.to(device="cuda")
    return pls, particles, best_lr, number_of_epochs
