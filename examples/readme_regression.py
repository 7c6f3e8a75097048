"""The reference README's 1-D regression demo (README.md:86-110,141-153,181-216,254-265 of jswu18/projected-langevin-sampling)
on the B200 path: same calls, same seeds, imports switched to projected_langevin_sampling_b200.

    python examples/readme_regression.py            # needs a B200 and the built library (python -c "import __graft_entry__ as g; g.build()")
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import projected_langevin_sampling_b200 as pls_b200  # noqa: E402
from projected_langevin_sampling_b200 import ConditionalVarianceInducingPointSelector, OrthonormalBasis, PLS, PLSKernel  # noqa: E402
from projected_langevin_sampling_b200.projected_langevin_sampling.costs import GaussianCost  # noqa: E402
from projected_langevin_sampling_b200.projected_langevin_sampling.link_functions import IdentityLinkFunction  # noqa: E402
from projected_langevin_sampling_b200.trainers import train_pls  # noqa: E402

torch.set_default_dtype(torch.float64)  # README.md:86-87
pls_b200.set_seed(0)

# data (README.md:94-110)
number_of_data_points, observation_noise, seed = 100, 0.1, 0
x = torch.linspace(-1, 1, number_of_data_points).reshape(-1, 1)
y = torch.sin(2 * torch.pi * x.reshape(-1)) + observation_noise * torch.normal(
    mean=0.0, std=1.0, size=(number_of_data_points,), generator=torch.Generator().manual_seed(seed))

# kernel + inducing points (README.md:141-153); a gpytorch ScaleKernel(RBFKernel()) is accepted here as well
kernel = pls_b200.ScaleKernel(pls_b200.RBFKernel(lengthscale=0.15), outputscale=3.0)
x_induce, induce_indices = ConditionalVarianceInducingPointSelector()(x=x, m=int(math.sqrt(number_of_data_points)), kernel=kernel)

# basis, cost, PLS (README.md:181-216)
pls_kernel = PLSKernel(base_kernel=kernel, approximation_samples=x_induce)
basis = OrthonormalBasis(kernel=pls_kernel, x_induce=x_induce, x_train=x)
cost = GaussianCost(observation_noise=0.5, y_train=y, link_function=IdentityLinkFunction())
pls = PLS(basis=basis, cost=cost)
particles = pls.initialise_particles(number_of_particles=100, seed=seed)

# the Langevin loop (README.md:254-265)
number_of_epochs, step_size = 200, 1e-3
for _ in range(number_of_epochs):
    particles += pls.calculate_particle_update(particles=particles, step_size=step_size)
print(f"energy potential after {number_of_epochs} epochs: {pls.calculate_energy_potential(particles):.6f}")

# the same loop through the fused trainer (one forward pass per epoch serves the update and the energy)
particles2 = pls.initialise_particles(number_of_particles=100, seed=seed)
pls_b200.set_seed(0)
particles2, energies = train_pls(pls, particles2, number_of_epochs, step_size, early_stopper_patience=1e9)
print(f"train_pls: {len(energies)} accepted epochs, final energy {energies[-1]:.6f}")

prediction = pls.predict_untransformed_samples(particles=particles, x=x, noise=torch.zeros(basis.approximation_dimension + x.shape[0], 100))
print("posterior mean |error| on the training inputs:", float((prediction.mean(dim=1).cpu() - torch.sin(2 * torch.pi * x.reshape(-1))).abs().mean()))
