"""Generate golden fixtures by EXECUTING THE UNMODIFIED REFERENCE (`/root/reference`) in the build container.

The reference imports gpytorch at module top level and gpytorch is not installed here, so its modules are run
behind `oracle/gpytorch_stub` (dense kernel evaluation; RBF arithmetic = the oracle's restatement of the
gpytorch formula).  Everything else -- OrthonormalBasis, the costs (incl. the autograd derivative), PLS, the
sampler, the ConditionalVariance selector -- is the reference's own code, so these fixtures pin the parts the
reference's unit tests leave unpinned (the Langevin update, whole trajectories, selector runs with m > 2).

Run from the repo root (only possible where /root/reference exists; the fixtures are committed):
    python tests/golden/make_golden.py
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "gpytorch_stub"), ROOT, "/root/reference"]

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

import gpytorch  # noqa: E402  (the stub)
from src.inducing_point_selectors import ConditionalVarianceInducingPointSelector  # noqa: E402
from src.projected_langevin_sampling import PLS, PLSKernel  # noqa: E402
from src.projected_langevin_sampling.basis import OrthonormalBasis  # noqa: E402
from src.projected_langevin_sampling.costs import (  # noqa: E402
    BernoulliCost,
    GaussianCost,
    MultiModalCost,
    PoissonCost,
    StudentTCost,
)
from src.projected_langevin_sampling.link_functions import (  # noqa: E402
    IdentityLinkFunction,
    ProbitLinkFunction,
    SigmoidLinkFunction,
    SquareLinkFunction,
)
from src.utils import set_seed  # noqa: E402


def make_kernel(lengthscale, outputscale, ard=None):
    k = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=ard))
    k.base_kernel.lengthscale = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(1, -1)
    k.outputscale = outputscale
    return k


def readme_demo():
    """BASELINE config 1 = README.md:86-110,141-153,181-216,254-265."""
    set_seed(0)
    n, obs, seed = 100, 0.1, 0
    x = torch.linspace(-1, 1, n).reshape(-1, 1)
    y = torch.sin(2 * torch.pi * x.reshape(-1)) + obs * torch.normal(
        mean=torch.tensor(0), std=torch.tensor(1), generator=torch.Generator().manual_seed(seed), size=(n,)
    ).reshape(-1)
    kernel = make_kernel(0.15, 3.0)
    x_induce, induce_idx = ConditionalVarianceInducingPointSelector()(x=x, m=int(math.sqrt(n)), kernel=kernel)
    pls_kernel = PLSKernel(base_kernel=kernel, approximation_samples=x_induce)
    basis = OrthonormalBasis(kernel=pls_kernel, x_induce=x_induce, x_train=x)
    cost = GaussianCost(observation_noise=0.5, y_train=y, link_function=IdentityLinkFunction())
    pls = PLS(basis=basis, cost=cost)
    particles = pls.initialise_particles(number_of_particles=100, seed=seed)
    p0 = particles.clone()
    # the Langevin noise stream: replayable as `torch.manual_seed(1234)` + successive torch.normal((M_k, J)) calls
    torch.manual_seed(1234)
    snaps, energies = {}, []
    for step in range(200):
        particles += pls.calculate_particle_update(particles=particles, step_size=1e-3)
        energies.append(pls.calculate_energy_potential(particles=particles))
        if step + 1 in (1, 10, 200):
            snaps[step + 1] = particles.clone()
    np.savez(
        os.path.join(HERE, "readme_demo.npz"),
        x=x.numpy(), y=y.numpy(), induce_idx=induce_idx.numpy(), x_induce=x_induce.numpy(),
        eigenvalues=basis.eigenvalues.numpy(), eigenvectors=basis.eigenvectors.numpy(),
        p0=p0.numpy(), p1=snaps[1].numpy(), p10=snaps[10].numpy(), p200=snaps[200].numpy(),
        energies=np.array(energies), noise_seed=1234, step_size=1e-3, lengthscale=0.15, outputscale=3.0,
        observation_noise=0.5,
    )
    print("readme_demo: M_k", basis.approximation_dimension, "idx", induce_idx.tolist(), "E0..", energies[:2], energies[-1])


def one_step_all_costs():
    """One Langevin update per (cost, link) on a small ARD problem; pins orthonormal.py:128-159 + every derivative."""
    g = torch.Generator().manual_seed(7)
    n, d, m, j = 60, 3, 9, 7
    x = torch.randn(n, d, generator=g)
    z = x[torch.randperm(n, generator=g)[:m]].clone()
    ls = torch.tensor([0.9, 1.3, 1.7])
    kernel = make_kernel(ls, 1.7, ard=d)
    basis = OrthonormalBasis(kernel=PLSKernel(base_kernel=kernel, approximation_samples=z), x_induce=z, x_train=x)
    p = torch.randn(basis.approximation_dimension, j, generator=g)
    y_real = torch.randn(n, generator=g)
    y_bin = (torch.rand(n, generator=g) > 0.5).double()
    y_cnt = torch.poisson(torch.full((n,), 2.0), generator=g)
    cases = {
        "gaussian_identity": GaussianCost(observation_noise=0.3, y_train=y_real, link_function=IdentityLinkFunction()),
        "gaussian_square": GaussianCost(observation_noise=0.3, y_train=y_real, link_function=SquareLinkFunction()),
        "bernoulli_sigmoid": BernoulliCost(y_train=y_bin, link_function=SigmoidLinkFunction()),
        "bernoulli_probit": BernoulliCost(y_train=y_bin, link_function=ProbitLinkFunction()),
        "poisson_square": PoissonCost(y_train=y_cnt, link_function=SquareLinkFunction()),
        "poisson_identity": PoissonCost(y_train=y_cnt, link_function=IdentityLinkFunction()),
        "student_t_identity": StudentTCost(degrees_of_freedom=4.0, y_train=y_real, link_function=IdentityLinkFunction(), scale=0.7),
        "multimodal_identity": MultiModalCost(observation_noise=0.4, shift=1.5, bernoulli_noise=0.3, y_train=y_real,
                                              link_function=IdentityLinkFunction()),
    }
    out = dict(x=x.numpy(), z=z.numpy(), lengthscale=ls.numpy(), outputscale=1.7, p=p.numpy(),
               eigenvalues=basis.eigenvalues.numpy(), eigenvectors=basis.eigenvectors.numpy(),
               y_real=y_real.numpy(), y_bin=y_bin.numpy(), y_cnt=y_cnt.numpy(), step_size=2e-3, noise_seed=99)
    f = basis.calculate_untransformed_train_prediction_samples(p)
    out["f"] = f.numpy()
    for name, cost in cases.items():
        pls = PLS(basis=basis, cost=cost)
        dc = pls.calculate_cost_derivative(p)
        torch.manual_seed(99)
        delta = pls.calculate_particle_update(p, 2e-3)
        out[name + "__dc"] = dc.numpy()
        out[name + "__delta"] = delta.numpy()
        out[name + "__cost"] = pls.calculate_cost(p).numpy()
        out[name + "__energy"] = pls.calculate_energy_potential(p)
        print(name, "ok", float(delta.abs().max()))
    np.savez(os.path.join(HERE, "one_step_all_costs.npz"), **out)


def selector_runs():
    """ConditionalVariance runs with m > 2 (the reference's own tests stop at m = 2)."""
    out = {}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(300, 4, generator=g)
    kernel = make_kernel(torch.tensor([1.1, 0.8, 1.4, 2.0]), 2.5, ard=4)
    set_seed(11)
    z, idx = ConditionalVarianceInducingPointSelector()(x=x, m=24, kernel=kernel)
    out.update(ard_x=x.numpy(), ard_ls=np.array([1.1, 0.8, 1.4, 2.0]), ard_os=2.5, ard_seed=11, ard_m=24,
               ard_idx=idx.numpy(), ard_z=z.numpy())
    x1 = torch.rand(500, 1, generator=g) * 6 - 3
    kernel1 = make_kernel(0.5, 1.0)
    set_seed(5)
    z1, idx1 = ConditionalVarianceInducingPointSelector()(x=x1, m=16, kernel=kernel1)
    out.update(one_x=x1.numpy(), one_ls=0.5, one_os=1.0, one_seed=5, one_m=16, one_idx=idx1.numpy(), one_z=z1.numpy())
    np.savez(os.path.join(HERE, "selector_runs.npz"), **out)
    print("selector:", idx.tolist()[:8], idx1.tolist()[:8])


def scale_problem(n, d, seed=0):
    """C4-style inputs (X ~ N(0, I), ARD lengthscales sqrt(D) (0.75 + 0.5 k / (D - 1))), regenerated from the
    seed by the tests -- only the selected indices are stored."""
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    ls = torch.tensor([math.sqrt(d) * (0.75 + 0.5 * k / max(d - 1, 1)) for k in range(d)], dtype=torch.float64)
    return x, ls


def selector_scale(sizes=((100_000, 256), (1_000_000, 1024))):
    """The reference's ConditionalVariance selector at C4-like scale (D = 8 ARD): N = 100 000, M = 256 and -- host memory
    permitting (the reference's C is 8.2 GB) -- the C4 shape itself, N = 1 000 000, M = 1024.  Indices only."""
    import time

    path = os.path.join(HERE, "selector_scale.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for n, m in sizes:
        x, ls = scale_problem(n, 8)
        kernel = make_kernel(ls, 1.0, ard=8)
        set_seed(123)
        t0 = time.time()
        z, idx = ConditionalVarianceInducingPointSelector()(x=x, m=m, kernel=kernel)
        tag = f"n{n}_m{m}"
        out.update({tag + "_idx": idx.numpy(), tag + "_seed": 123, tag + "_data_seed": 0, tag + "_d": 8, tag + "_outputscale": 1.0})
        print("selector_scale", tag, "reference run", round(time.time() - t0, 1), "s; first", idx.tolist()[:6], flush=True)
        np.savez(path, **out)


def train_loop_runs():
    """experiments/trainers.py:139-162 `train_pls` (the caller of the hot path) with experiments/early_stopper.py:4-24:
    one run that uses all its epochs and one that the EarlyStopper ends (energies appended only for accepted steps)."""
    from experiments.trainers import train_pls  # noqa: E402  (imports src.gaussian_process: placeholders in the stub)

    g = torch.Generator().manual_seed(21)
    n, d, m, j = 80, 2, 8, 12
    x = torch.randn(n, d, generator=g)
    z = x[torch.randperm(n, generator=g)[:m]].clone()
    ls = torch.tensor([0.8, 1.2])
    kernel = make_kernel(ls, 1.4, ard=d)
    basis = OrthonormalBasis(kernel=PLSKernel(base_kernel=kernel, approximation_samples=z), x_induce=z, x_train=x,
                             eigenvalue_threshold=1e-6)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g)
    y_cnt = torch.poisson(torch.full((n,), 3.0), generator=g)
    out = dict(x=x.numpy(), z=z.numpy(), lengthscale=ls.numpy(), outputscale=1.4, y=y.numpy(), y_cnt=y_cnt.numpy(),
               eigenvalues=basis.eigenvalues.numpy(), eigenvectors=basis.eigenvectors.numpy(), threshold=1e-6)
    runs = {
        "full": (GaussianCost(observation_noise=0.2, y_train=y, link_function=IdentityLinkFunction()), 40, 2e-3, 1.0, 77),
        "stopped": (StudentTCost(degrees_of_freedom=4.0, y_train=y, link_function=IdentityLinkFunction(), scale=0.5), 400, 4e-3, 2e-2, 78),
    }
    for name, (cost, epochs, eta, patience, noise_seed) in runs.items():
        pls = PLS(basis=basis, cost=cost)
        p0 = pls.initialise_particles(number_of_particles=j, seed=3)
        torch.manual_seed(noise_seed)
        particles, energies = train_pls(pls=pls, particles=p0.clone(), number_of_epochs=epochs, step_size=eta,
                                        early_stopper_patience=patience)
        out.update({f"{name}__p0": p0.numpy(), f"{name}__p": particles.numpy(), f"{name}__energies": np.array(energies),
                    f"{name}__epochs": epochs, f"{name}__eta": eta, f"{name}__patience": patience, f"{name}__noise_seed": noise_seed})
        print("train_pls", name, "accepted", len(energies), "of", epochs, "E", energies[0], energies[-1])
    np.savez(os.path.join(HERE, "train_loop_runs.npz"), **out)


def ipb_runs():
    """The reference's InducingPointBasis (basis/inducing_point.py): forward, cost derivative, energy and one Langevin update
    on a small, well-conditioned ARD problem (k(Z, Z) is solved by Cholesky on both sides)."""
    from src.projected_langevin_sampling.basis import InducingPointBasis  # noqa: E402

    g = torch.Generator().manual_seed(41)
    n, d, m, j = 70, 2, 7, 5
    x = torch.randn(n, d, generator=g)
    z = torch.tensor([[-1.5, -1.0], [-0.7, 1.2], [0.0, 0.0], [0.8, -1.3], [1.6, 0.9], [-1.8, 1.9], [2.2, -0.4]])
    ls = torch.tensor([0.9, 1.1])
    kernel = make_kernel(ls, 1.3, ard=d)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g)
    y_induce = torch.sin(z.sum(1))
    basis = InducingPointBasis(kernel=PLSKernel(base_kernel=kernel, approximation_samples=z), x_induce=z, y_induce=y_induce, x_train=x)
    pls = PLS(basis=basis, cost=GaussianCost(observation_noise=0.3, y_train=y, link_function=IdentityLinkFunction()))
    p_noise = pls.initialise_particles(number_of_particles=j, seed=5)
    p = pls.initialise_particles(number_of_particles=j, seed=5, noise_only=False)
    torch.manual_seed(17)
    delta = pls.calculate_particle_update(p, 1e-3)
    np.savez(os.path.join(HERE, "ipb_runs.npz"), x=x.numpy(), z=z.numpy(), y=y.numpy(), y_induce=y_induce.numpy(), lengthscale=ls.numpy(),
             outputscale=1.3, observation_noise=0.3, p_noise=p_noise.numpy(), p=p.numpy(),
             f=basis.calculate_untransformed_train_prediction_samples(p).numpy(), dc=pls.calculate_cost_derivative(p).numpy(),
             cost=pls.calculate_cost(p).numpy(), energy=pls.calculate_energy_potential(p), delta=delta.numpy(), step_size=1e-3,
             noise_seed=17, cond=float(torch.linalg.cond(basis.base_gram_induce)))
    print("ipb: cond(k(Z,Z)) =", float(torch.linalg.cond(basis.base_gram_induce)), "energy", pls.calculate_energy_potential(p))


def runner_runs():
    """experiments/runners.py:331-446 `train_pls_runner` (step-size search around train_pls; metric "loss") run UNMODIFIED.  Its
    module imports the plotting layer (matplotlib is not installed): the names are satisfied by empty placeholder modules, no
    plotting function is called (plot_energy_potential_path=None).  The runner also evaluates pls.predict on the training inputs
    for every finite run (whatever the metric); set_seed(seed) at the start of every search makes the trainings independent of it."""
    import types

    class _Placeholder(types.ModuleType):
        def __getattr__(self, name):  # plt.Figure, plt.Axes, ... appear in annotations of the plotting layer only
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})

    for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot"):
        sys.modules.setdefault(name, _Placeholder(name))
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from experiments.data import Data, ExperimentData, ProblemType  # noqa: E402
    from experiments.runners import train_pls_runner  # noqa: E402

    g = torch.Generator().manual_seed(33)
    n, d, m, j = 90, 2, 8, 10
    x = torch.randn(n, d, generator=g)
    z = x[torch.randperm(n, generator=g)[:m]].clone()
    ls = torch.tensor([0.9, 1.3])
    kernel = make_kernel(ls, 1.2, ard=d)
    basis = OrthonormalBasis(kernel=PLSKernel(base_kernel=kernel, approximation_samples=z), x_induce=z, x_train=x, eigenvalue_threshold=1e-6)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g)
    pls = PLS(basis=basis, cost=GaussianCost(observation_noise=0.2, y_train=y, link_function=IdentityLinkFunction()))
    p0 = pls.initialise_particles(number_of_particles=j, seed=4)
    data = ExperimentData(name="synthetic", problem_type=ProblemType.REGRESSION, full=Data(x=x, y=y), train=Data(x=x, y=y, name="train"))
    kw = dict(simulation_duration=0.04, maximum_number_of_steps=160, early_stopper_patience=0.01, number_of_step_searches=4,
              step_size_upper=2e-3, minimum_change_in_energy_potential=1e-4, seed=11)
    particles, best_lr, epochs = train_pls_runner(pls=pls, particle_name="p", experiment_data=data, particles=p0.clone(),
                                                  metric_to_optimise="loss", **kw)
    np.savez(os.path.join(HERE, "runner_runs.npz"), x=x.numpy(), z=z.numpy(), y=y.numpy(), lengthscale=ls.numpy(), outputscale=1.2,
             observation_noise=0.2, threshold=1e-6, eigenvalues=basis.eigenvalues.numpy(), eigenvectors=basis.eigenvectors.numpy(),
             p0=p0.numpy(), particles=particles.numpy(), best_lr=float(best_lr), epochs=int(epochs),
             **{f"kw__{k}": v for k, v in kw.items()})
    print("train_pls_runner: best_lr", best_lr, "epochs", epochs, "E", pls.calculate_energy_potential(particles))


if __name__ == "__main__":
    # selector_scale takes ~15 min and 10 GB of host memory: run it by name
    which = sys.argv[1:] or ["readme_demo", "one_step_all_costs", "selector_runs", "train_loop_runs", "ipb_runs", "runner_runs"]
    for name in which:
        globals()[name]()
