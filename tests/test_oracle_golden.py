"""The oracle against every golden vector the reference's own tests hold for the hot path (ported here with the
reference test file:line each one comes from) and against fixtures produced by executing the unmodified
reference behind a gpytorch stub (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle.pls_oracle import (
    Cost,
    LinearKernel,
    Link,
    OrthonormalBasisOracle,
    PLSOracle,
    RBFScaleKernel,
    conditional_variance_select,
    r_kernel,
    sample_multivariate_normal,
    set_seed,
)

Z2 = torch.tensor([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])
X5 = torch.tensor([[1.1, 3.5, 3.5], [1.3, 7.5, 1.5], [2.5, 2.5, 0.5], [1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])
P23 = torch.tensor(
    [[1.5409960746765137, -0.293428897857666, -2.1787893772125244], [0.5684312582015991, -1.0845223665237427, -1.3985954523086548]]
)
F53 = torch.tensor(
    [
        [2.8012976646, -5.3903346062, -6.9202804565],
        [6.0996646881, -6.6724033356, -11.9823102951],
        [5.1316790581, -3.4285004139, -8.7493305206],
        [1.8773055077, -4.0070199966, -4.8781495094],
        [2.7915179729, -4.9768695831, -6.6556425095],
    ]
)


def onb(threshold=0.0, **kw):
    return OrthonormalBasisOracle(LinearKernel(), Z2, X5, eigenvalue_threshold=threshold, **kw)


# reference tests/test_basis.py:11-71
@pytest.mark.parametrize("threshold,dim", [(0.0, 2), (1.0, 1)])
def test_onb_approximation_dimension(threshold, dim):
    assert onb(threshold).approximation_dimension == dim


# reference tests/test_basis.py:118-188
@pytest.mark.parametrize("threshold,expected", [(0.0, P23), (1.0, P23[:1])])
def test_onb_initialised_particles(threshold, expected):
    assert torch.allclose(onb(threshold).initialise_particles(3, seed=0), expected)


# reference tests/test_basis.py:272-328
def test_onb_forward():
    assert torch.allclose(onb().forward(P23), F53)


# reference tests/test_basis.py:390-453 (MockCost returns ones, mockers/cost.py)
def test_onb_energy_potential():
    assert np.allclose(onb().energy_potential(P23, torch.ones(3)), 56.62522888183594)


# reference tests/test_basis.py:522-635 (first case; second adds a StudentT draw)
def test_onb_sample_predictive_noise():
    xs = torch.tensor([[3.0, 2.0, 3.2], [1.5, 6.5, 1.5]])
    expected = torch.tensor([[0.0851, -0.1569, -0.2067], [3.1662, 4.6236, -1.2954], [1.3697, 1.4171, 0.7368], [3.9759, 6.4164, -2.9854]])
    # the reference test builds the ONB with MockProjectedLangevinSamplingKernel whose forward is the plain linear
    # kernel (mockers/kernel.py:26-43), not the r-kernel; replay that
    basis = onb()
    gram_x = LinearKernel()(xs, xs)
    off = LinearKernel()(xs, Z2) @ basis.scaled_eigenvectors @ torch.diag(basis.eigenvalues)
    cov = torch.cat([torch.cat([torch.diag(basis.eigenvalues), off.T], 1), torch.cat([off, gram_x], 1)], 0)
    set_seed(0)
    got = sample_multivariate_normal(torch.zeros(4), cov, size=(3,)).T
    assert torch.allclose(got, expected, rtol=1e-3, atol=2e-4)


# reference tests/test_basis.py:754-862 (first case: noise given)
def test_onb_predict_untransformed_samples():
    xs = torch.tensor([[3.0, 2.0, 3.2], [1.5, 6.5, 1.5]])
    noise = torch.tensor([[0.0851, -0.1569, -0.2067], [3.1662, 4.6236, -1.2954], [1.3697, 1.4171, 0.7368], [3.9759, 6.4164, -2.9854]])
    expected = torch.tensor([[-8.1948, -24.4906, -2.6199], [-6.7366, -23.3037, -7.1726]])
    assert torch.allclose(onb().predict_untransformed_samples(P23, xs, noise=noise), expected, rtol=1e-3)


F22 = torch.tensor([[4.1, 3.2], [-9.3, 2.5]])
FB = torch.tensor([[0.1, 0.2], [0.9, 0.5]]).double()
COSTS = {
    "bernoulli": lambda link="sigmoid": Cost("bernoulli", torch.tensor([0.0, 1.0]), Link(link)),
    "gaussian": lambda link="identity": Cost("gaussian", torch.tensor([2.4, -2.3]), Link(link), observation_noise=1.0),
    "poisson": lambda link="square": Cost("poisson", torch.tensor([2.4, 2.3]), Link(link)),
    "student_t": lambda link="identity": Cost("student_t", torch.tensor([2.4, 2.3]), Link(link), degrees_of_freedom=3),
    "multimodal": lambda link="identity": Cost(
        "multimodal", torch.tensor([2.4, 2.3]), Link(link), observation_noise=1.0, shift=15.8, bernoulli_noise=0.8
    ),
}


# reference tests/test_costs.py:78-143
@pytest.mark.parametrize(
    "kind,f,expected",
    [
        ("bernoulli", FB, torch.tensor([1.0856, 1.2722]).double()),
        ("gaussian", F22, torch.tensor([[25.9450, 11.8400]])),
        ("poisson", F22, torch.tensor([86.2692, 6.6919])),
        ("student_t", F22, torch.tensor([9.0002, 0.4132])),
        ("multimodal", F22, torch.tensor([73.7818, 5.3968])),
    ],
)
def test_cost_values(kind, f, expected):
    assert torch.allclose(COSTS[kind]().value(f), expected, rtol=1e-3)


# reference tests/test_costs.py:146-213 (closed forms; multimodal = autograd in the reference)
@pytest.mark.parametrize(
    "kind,f,expected",
    [
        ("bernoulli", FB, torch.tensor([[0.5250, 0.5498], [-0.2891, -0.3775]]).double()),
        ("gaussian", F22, torch.tensor([[1.7000, 0.8000], [-7.0000, 4.8000]])),
        ("poisson", F22, torch.tensor([[7.0293, 4.9000], [-18.1054, 3.1600]])),
        ("student_t", F22, torch.tensor([[1.1545, 0.8791], [-0.3373, 0.2632]])),
        ("multimodal", F22, torch.tensor([[1.7000, 0.8000], [-11.6000, 0.2000]])),
    ],
)
def test_cost_derivatives(kind, f, expected):
    c = COSTS[kind]()
    assert torch.allclose(c.derivative(f), expected, rtol=1e-3)
    # the closed chain-rule form the CUDA functors implement gives the same numbers
    assert torch.allclose(c.derivative_chain_rule(f), expected, rtol=1e-3)


# reference tests/test_costs.py:216-271 (force_autograd=True, incl. Bernoulli-probit and Poisson-identity)
@pytest.mark.parametrize(
    "kind,link,f,expected",
    [
        ("bernoulli", "sigmoid", FB, torch.tensor([[0.5250, 0.5498], [-0.2891, -0.3775]]).double()),
        ("bernoulli", "probit", FB, torch.tensor([[0.8626, 0.9294], [-0.3261, -0.5092]]).double()),
        ("gaussian", "identity", F22, torch.tensor([[1.7000, 0.8000], [-7.0000, 4.8000]])),
        ("poisson", "identity", F22, torch.tensor([[-0.1707, -0.5000], [1.4946, -0.8400]])),
    ],
)
def test_cost_derivatives_autograd(kind, link, f, expected):
    c = COSTS[kind](link)
    assert torch.allclose(c.derivative(f, force_autograd=True), expected, rtol=1e-3)
    assert torch.allclose(c.derivative_chain_rule(f), expected, rtol=1e-3)


# reference tests/test_pls_kernel.py:8-52
def test_r_kernel():
    z = torch.tensor([[1.1, 3.5, 3.5], [1.3, 7.5, 1.5], [2.5, 2.5, 0.5]])
    got = r_kernel(LinearKernel(), z, torch.tensor([[1.0, 2.0, 3.0]]), torch.tensor([[1.5, 2.5, 3.5]]))
    assert torch.allclose(got, torch.tensor(355.60004))
    z = torch.tensor([[1.1, 3.5], [1.3, 7.5], [2.5, 2.5]])
    got = r_kernel(LinearKernel(), z, torch.tensor([[1.0, 3.0], [3.0, 5.0]]), torch.tensor([[1.5, 3.5]]))
    assert torch.allclose(got, torch.tensor([[319.13333], [568.8667]]))


# reference tests/test_inducing_point_selectors.py:65-120
@pytest.mark.parametrize(
    "threshold,x,z",
    [
        (0.0, torch.tensor([[1.1, 3.5, 3.5], [1.3, 7.5, 1.5], [2.5, 2.5, 0.5], [1.5, 2.5, 3.5]]), torch.tensor([[1.3, 7.5, 1.5], [1.5, 2.5, 3.5]])),
        (10.0, torch.tensor([[1.0, 3.0], [3.0, 5.0], [1.1, 3.5], [1.3, 7.5], [2.5, 2.5]]), torch.tensor([[1.3, 7.5], [3.0, 5.0]])),
    ],
)
def test_selector_reference_cases(threshold, x, z):
    set_seed(0)
    got, _ = conditional_variance_select(x, 2, LinearKernel(), threshold=threshold)
    assert torch.allclose(got, z)


# reference tests/test_samplers.py:11-65
def test_sampler():
    set_seed(0)
    s = sample_multivariate_normal(torch.zeros(2), torch.eye(2), size=(2,), seed=0)
    exp = torch.tensor([[1.5409960746765137, -2.1787893772125244], [-0.293428897857666, 0.5684312582015991]])
    assert np.allclose(s, exp, rtol=1e-3)
    set_seed(0)
    s = sample_multivariate_normal(torch.zeros(2), torch.eye(2), size=(2,), seed=None)
    assert np.allclose(s, exp, rtol=1e-3)
    set_seed(0)
    s = sample_multivariate_normal(torch.zeros(2), torch.eye(2))
    assert np.allclose(s, torch.tensor([[1.5410, -0.2934]]), rtol=1e-3)


# reference tests/test_set_seed.py:7-16
@pytest.mark.parametrize("seed,expected", [(0, 4), (1, 5)])
def test_set_seed(seed, expected):
    set_seed(seed)
    assert torch.randint(low=0, high=10, size=(1,)).item() == expected


# ---------------------------------------------------------------------------------------------------------------
# fixtures produced by running the unmodified reference (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------
def _cost_from_name(name, g):
    kind, link = name.rsplit("_", 1)
    y = {"gaussian": "y_real", "student_t": "y_real", "multimodal": "y_real", "bernoulli": "y_bin", "poisson": "y_cnt"}[kind]
    kw = dict(
        gaussian=dict(observation_noise=0.3),
        bernoulli={},
        poisson={},
        student_t=dict(degrees_of_freedom=4.0, scale=0.7),
        multimodal=dict(observation_noise=0.4, shift=1.5, bernoulli_noise=0.3),
    )[kind]
    return Cost(kind, torch.from_numpy(g[y]), Link(link), **kw)


ALL_COSTS = [
    "gaussian_identity", "gaussian_square", "bernoulli_sigmoid", "bernoulli_probit",
    "poisson_square", "poisson_identity", "student_t_identity", "multimodal_identity",
]


@pytest.mark.parametrize("name", ALL_COSTS)
def test_one_step_against_reference_run(name, golden_dir):
    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(os.path.join(golden_dir, "one_step_all_costs.npz"))
        kern = RBFScaleKernel(lengthscale=torch.from_numpy(g["lengthscale"]), outputscale=float(g["outputscale"]))
        basis = OrthonormalBasisOracle(kern, torch.from_numpy(g["z"]), torch.from_numpy(g["x"]))
        assert torch.allclose(basis.eigenvalues, torch.from_numpy(g["eigenvalues"]), rtol=1e-12, atol=0)
        p = torch.from_numpy(g["p"])
        assert torch.allclose(basis.forward(p), torch.from_numpy(g["f"]), rtol=1e-12, atol=1e-13)
        pls = PLSOracle(basis, _cost_from_name(name, g))
        dc = pls.calculate_cost_derivative(p)
        assert torch.allclose(dc, torch.from_numpy(g[name + "__dc"]), rtol=1e-11, atol=1e-12)
        # chain-rule closed form == what the reference computed (autograd for the non-matching links)
        assert torch.allclose(pls.cost.derivative_chain_rule(basis.forward(p)), torch.from_numpy(g[name + "__dc"]), rtol=1e-9, atol=1e-11)
        torch.manual_seed(int(g["noise_seed"]))
        delta = pls.calculate_particle_update(p, float(g["step_size"]))
        assert torch.allclose(delta, torch.from_numpy(g[name + "__delta"]), rtol=1e-11, atol=1e-12)
        assert torch.allclose(pls.calculate_cost(p), torch.from_numpy(g[name + "__cost"]), rtol=1e-11)
        assert np.isclose(pls.calculate_energy_potential(p), float(g[name + "__energy"]), rtol=1e-11)
    finally:
        torch.set_default_dtype(torch.float32)


def test_readme_demo_against_reference_run(golden_dir):
    """BASELINE config 1: selector -> ONB -> 200 Langevin steps, replayed with the oracle."""
    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(os.path.join(golden_dir, "readme_demo.npz"))
        x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
        kern = RBFScaleKernel(lengthscale=float(g["lengthscale"]), outputscale=float(g["outputscale"]))
        set_seed(0)
        z, idx = conditional_variance_select(x, 10, kern)
        assert idx.tolist() == g["induce_idx"].tolist()
        assert torch.equal(z, torch.from_numpy(g["x_induce"]))
        basis = OrthonormalBasisOracle(kern, z, x)
        pls = PLSOracle(basis, Cost("gaussian", y, Link("identity"), observation_noise=float(g["observation_noise"])))
        p = basis.initialise_particles(100, seed=0)
        assert torch.equal(p, torch.from_numpy(g["p0"]))
        torch.manual_seed(int(g["noise_seed"]))
        snaps = {}
        energies = []
        for s in range(200):
            p = p + pls.calculate_particle_update(p, float(g["step_size"]))
            energies.append(pls.calculate_energy_potential(p))
            if s + 1 in (1, 10, 200):
                snaps[s + 1] = p.clone()
        for k in (1, 10, 200):
            assert torch.allclose(snaps[k], torch.from_numpy(g[f"p{k}"]), rtol=1e-10, atol=1e-11), k
        assert np.allclose(energies, g["energies"], rtol=1e-10)
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("tag", ["ard", "one"])
def test_selector_against_reference_run(tag, golden_dir):
    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(os.path.join(golden_dir, "selector_runs.npz"))
        kern = RBFScaleKernel(lengthscale=torch.as_tensor(g[f"{tag}_ls"]), outputscale=float(g[f"{tag}_os"]))
        set_seed(int(g[f"{tag}_seed"]))
        z, idx = conditional_variance_select(torch.from_numpy(g[f"{tag}_x"]), int(g[f"{tag}_m"]), kern)
        assert idx.tolist() == g[f"{tag}_idx"].tolist()
        assert torch.equal(z, torch.from_numpy(g[f"{tag}_z"]))
    finally:
        torch.set_default_dtype(torch.float32)


# ---- the caller of the hot path: experiments/trainers.py:139-162 + experiments/early_stopper.py:4-24 -------------------
@pytest.mark.parametrize("name,kind", [("full", "gaussian"), ("stopped", "student_t")])
def test_train_loop_against_reference_run(name, kind, golden_dir):
    """The reference's own `train_pls` was run by tests/golden/make_golden.py; `full` uses all 40 epochs, `stopped` is
    ended by the EarlyStopper after 32 accepted epochs of 400."""
    from oracle.pls_oracle import train_pls_oracle

    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(os.path.join(golden_dir, "train_loop_runs.npz"))
        x, z = torch.from_numpy(g["x"]), torch.from_numpy(g["z"])
        eig = (torch.from_numpy(g["eigenvalues"]), torch.from_numpy(g["eigenvectors"]))
        basis = OrthonormalBasisOracle(RBFScaleKernel(torch.from_numpy(g["lengthscale"]), float(g["outputscale"])), z, x,
                                       eigenvalue_threshold=float(g["threshold"]), eig=eig)
        cost = (Cost("gaussian", torch.from_numpy(g["y"]), Link("identity"), observation_noise=0.2) if kind == "gaussian"
                else Cost("student_t", torch.from_numpy(g["y"]), Link("identity"), degrees_of_freedom=4.0, scale=0.5))
        pls = PLSOracle(basis, cost)
        torch.manual_seed(int(g[name + "__noise_seed"]))
        p, energies = train_pls_oracle(pls, torch.from_numpy(g[name + "__p0"]).clone(), int(g[name + "__epochs"]),
                                       float(g[name + "__eta"]), float(g[name + "__patience"]))
        want_e = g[name + "__energies"]
        assert len(energies) == len(want_e)
        assert np.allclose(energies, want_e, rtol=1e-10)
        want_p = torch.from_numpy(g[name + "__p"])
        assert (p - want_p).abs().max().item() <= 1e-10 * want_p.abs().max().item()
    finally:
        torch.set_default_dtype(torch.float32)


# ---- InducingPointBasis ("next" row): reference tests/test_basis.py:74-115,191-269,331-387,456-519 + a reference run ------
Y2 = torch.tensor([2.1, 3.3])
F53_IPB = torch.tensor(
    [
        [1.2656511068, -0.8267806172, -2.1464431286],
        [-6.1349906921, -5.1450047493, 4.8272337914],
        [-8.9970912933, -5.7714910507, 8.1600494385],
        [1.5409765244, -0.2934455872, -2.1787719727],
        [0.5684509277, -1.0845184326, -1.3986053467],
    ]
)


def ipb():
    from oracle.pls_oracle import InducingPointBasisOracle

    return InducingPointBasisOracle(LinearKernel(), Z2.double(), Y2.double(), X5.double(), r_kernel_is_base=True)


def test_ipb_reference_vectors():
    b = ipb()
    assert b.approximation_dimension == 2  # tests/test_basis.py:98-116
    assert torch.allclose(b.initialise_particles(3, seed=0).float(), P23)  # :248-269 (noise only)
    assert torch.allclose(b.initialise_particles(3, seed=0, noise_only=False).float(), Y2[:, None] + P23)
    # the 2 x 2 linear Gram has condition number ~1e4: float32 golden values, float64 here
    assert torch.allclose(b.forward(P23.double()).float(), F53_IPB, rtol=1e-3, atol=1e-3)  # :369-387
    assert np.isclose(b.energy_potential(P23.double(), torch.ones(3, dtype=torch.float64)), 275.2294006347656, rtol=1e-3)  # :495-519


def test_ipb_against_reference_run(golden_dir):
    from oracle.pls_oracle import InducingPointBasisOracle

    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(os.path.join(golden_dir, "ipb_runs.npz"))
        x, z, y = torch.from_numpy(g["x"]), torch.from_numpy(g["z"]), torch.from_numpy(g["y"])
        basis = InducingPointBasisOracle(RBFScaleKernel(torch.from_numpy(g["lengthscale"]), float(g["outputscale"])), z,
                                         torch.from_numpy(g["y_induce"]), x)
        pls = PLSOracle(basis, Cost("gaussian", y, Link("identity"), observation_noise=float(g["observation_noise"])))
        assert torch.equal(basis.initialise_particles(5, seed=5), torch.from_numpy(g["p_noise"]))
        p = basis.initialise_particles(5, seed=5, noise_only=False)
        assert torch.allclose(p, torch.from_numpy(g["p"]), rtol=1e-14)
        tol = 1e-10
        rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()  # noqa: E731
        assert rel(basis.forward(p), torch.from_numpy(g["f"])) < tol
        assert rel(pls.calculate_cost_derivative(p), torch.from_numpy(g["dc"])) < tol
        assert rel(pls.calculate_cost(p), torch.from_numpy(g["cost"])) < tol
        assert abs(pls.calculate_energy_potential(p) - float(g["energy"])) < tol * abs(float(g["energy"]))
        torch.manual_seed(int(g["noise_seed"]))
        assert rel(pls.calculate_particle_update(p, float(g["step_size"])), torch.from_numpy(g["delta"])) < tol
    finally:
        torch.set_default_dtype(torch.float32)


def test_ipb_predictive_noise_and_predict_reference_vectors():
    """reference tests/test_basis.py:640-838 (predictive noise, seed 0, no extra distribution) and :862-1010 (predict with the
    noise given); the reference builds the basis with its mock r-kernel = the plain linear kernel."""
    xs = torch.tensor([[3.0, 2.0, 3.2], [1.5, 6.5, 1.5]])
    noise_want = torch.tensor([[1.4442, 3.7593, -0.4158], [1.9489, 4.2264, -0.9286], [3.0377, 3.1129, -3.4442], [1.4840, 1.3103, 0.6729]])
    k = LinearKernel()
    cov = torch.cat([torch.cat([k(Z2, Z2), k(Z2, xs)], 1), torch.cat([k(Z2, xs).T, k(xs, xs)], 1)], 0)
    set_seed(0)
    got = sample_multivariate_normal(torch.zeros(4), cov, size=(3,)).T
    # the 4 x 4 linear Gram of 3-D points has rank 3: its fourth eigenvalue is float32 round-off (~1e-6) and contributes
    # sqrt(1e-6) ~ 1e-3 of noise that depends on how the matrix was assembled, hence the absolute tolerance
    assert torch.allclose(got, noise_want, rtol=1e-3, atol=1e-2)
    pred = ipb().predict_untransformed_samples(P23.double(), xs.double(), noise_want.double())
    assert torch.allclose(pred.float(), torch.tensor([[-4.4373, -3.6672, 2.9305], [-7.8718, -6.6582, 8.8616]]), rtol=2e-3, atol=2e-3)


def test_gaussian_normal_equations_identity_on_the_oracle():
    """The algebra behind the opt-in Gaussian shortcut (LangevinEngine._normal_equations), checked on the CPU oracle alone: with the
    Gaussian cost and the identity link the oracle's update and energy equal the M x M forms built from A' = k(Z,X) k(X,Z) / s,
    b' = k(Z,X) y / s (orthonormal.py:98-108,151-158 re-associated with costs/gaussian.py:54-88)."""
    g = torch.Generator().manual_seed(3)
    n, d, m, j, s_obs, eta = 400, 3, 24, 17, 0.3, 1e-3
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    z = x[:m].clone()
    kernel = RBFScaleKernel(torch.tensor([1.2, 1.5, 1.8], dtype=torch.float64), 1.4)
    basis = OrthonormalBasisOracle(kernel, z, x, eigenvalue_threshold=1e-9)
    orc = PLSOracle(basis, Cost("gaussian", y, Link("identity"), observation_noise=s_obs))
    m_k = basis.approximation_dimension
    p = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    xi = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    k_xz = kernel(x, z)
    vt, lam = basis.scaled_eigenvectors, basis.eigenvalues
    a, b = k_xz.T @ k_xz / s_obs, k_xz.T @ y / s_obs
    w = vt @ p
    aw = a @ w
    delta = -eta * (vt.T @ (aw - b[:, None])) - eta * p / lam[:, None] + np.sqrt(2 * eta) * xi
    want = orc.calculate_particle_update(p, eta, noise=xi)
    assert (delta - want).abs().max() <= 1e-11 * want.abs().max()
    cost = 0.5 * (w * aw).sum(0) - b @ w + (y @ y) / (2 * s_obs)
    energy = (cost + 0.5 * (p * p / lam[:, None]).sum(0)).mean().item()
    assert abs(energy - orc.calculate_energy_potential(p)) <= 1e-11 * abs(energy)
