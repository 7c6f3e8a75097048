"""GPU parity tests: the CUDA path (through the Python mirror of the reference API, i.e. through the C ABI) against
  (1) the reference's own golden vectors (tests/test_basis.py, test_costs.py, test_inducing_point_selectors.py),
  (2) fixtures produced by executing the unmodified reference (tests/golden/*.npz),
  (3) the CPU oracle on seeded random inputs, incl. ragged shapes (odd J, N/M not multiples of the tile, row chunks).
Tolerance for float64 results: 1e-10 relative to the result's infinity norm (BASELINE.json north_star); selector indices
must be identical."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.pls_oracle import (  # noqa: E402
    Cost,
    LinearKernel as OracleLinear,
    Link,
    OrthonormalBasisOracle,
    PLSOracle,
    RBFScaleKernel,
    conditional_variance_select,
    set_seed as oracle_set_seed,
)

TOL = 1e-10


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    scale = want.abs().max().item()
    return (got - want).abs().max().item() / (scale if scale > 0 else 1.0)


@pytest.fixture(scope="module")
def b200():
    import projected_langevin_sampling_b200 as pkg
    from projected_langevin_sampling_b200 import _native

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _native.context()  # fails loudly if the library is missing / not sm_100
    return pkg


def _costs_mod():
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions

    return costs, link_functions


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


def linear_onb(pkg, threshold=0.0):
    return pkg.OrthonormalBasis(pkg.PLSKernel(pkg.LinearKernel(), Z2), Z2, X5, eigenvalue_threshold=threshold, verbose=False)


# ---- (1) the reference's golden vectors through the CUDA path -------------------------------------------------------
@pytest.mark.parametrize("threshold,dim", [(0.0, 2), (1.0, 1)])
def test_reference_onb_dimension(b200, threshold, dim):  # reference tests/test_basis.py:11-71
    assert linear_onb(b200, threshold).approximation_dimension == dim


@pytest.mark.parametrize("threshold,expected", [(0.0, P23), (1.0, P23[:1])])
def test_reference_onb_initialised_particles(b200, threshold, expected):  # tests/test_basis.py:118-188
    p = linear_onb(b200, threshold).initialise_particles(3, seed=0)
    assert p.is_cuda and p.dtype == torch.float64
    assert torch.allclose(p.cpu().float(), expected)


def test_reference_onb_forward(b200):  # tests/test_basis.py:272-328
    f = linear_onb(b200).calculate_untransformed_train_prediction_samples(P23)
    assert torch.allclose(f.cpu().float(), F53, rtol=1e-4, atol=1e-5)


def test_reference_onb_energy(b200):  # tests/test_basis.py:390-453
    e = linear_onb(b200).calculate_energy_potential(P23, torch.ones(3))
    assert np.allclose(e, 56.62522888183594, rtol=1e-5)


def test_reference_onb_predict(b200):  # tests/test_basis.py:754-862 (noise given)
    xs = torch.tensor([[3.0, 2.0, 3.2], [1.5, 6.5, 1.5]])
    noise = torch.tensor([[0.0851, -0.1569, -0.2067], [3.1662, 4.6236, -1.2954], [1.3697, 1.4171, 0.7368], [3.9759, 6.4164, -2.9854]])
    expected = torch.tensor([[-8.1948, -24.4906, -2.6199], [-6.7366, -23.3037, -7.1726]])
    got = linear_onb(b200).predict_untransformed_samples(P23, xs, noise=noise)
    assert torch.allclose(got.cpu().float(), expected, rtol=1e-3)


def test_reference_pls_kernel(b200):  # tests/test_pls_kernel.py:8-52
    z = torch.tensor([[1.1, 3.5, 3.5], [1.3, 7.5, 1.5], [2.5, 2.5, 0.5]])
    k = b200.PLSKernel(b200.LinearKernel(), z)
    got = k(torch.tensor([[1.0, 2.0, 3.0]]), torch.tensor([[1.5, 2.5, 3.5]]))
    assert torch.allclose(got.cpu().float(), torch.tensor(355.60004))


F22 = torch.tensor([[4.1, 3.2], [-9.3, 2.5]])
FB = torch.tensor([[0.1, 0.2], [0.9, 0.5]]).double()


def _ref_cost(kind, link=None):
    costs, links = _costs_mod()
    lk = {"identity": links.IdentityLinkFunction, "sigmoid": links.SigmoidLinkFunction, "probit": links.ProbitLinkFunction,
          "square": links.SquareLinkFunction}
    if kind == "bernoulli":
        return costs.BernoulliCost(y_train=torch.tensor([0.0, 1.0]), link_function=lk[link or "sigmoid"]())
    if kind == "gaussian":
        return costs.GaussianCost(observation_noise=1.0, y_train=torch.tensor([2.4, -2.3]), link_function=lk[link or "identity"]())
    if kind == "poisson":
        return costs.PoissonCost(y_train=torch.tensor([2.4, 2.3]), link_function=lk[link or "square"]())
    if kind == "student_t":
        return costs.StudentTCost(degrees_of_freedom=3, y_train=torch.tensor([2.4, 2.3]), link_function=lk[link or "identity"]())
    return costs.MultiModalCost(observation_noise=1.0, shift=15.8, bernoulli_noise=0.8, y_train=torch.tensor([2.4, 2.3]),
                                link_function=lk[link or "identity"]())


@pytest.mark.parametrize(
    "kind,f,expected",
    [
        ("bernoulli", FB, [1.0856, 1.2722]),
        ("gaussian", F22, [25.9450, 11.8400]),
        ("poisson", F22, [86.2692, 6.6919]),
        ("student_t", F22, [9.0002, 0.4132]),
        ("multimodal", F22, [73.7818, 5.3968]),
    ],
)
def test_reference_cost_values(b200, kind, f, expected):  # tests/test_costs.py:78-143
    got = _ref_cost(kind).calculate_cost(f)
    assert torch.allclose(got.cpu(), torch.tensor(expected).double(), rtol=1e-3)


@pytest.mark.parametrize(
    "kind,link,force,f,expected",
    [
        ("bernoulli", None, False, FB, [[0.5250, 0.5498], [-0.2891, -0.3775]]),
        ("gaussian", None, False, F22, [[1.7000, 0.8000], [-7.0000, 4.8000]]),
        ("poisson", None, False, F22, [[7.0293, 4.9000], [-18.1054, 3.1600]]),
        ("student_t", None, False, F22, [[1.1545, 0.8791], [-0.3373, 0.2632]]),
        ("multimodal", None, False, F22, [[1.7000, 0.8000], [-11.6000, 0.2000]]),
        ("bernoulli", "sigmoid", True, FB, [[0.5250, 0.5498], [-0.2891, -0.3775]]),
        ("bernoulli", "probit", True, FB, [[0.8626, 0.9294], [-0.3261, -0.5092]]),
        ("gaussian", "identity", True, F22, [[1.7000, 0.8000], [-7.0000, 4.8000]]),
        ("poisson", "identity", True, F22, [[-0.1707, -0.5000], [1.4946, -0.8400]]),
    ],
)
def test_reference_cost_derivatives(b200, kind, link, force, f, expected):  # tests/test_costs.py:146-271
    got = _ref_cost(kind, link).calculate_cost_derivative(f, force_autograd=force)
    assert torch.allclose(got.cpu(), torch.tensor(expected).double(), rtol=1e-3)


@pytest.mark.parametrize(
    "threshold,x,z",
    [
        (0.0, [[1.1, 3.5, 3.5], [1.3, 7.5, 1.5], [2.5, 2.5, 0.5], [1.5, 2.5, 3.5]], [[1.3, 7.5, 1.5], [1.5, 2.5, 3.5]]),
        (10.0, [[1.0, 3.0], [3.0, 5.0], [1.1, 3.5], [1.3, 7.5], [2.5, 2.5]], [[1.3, 7.5], [3.0, 5.0]]),
    ],
)
def test_reference_selector_cases(b200, threshold, x, z):  # tests/test_inducing_point_selectors.py:65-120
    b200.set_seed(0)
    sel = b200.ConditionalVarianceInducingPointSelector(threshold=threshold)
    got, _ = sel.compute_induce_data(x=torch.tensor(x), m=2, kernel=b200.LinearKernel())
    assert torch.allclose(got, torch.tensor(z))


# ---- (2) fixtures produced by running the unmodified reference ---------------------------------------------------------
ALL_COSTS = ["gaussian_identity", "gaussian_square", "bernoulli_sigmoid", "bernoulli_probit", "poisson_square",
             "poisson_identity", "student_t_identity", "multimodal_identity"]


def _pkg_cost(name, g):
    costs, links = _costs_mod()
    kind, link = name.rsplit("_", 1)
    lk = {"identity": links.IdentityLinkFunction, "sigmoid": links.SigmoidLinkFunction, "probit": links.ProbitLinkFunction,
          "square": links.SquareLinkFunction}[link]()
    if kind == "gaussian":
        return costs.GaussianCost(observation_noise=0.3, y_train=torch.from_numpy(g["y_real"]), link_function=lk)
    if kind == "bernoulli":
        return costs.BernoulliCost(y_train=torch.from_numpy(g["y_bin"]), link_function=lk)
    if kind == "poisson":
        return costs.PoissonCost(y_train=torch.from_numpy(g["y_cnt"]), link_function=lk)
    if kind == "student_t":
        return costs.StudentTCost(degrees_of_freedom=4.0, y_train=torch.from_numpy(g["y_real"]), link_function=lk, scale=0.7)
    return costs.MultiModalCost(observation_noise=0.4, shift=1.5, bernoulli_noise=0.3, y_train=torch.from_numpy(g["y_real"]), link_function=lk)


@pytest.mark.parametrize("name", ALL_COSTS)
def test_one_step_against_reference_run(b200, name, golden_dir):
    torch.set_default_dtype(torch.float64)  # as the reference run (affects the probit constant and the noise dtype)
    try:
        g = np.load(os.path.join(golden_dir, "one_step_all_costs.npz"))
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=3, lengthscale=torch.from_numpy(g["lengthscale"])), outputscale=float(g["outputscale"]))
        z, x = torch.from_numpy(g["z"]), torch.from_numpy(g["x"])
        eig = (torch.from_numpy(g["eigenvalues"]), torch.from_numpy(g["eigenvectors"]))
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigendecomposition=eig, verbose=False)
        pls = b200.PLS(basis, _pkg_cost(name, g))
        p = torch.from_numpy(g["p"]).cuda()
        assert rel_err(basis.calculate_untransformed_train_prediction_samples(p), torch.from_numpy(g["f"])) < TOL
        assert rel_err(pls.calculate_cost_derivative(p), torch.from_numpy(g[name + "__dc"])) < TOL
        assert rel_err(pls.calculate_cost(p), torch.from_numpy(g[name + "__cost"])) < TOL
        assert abs(pls.calculate_energy_potential(p) - float(g[name + "__energy"])) <= TOL * abs(float(g[name + "__energy"]))
        torch.manual_seed(int(g["noise_seed"]))  # replay the reference's torch.normal draw
        delta = pls.calculate_particle_update(p, float(g["step_size"]))
        assert rel_err(delta, torch.from_numpy(g[name + "__delta"])) < TOL
        # the unfused composition (basis + cost objects separately) gives the same update
        torch.manual_seed(int(g["noise_seed"]))
        f = basis.calculate_untransformed_train_prediction_samples(p)
        dc = pls.cost.calculate_cost_derivative(f)
        delta2 = basis.calculate_particle_update(p, dc, float(g["step_size"]))
        assert rel_err(delta2, torch.from_numpy(g[name + "__delta"])) < TOL
    finally:
        torch.set_default_dtype(torch.float32)


def test_readme_demo_against_reference_run(b200, golden_dir):
    """BASELINE config 1: selector -> ONB -> 200 Langevin steps with the reference's noise stream."""
    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = np.load(os.path.join(golden_dir, "readme_demo.npz"))
        x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
        kernel = b200.ScaleKernel(b200.RBFKernel(lengthscale=float(g["lengthscale"])), outputscale=float(g["outputscale"]))
        # The README's linspace inputs produce EXACT ties in the conditional variances (every point further than ~6
        # lengthscales from the pivots keeps d = outputscale + jitter to the last bit).  The reference resolves them with
        # `reversed(np.argsort(d))`; the selector detects the ties on the device and asks the same numpy routine, so the
        # indices are those of the reference run, and the Langevin run below starts from the points selected HERE.
        b200.set_seed(0)
        sel = b200.ConditionalVarianceInducingPointSelector()
        z_sel, idx = sel(x=x, m=10, kernel=kernel)
        assert idx.tolist() == g["induce_idx"].tolist()  # bit-exact indices, ties included
        assert torch.equal(z_sel, torch.from_numpy(g["x_induce"]))
        assert sel.last_run_info["host_tie_calls"] >= 1 and sel.last_run_info["min_top2_rel_gap"] == 0.0
        # ... and the device-only rule (highest permuted index among tied points) equals the oracle with a stable argsort
        b200.set_seed(0)
        _, idx_s = b200.ConditionalVarianceInducingPointSelector(tie_rule="stable")(x=x, m=10, kernel=kernel)
        oracle_set_seed(0)
        _, idx_orc = conditional_variance_select(x, 10, RBFScaleKernel(float(g["lengthscale"]), float(g["outputscale"])),
                                                 argsort_kind="stable")
        assert idx_s.tolist() == idx_orc.tolist()
        z = z_sel
        eig = (torch.from_numpy(g["eigenvalues"]), torch.from_numpy(g["eigenvectors"]))
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigendecomposition=eig, verbose=False)
        # the package's own eigendecomposition agrees on the spectrum
        own = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, verbose=False)
        assert rel_err(own.eigenvalues, eig[0]) < 1e-9
        pls = b200.PLS(basis, costs.GaussianCost(observation_noise=float(g["observation_noise"]), y_train=y,
                                                 link_function=links.IdentityLinkFunction()))
        p = pls.initialise_particles(number_of_particles=100, seed=0)
        assert torch.equal(p.cpu(), torch.from_numpy(g["p0"]))
        torch.manual_seed(int(g["noise_seed"]))
        energies = []
        for s in range(200):
            p += pls.calculate_particle_update(particles=p, step_size=float(g["step_size"]))
            energies.append(pls.calculate_energy_potential(particles=p))
            if s + 1 in (1, 10, 200):
                assert rel_err(p, torch.from_numpy(g[f"p{s + 1}"])) < 1e-9, s + 1
        assert np.allclose(energies, g["energies"], rtol=1e-9)
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("tag", ["ard", "one"])
def test_selector_against_reference_run(b200, tag, golden_dir):
    """`ard`: generic 4-D inputs, no ties.  `one`: 1-D inputs where every point further than ~6 lengthscales from all pivots
    keeps d = outputscale + jitter to the last bit (exact ties, which the reference resolves with numpy's default argsort and
    the selector hands to the same routine).  Both: indices and points identical to the reference run."""
    g = np.load(os.path.join(golden_dir, "selector_runs.npz"))
    if tag == "one":
        x = torch.from_numpy(g["one_x"])
        kernel = b200.ScaleKernel(b200.RBFKernel(lengthscale=float(g["one_ls"])), outputscale=float(g["one_os"]))
        b200.set_seed(int(g["one_seed"]))
        sel = b200.ConditionalVarianceInducingPointSelector()
        z, idx = sel(x=x, m=int(g["one_m"]), kernel=kernel)
        assert idx.tolist() == g["one_idx"].tolist()
        assert torch.equal(z, torch.from_numpy(g["one_z"]))
        assert sel.last_run_info["host_tie_calls"] >= 1
        b200.set_seed(int(g["one_seed"]))
        _, idx_s = b200.ConditionalVarianceInducingPointSelector(tie_rule="stable")(x=x, m=int(g["one_m"]), kernel=kernel)
        oracle_set_seed(int(g["one_seed"]))
        _, idxo = conditional_variance_select(x, int(g["one_m"]), RBFScaleKernel(float(g["one_ls"]), float(g["one_os"])),
                                              argsort_kind="stable")
        assert idx_s.tolist() == idxo.tolist()
        return
    x = torch.from_numpy(g[f"{tag}_x"])
    ls = torch.as_tensor(g[f"{tag}_ls"]).reshape(-1)
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=x.shape[1], lengthscale=ls), outputscale=float(g[f"{tag}_os"]))
    b200.set_seed(int(g[f"{tag}_seed"]))
    z, idx = b200.ConditionalVarianceInducingPointSelector()(x=x, m=int(g[f"{tag}_m"]), kernel=kernel)
    assert idx.tolist() == g[f"{tag}_idx"].tolist()  # bit-exact indices
    assert torch.equal(z, torch.from_numpy(g[f"{tag}_z"]))


# ---- (3) seeded random problems against the oracle -----------------------------------------------------------------------
def _problem(n, d, m, j, seed, cost_kind="gaussian", link="identity"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    z = x[torch.randperm(n, generator=g)[:m]].clone()
    ls = (0.8 + 0.6 * torch.rand(d, generator=g, dtype=torch.float64)) * (d**0.5)
    if cost_kind == "bernoulli":
        y = (torch.rand(n, generator=g) > 0.5).double()
    elif cost_kind == "poisson":
        y = torch.poisson(torch.full((n,), 3.0), generator=g).double()
    else:
        y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    return x, y, z, ls, g


def _build_pair(pkg, x, y, z, ls, outputscale, cost_kind, link, threshold=1e-8, **basis_kw):
    costs, links = _costs_mod()
    orc_kernel = RBFScaleKernel(ls, outputscale)
    kzz = orc_kernel(z, z)
    eig = torch.linalg.eigh((1 / z.shape[0]) * kzz)
    orc_basis = OrthonormalBasisOracle(orc_kernel, z, x, eigenvalue_threshold=threshold, eig=eig)
    kernel = pkg.ScaleKernel(pkg.RBFKernel(ard_num_dims=x.shape[1], lengthscale=ls), outputscale=outputscale)
    basis = pkg.OrthonormalBasis(pkg.PLSKernel(kernel, z), z, x, eigenvalue_threshold=threshold, eigendecomposition=eig,
                                 verbose=False, **basis_kw)
    lk = {"identity": links.IdentityLinkFunction, "sigmoid": links.SigmoidLinkFunction, "probit": links.ProbitLinkFunction,
          "square": links.SquareLinkFunction}[link]()
    if cost_kind == "gaussian":
        cost, ocost = costs.GaussianCost(0.25, y, lk), Cost("gaussian", y, Link(link), observation_noise=0.25)
    elif cost_kind == "bernoulli":
        cost, ocost = costs.BernoulliCost(y, lk), Cost("bernoulli", y, Link(link))
    elif cost_kind == "poisson":
        cost, ocost = costs.PoissonCost(y, lk), Cost("poisson", y, Link(link))
    elif cost_kind == "student_t":
        cost, ocost = costs.StudentTCost(5.0, y, lk, scale=0.8), Cost("student_t", y, Link(link), degrees_of_freedom=5.0, scale=0.8)
    else:
        cost = costs.MultiModalCost(0.5, 1.2, 0.35, y, lk)
        ocost = Cost("multimodal", y, Link(link), observation_noise=0.5, shift=1.2, bernoulli_noise=0.35)
    return pkg.PLS(basis, cost), PLSOracle(orc_basis, ocost)


@pytest.mark.parametrize(
    "n,d,m,j,cost_kind,link",
    [
        (1000, 1, 64, 128, "bernoulli", "sigmoid"),  # 1-D, exact tile multiples in J
        (2500, 8, 200, 300, "gaussian", "identity"),  # ragged N, M, J
        (777, 3, 33, 131, "poisson", "square"),  # odd J (padded leading dimension), tiny M
        (1500, 16, 150, 64, "student_t", "identity"),  # D = 16 (five exponent k-steps)
        (900, 5, 70, 50, "multimodal", "identity"),
        (640, 2, 40, 96, "bernoulli", "probit"),
        (129, 10, 17, 7, "gaussian", "square"),
    ],
)
def test_langevin_step_matches_oracle(b200, n, d, m, j, cost_kind, link):
    x, y, z, ls, g = _problem(n, d, m, j, seed=n + j, cost_kind=cost_kind, link=link)
    pls, orc = _build_pair(b200, x, y, z, ls, 1.4, cost_kind, link)
    m_k = orc.basis.approximation_dimension
    assert pls.basis.approximation_dimension == m_k
    p = 0.5 * torch.randn(m_k, j, generator=g, dtype=torch.float64)
    if cost_kind == "poisson":
        # -2y/F is singular at F = 0 and amplifies the 1e-13 re-association noise of F without bound; fit the particles
        # so that F stays near 2 (least squares on the oracle's dense features) to keep the comparison well conditioned
        phi = orc.basis.k_zx.T @ orc.basis.scaled_eigenvectors
        target = 2.0 + 0.2 * torch.randn(n, j, generator=g, dtype=torch.float64)
        p = torch.linalg.lstsq(phi, target).solution.contiguous()
        assert orc.basis.forward(p).abs().min() > 0.2
    xi = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    pc = p.cuda()
    assert rel_err(pls.basis.calculate_untransformed_train_prediction_samples(pc), orc.basis.forward(p)) < TOL
    assert rel_err(pls.calculate_cost_derivative(pc), orc.calculate_cost_derivative(p)) < TOL
    assert rel_err(pls.calculate_cost(pc), orc.calculate_cost(p)) < TOL
    want = orc.calculate_particle_update(p, 1e-3, noise=xi)
    assert rel_err(pls.calculate_particle_update(pc, 1e-3, noise=xi), want) < TOL
    e_want = orc.calculate_energy_potential(p)
    assert abs(pls.calculate_energy_potential(pc) - e_want) <= TOL * abs(e_want)
    # in-place step == particles + delta
    q = pc.clone()
    pls.step_(q, 1e-3, noise=xi)
    assert rel_err(q, p + want) < TOL


def test_row_chunked_gradient_matches_unchunked(b200):
    """The Dc row-chunk loop (accumulating back-projection) gives the same update as one chunk."""
    x, y, z, ls, g = _problem(3000, 4, 96, 200, seed=5)
    pls_a, orc = _build_pair(b200, x, y, z, ls, 1.0, "gaussian", "identity")
    pls_b, _ = _build_pair(b200, x, y, z, ls, 1.0, "gaussian", "identity", dc_budget_bytes=512 * 200 * 8)  # 512-row chunks
    m_k = orc.basis.approximation_dimension
    p = torch.randn(m_k, 200, generator=g, dtype=torch.float64)
    xi = torch.randn(m_k, 200, generator=g, dtype=torch.float64)
    want = orc.calculate_particle_update(p, 5e-4, noise=xi)
    da = pls_a.calculate_particle_update(p.cuda(), 5e-4, noise=xi)
    db = pls_b.calculate_particle_update(p.cuda(), 5e-4, noise=xi)
    assert len(pls_b.basis.engine(200).chunks) > 1
    assert rel_err(da, want) < TOL and rel_err(db, want) < TOL


def test_trajectory_matches_oracle(b200):
    """20 steps with the reference's host noise stream replayed on both sides."""
    torch.set_default_dtype(torch.float64)
    try:
        x, y, z, ls, g = _problem(1200, 2, 48, 160, seed=21)
        pls, orc = _build_pair(b200, x, y, z, ls, 2.0, "gaussian", "identity")
        p0 = torch.randn(orc.basis.approximation_dimension, 160, generator=g, dtype=torch.float64)
        torch.manual_seed(77)
        want = orc.run(p0, 1e-3, 20)
        torch.manual_seed(77)
        p = p0.cuda()
        for _ in range(20):
            p += pls.calculate_particle_update(p, 1e-3)
        assert rel_err(p, want) < 1e-9
    finally:
        torch.set_default_dtype(torch.float32)


def test_linearity_of_forward_at_scale(b200):
    """Size-independent property at a larger shape: F(aP + bQ) = aF(P) + bF(Q), and the energy of the Gaussian cost is
    consistent with the materialised prediction."""
    x, y, z, ls, g = _problem(20000, 8, 256, 512, seed=3)
    pls, orc = _build_pair(b200, x, y, z, ls, 1.0, "gaussian", "identity", threshold=1e-6)
    m_k = pls.basis.approximation_dimension
    p = torch.randn(m_k, 512, generator=g, dtype=torch.float64).cuda()
    q = torch.randn(m_k, 512, generator=g, dtype=torch.float64).cuda()
    fwd = pls.basis.calculate_untransformed_train_prediction_samples
    lhs = fwd(2.0 * p - 3.0 * q)
    rhs = 2.0 * fwd(p) - 3.0 * fwd(q)
    assert rel_err(lhs, rhs) < 1e-11
    cost_direct = pls.calculate_cost(p)
    cost_from_f = ((fwd(p) - y.cuda()[:, None]) ** 2).sum(0) / (2 * 0.25)
    assert rel_err(cost_direct, cost_from_f) < 1e-11
    # spot-check 64 random rows of F against the oracle's dense product
    rows = torch.randperm(20000, generator=g)[:64]
    want = (orc.basis.k_zx.T[rows] @ orc.basis.scaled_eigenvectors) @ p.cpu()
    assert rel_err(fwd(p)[rows.cuda()], want) < TOL


def test_selector_matches_oracle_random(b200):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4000, 6, generator=g, dtype=torch.float64)
    ls = torch.tensor([1.5, 2.0, 2.5, 1.8, 2.2, 3.0], dtype=torch.float64)
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=6, lengthscale=ls), outputscale=1.7)
    b200.set_seed(4)
    z, idx = b200.ConditionalVarianceInducingPointSelector()(x=x, m=96, kernel=kernel)
    oracle_set_seed(4)
    (zo, idxo, trace) = conditional_variance_select(x, 96, RBFScaleKernel(ls, 1.7), return_trace=True)
    assert trace["min_top2_rel_gap"] > 1e-9  # the comparison is meaningful: no near-ties in this run
    assert idx.tolist() == idxo.tolist()
    assert torch.equal(z, zo)


@pytest.mark.parametrize("rule", ["numpy", "stable"])
def test_selector_ties_duplicated_rows(b200, rule):
    """Exact ties from duplicated rows.  Default rule: the reference's `reversed(np.argsort(d))`, called on the host for the
    tied iterations only (= the oracle's literal restatement).  "stable": the highest permuted index, on the device (= the
    oracle with a stable sort)."""
    g = torch.Generator().manual_seed(2)
    base = torch.randn(40, 2, generator=g, dtype=torch.float64)
    x = torch.cat([base, base, base], dim=0)
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=2, lengthscale=1.0), outputscale=1.0)
    b200.set_seed(1)
    sel = b200.ConditionalVarianceInducingPointSelector(tie_rule=rule)
    _, idx = sel(x=x, m=12, kernel=kernel)
    oracle_set_seed(1)
    _, idxo = conditional_variance_select(x, 12, RBFScaleKernel(1.0, 1.0), argsort_kind=None if rule == "numpy" else "stable")
    assert idx.tolist() == idxo.tolist()
    assert sel.last_run_info["tied_picks"] + sel.last_run_info["host_tie_calls"] >= 1
    assert (sel.last_run_info["host_tie_calls"] == 0) == (rule == "stable")


def test_selector_custom_tie_rule_and_bad_rule(b200):
    """A callable tie rule is honoured; one that returns an already chosen point is rejected loudly."""
    x = torch.linspace(-1, 1, 200, dtype=torch.float64)[:, None]
    kernel = b200.ScaleKernel(b200.RBFKernel(lengthscale=0.05), outputscale=1.0)
    b200.set_seed(3)
    lowest = lambda d, chosen: int(np.flatnonzero((d == np.delete(d, chosen).max()) & ~np.isin(np.arange(d.size), chosen))[0])  # noqa: E731
    _, idx_low = b200.ConditionalVarianceInducingPointSelector(tie_rule=lowest)(x=x, m=6, kernel=kernel)
    b200.set_seed(3)
    _, idx_high = b200.ConditionalVarianceInducingPointSelector(tie_rule="stable")(x=x, m=6, kernel=kernel)
    assert idx_low.tolist() != idx_high.tolist() and len(set(idx_low.tolist())) == 6
    b200.set_seed(3)
    with pytest.raises(Exception, match="already chosen|out of range"):
        b200.ConditionalVarianceInducingPointSelector(tie_rule=lambda d, chosen: int(chosen[0]))(x=x, m=6, kernel=kernel)
    with pytest.raises(ValueError):
        b200.ConditionalVarianceInducingPointSelector(tie_rule="random")(x=x, m=6, kernel=kernel)


def test_selector_early_stop_raises_like_reference(b200):
    x = torch.tensor([[1.0, 3.0], [3.0, 5.0], [1.1, 3.5], [1.3, 7.5], [2.5, 2.5]])
    b200.set_seed(0)
    sel = b200.ConditionalVarianceInducingPointSelector(threshold=1e9)
    with pytest.raises(IndexError):
        sel.compute_induce_data(x=x, m=4, kernel=b200.LinearKernel())
    with pytest.raises(AssertionError):
        sel.compute_induce_data(x=x, m=1, kernel=b200.LinearKernel())


def test_philox_noise_is_shard_invariant_and_standard_normal(b200):
    from projected_langevin_sampling_b200 import _native, ops

    ctx = _native.context()
    full = ops.philox_normal(ctx, seed=123, step=7, rows=64, j=4096)
    left = ops.philox_normal(ctx, seed=123, step=7, rows=64, j=1000, j_global_offset=0)
    right = ops.philox_normal(ctx, seed=123, step=7, rows=64, j=3096, j_global_offset=1000)
    assert torch.equal(full, torch.cat([left, right], dim=1))
    other = ops.philox_normal(ctx, seed=123, step=8, rows=64, j=4096)
    assert not torch.equal(full, other)
    assert abs(full.mean().item()) < 0.01 and abs(full.var().item() - 1.0) < 0.02
    assert abs((full**4).mean().item() - 3.0) < 0.1


def test_philox_step_equals_given_noise(b200):
    from projected_langevin_sampling_b200 import _native, ops

    x, y, z, ls, g = _problem(600, 3, 40, 90, seed=8)
    pls, orc = _build_pair(b200, x, y, z, ls, 1.0, "gaussian", "identity")
    m_k = orc.basis.approximation_dimension
    p = torch.randn(m_k, 90, generator=g, dtype=torch.float64)
    xi = ops.philox_normal(_native.context(), seed=5, step=3, rows=m_k, j=90, j_global_offset=10)
    q = p.cuda()
    pls.step_(q, 2e-3, philox=(5, 3, 10))
    want = p + orc.calculate_particle_update(p, 2e-3, noise=xi.cpu())
    assert rel_err(q, want) < TOL


def test_error_behaviour_matches_reference(b200):
    basis = linear_onb(b200)
    costs, links = _costs_mod()
    pls = b200.PLS(basis, costs.GaussianCost(1.0, torch.zeros(5), links.IdentityLinkFunction()))
    with pytest.raises(AssertionError):  # basis/base.py:156-158
        pls.calculate_particle_update(torch.zeros(3, 4), 1e-3)
    with pytest.raises(AssertionError):  # projected_langevin_sampling.py:131-133
        pls.calculate_energy_potential(torch.zeros(3, 4))
    with pytest.raises(ValueError):  # orthonormal.py:91-92
        pls.initialise_particles(4, noise_only=False)

    class Matern:  # no CUDA path and no CPU fallback
        pass

    with pytest.raises(TypeError):
        b200.OrthonormalBasis(b200.PLSKernel(Matern(), Z2), Z2, X5)


# ---- the caller of the hot path: train_pls with the energy fused into the step's forward ---------------------------------
@pytest.mark.parametrize("name,kind", [("full", "gaussian"), ("stopped", "student_t")])
def test_train_pls_against_reference_run(b200, name, kind, golden_dir):
    """projected_langevin_sampling_b200.trainers.train_pls (one forward per epoch: cost derivative and energy from the same
    F tiles) against the reference's own experiments/trainers.py:139-162 run (tests/golden/make_golden.py)."""
    from projected_langevin_sampling_b200.trainers import train_pls

    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = np.load(os.path.join(golden_dir, "train_loop_runs.npz"))
        x, z = torch.from_numpy(g["x"]), torch.from_numpy(g["z"])
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=2, lengthscale=torch.from_numpy(g["lengthscale"])),
                                  outputscale=float(g["outputscale"]))
        eig = (torch.from_numpy(g["eigenvalues"]), torch.from_numpy(g["eigenvectors"]))
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=float(g["threshold"]),
                                      eigendecomposition=eig, verbose=False)
        cost = (costs.GaussianCost(observation_noise=0.2, y_train=torch.from_numpy(g["y"]), link_function=links.IdentityLinkFunction())
                if kind == "gaussian" else costs.StudentTCost(degrees_of_freedom=4.0, y_train=torch.from_numpy(g["y"]),
                                                              link_function=links.IdentityLinkFunction(), scale=0.5))
        pls = b200.PLS(basis, cost)
        for on_device in (True, False):  # in place on a CUDA tensor; copied back into a host tensor
            p0 = torch.from_numpy(g[name + "__p0"]).clone()
            p0 = p0.cuda() if on_device else p0
            torch.manual_seed(int(g[name + "__noise_seed"]))
            p, energies = train_pls(pls, p0, int(g[name + "__epochs"]), float(g[name + "__eta"]), float(g[name + "__patience"]))
            assert p is p0
            want_e = g[name + "__energies"]
            assert len(energies) == len(want_e)  # same epochs accepted, same stopping epoch
            # one step agrees to 1e-10; 1e-9 over the trajectories (40 and 32 epochs) leaves room for round-off growth
            traj_tol = 1e-9
            assert abs(energies[0] - want_e[0]) <= 1e-10 * abs(want_e[0])
            assert np.allclose(energies, want_e, rtol=traj_tol)
            assert rel_err(p, torch.from_numpy(g[name + "__p"])) < traj_tol
    finally:
        torch.set_default_dtype(torch.float32)


def test_train_pls_matches_stepwise_api(b200):
    """Fused loop == the reference-shaped loop over calculate_particle_update + calculate_energy_potential, at a size with
    several row tiles, a ragged J and both Dc chunks (row-chunked gradient and cost partials)."""
    from projected_langevin_sampling_b200.trainers import train_pls

    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = torch.Generator().manual_seed(8)
        n, d, m, j = 700, 3, 40, 37
        x = torch.randn(n, d, generator=g)
        y = (torch.rand(n, generator=g) > 0.4).double()
        z = x[:m].clone()
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=torch.tensor([1.0, 1.4, 0.9])), outputscale=1.1)
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=1e-8, verbose=False,
                                      dc_budget_bytes=256 * 38 * 8)  # 256-row chunks -> 3 chunks
        pls = b200.PLS(basis, costs.BernoulliCost(y_train=y, link_function=links.SigmoidLinkFunction()))
        p0 = pls.initialise_particles(number_of_particles=j, seed=4)
        torch.manual_seed(31)
        p_ref, e_ref = p0.clone(), []
        for _ in range(6):
            p_ref += pls.calculate_particle_update(p_ref, 1e-3)
            e_ref.append(pls.calculate_energy_potential(p_ref))
        torch.manual_seed(31)
        p, energies = train_pls(pls, p0.clone(), 6, 1e-3, early_stopper_patience=10.0)
        assert np.allclose(energies, e_ref, rtol=1e-11)
        assert rel_err(p, p_ref) < 1e-11
    finally:
        torch.set_default_dtype(torch.float32)


def test_cuda_graph_run_equals_eager_run(b200):
    """PLS.run(cuda_graph=True): one captured step replayed with a device-side Philox step counter gives exactly the eager
    loop's particles (same kernels, same arguments), at a launch-bound size."""
    x, y, z, ls, g = _problem(2000, 1, 32, 256, seed=12, cost_kind="bernoulli")
    pls, _ = _build_pair(b200, x, y, z, ls, 1.0, "bernoulli", "sigmoid")
    p0 = torch.randn(pls.basis.approximation_dimension, 256, generator=g, dtype=torch.float64).cuda()
    eager, _ = pls.run(p0.clone(), 1e-3, 25, seed=5, j_global_offset=64)
    graphed, energies = pls.run(p0.clone(), 1e-3, 25, seed=5, j_global_offset=64, cuda_graph=True, energy_every=10)
    assert torch.equal(eager, graphed)
    assert len(energies) == 2 and all(np.isfinite(energies))
    again, _ = pls.run(p0.clone(), 1e-3, 25, seed=5, j_global_offset=64)  # the counter hook is uninstalled afterwards
    assert torch.equal(eager, again)
    with pytest.raises(ValueError):
        pls.run(p0.clone(), 1e-3, 2, cuda_graph=True)


def test_predict_untransformed_samples_matches_oracle(b200):
    """Prediction at new inputs with the predictive noise given (orthonormal.py:216-244): RBF/ARD kernel, ragged sizes."""
    x, y, z, ls, g = _problem(900, 4, 50, 33, seed=14)
    pls, orc = _build_pair(b200, x, y, z, ls, 1.3, "gaussian", "identity")
    m_k = orc.basis.approximation_dimension
    xs = torch.randn(257, 4, generator=g, dtype=torch.float64)
    p = torch.randn(m_k, 33, generator=g, dtype=torch.float64)
    noise = 0.1 * torch.randn(m_k + 257, 33, generator=g, dtype=torch.float64)
    want = orc.basis.predict_untransformed_samples(p, xs, noise=noise)
    got = pls.predict_untransformed_samples(p.cuda(), xs, noise=noise)
    assert rel_err(got, want) < TOL


def test_step_size_search_runner_and_checkpoint(b200, tmp_path):
    """runners.train_pls_runner (experiments/runners.py:331-446, metric "loss") against the oracle's restatement: same best
    step size, same number of accepted epochs, same particles; then the reference's .pth checkpoint format round-trips."""
    from oracle.pls_oracle import train_pls_runner_oracle
    from projected_langevin_sampling_b200.runners import load_pls, save_pls, train_pls_runner

    torch.set_default_dtype(torch.float64)
    try:
        x, y, z, ls, g = _problem(400, 2, 12, 16, seed=33)
        pls, orc = _build_pair(b200, x, y, z, ls, 1.2, "gaussian", "identity", threshold=1e-6)
        p0 = torch.randn(orc.basis.approximation_dimension, 16, generator=g, dtype=torch.float64)
        kw = dict(simulation_duration=0.02, maximum_number_of_steps=40, early_stopper_patience=1.0, number_of_step_searches=4,
                  step_size_upper=5e-3, minimum_change_in_energy_potential=1e-9, seed=3)
        want_p, want_lr, want_n, want_hist = train_pls_runner_oracle(orc, p0.clone(), **kw)
        hist = {}
        got_p, got_lr, got_n = train_pls_runner(pls, p0.cuda(), energy_potentials_history=hist, **kw)
        assert got_lr == want_lr and got_n == want_n and list(hist) == list(want_hist)
        for step_size in hist:
            assert np.allclose(hist[step_size], want_hist[step_size], rtol=1e-9)
        assert rel_err(got_p, want_p) < 1e-9
        path = str(tmp_path / "pls.pth")
        save_pls(pls, got_p, path, best_lr=got_lr, number_of_epochs=got_n)
        pls.observation_noise = 123.0
        pls2, p2, lr2, n2 = load_pls(pls, path)
        assert torch.equal(p2, got_p) and lr2 == got_lr and n2 == got_n and pls2.observation_noise == 0.25
    finally:
        torch.set_default_dtype(torch.float32)


def test_step_size_search_runner_against_reference_run(b200, golden_dir):
    """runners.train_pls_runner against a run of the reference's OWN experiments/runners.py:331-446 (tests/golden/make_golden.py
    runner_runs; metric "loss", four log-spaced step sizes from the same particles and seed): the same best step size, the
    same number of accepted epochs and the same particles."""
    from projected_langevin_sampling_b200.runners import train_pls_runner

    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = np.load(os.path.join(golden_dir, "runner_runs.npz"))
        x, z = torch.from_numpy(g["x"]), torch.from_numpy(g["z"])
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=2, lengthscale=torch.from_numpy(g["lengthscale"])),
                                  outputscale=float(g["outputscale"]))
        eig = (torch.from_numpy(g["eigenvalues"]), torch.from_numpy(g["eigenvectors"]))
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=float(g["threshold"]),
                                      eigendecomposition=eig, verbose=False)
        pls = b200.PLS(basis, costs.GaussianCost(observation_noise=float(g["observation_noise"]), y_train=torch.from_numpy(g["y"]),
                                                 link_function=links.IdentityLinkFunction()))
        kw = {k[4:]: (int(g[k]) if k[4:] in ("maximum_number_of_steps", "number_of_step_searches", "seed") else float(g[k]))
              for k in g.files if k.startswith("kw__")}
        hist = {}
        got_p, got_lr, got_n = train_pls_runner(pls, torch.from_numpy(g["p0"]).cuda(), energy_potentials_history=hist, **kw)
        assert got_lr == float(g["best_lr"]) and got_n == int(g["epochs"])
        assert len(hist) == kw["number_of_step_searches"]  # every search ran (no early break in the reference run either)
        assert rel_err(got_p, torch.from_numpy(g["particles"])) < 1e-9  # 160 epochs of round-off growth on a 1e-10 step
    finally:
        torch.set_default_dtype(torch.float32)


# ---- InducingPointBasis ("next" row) ---------------------------------------------------------------------------------------
def test_ipb_reference_vectors(b200):  # reference tests/test_basis.py:98-116,248-269,369-387,495-519 (linear mock kernel)
    y2 = torch.tensor([2.1, 3.3])
    f_want = torch.tensor([[1.2656511068, -0.8267806172, -2.1464431286], [-6.1349906921, -5.1450047493, 4.8272337914],
                           [-8.9970912933, -5.7714910507, 8.1600494385], [1.5409765244, -0.2934455872, -2.1787719727],
                           [0.5684509277, -1.0845184326, -1.3986053467]])
    basis = b200.InducingPointBasis(b200.PLSKernel(b200.LinearKernel(), Z2), Z2, y2, X5)
    assert basis.approximation_dimension == 2
    assert torch.allclose(basis.initialise_particles(3, seed=0).cpu().float(), P23)
    assert torch.allclose(basis.initialise_particles(3, seed=0, noise_only=False).cpu().float(), y2[:, None] + P23)
    assert torch.allclose(basis.calculate_untransformed_train_prediction_samples(P23).cpu().float(), f_want, rtol=1e-3, atol=1e-3)
    assert np.isclose(basis.calculate_energy_potential(P23, torch.ones(3)), 275.2294006347656, rtol=1e-3)


def test_ipb_against_reference_run(b200, golden_dir):
    """The CUDA InducingPointBasis against a run of the reference's own class (tests/golden/make_golden.py ipb_runs):
    forward, cost derivative, cost, energy and one Langevin update with the reference's noise stream replayed."""
    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = np.load(os.path.join(golden_dir, "ipb_runs.npz"))
        x, z, y = torch.from_numpy(g["x"]), torch.from_numpy(g["z"]), torch.from_numpy(g["y"])
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=2, lengthscale=torch.from_numpy(g["lengthscale"])), outputscale=float(g["outputscale"]))
        basis = b200.InducingPointBasis(b200.PLSKernel(kernel, z), z, torch.from_numpy(g["y_induce"]), x)
        pls = b200.PLS(basis, costs.GaussianCost(float(g["observation_noise"]), y, links.IdentityLinkFunction()))
        p = pls.initialise_particles(number_of_particles=5, seed=5, noise_only=False)
        assert rel_err(p, torch.from_numpy(g["p"])) < 1e-14
        assert rel_err(basis.calculate_untransformed_train_prediction_samples(p), torch.from_numpy(g["f"])) < TOL
        assert rel_err(pls.calculate_cost_derivative(p), torch.from_numpy(g["dc"])) < TOL
        assert rel_err(pls.calculate_cost(p), torch.from_numpy(g["cost"])) < TOL
        assert abs(pls.calculate_energy_potential(p) - float(g["energy"])) <= TOL * abs(float(g["energy"]))
        torch.manual_seed(int(g["noise_seed"]))
        assert rel_err(pls.calculate_particle_update(p, float(g["step_size"])), torch.from_numpy(g["delta"])) < TOL
        # the unfused composition (materialised cost derivative) gives the same update
        torch.manual_seed(int(g["noise_seed"]))
        dc = pls.cost.calculate_cost_derivative(basis.calculate_untransformed_train_prediction_samples(p))
        assert rel_err(basis.calculate_particle_update(p, dc, float(g["step_size"])), torch.from_numpy(g["delta"])) < TOL
    finally:
        torch.set_default_dtype(torch.float32)


def test_ipb_step_matches_oracle_at_size(b200):
    """InducingPointBasis at a size with several row tiles and Dc chunks, against the oracle (well-separated inducing points
    keep k(Z, Z) well conditioned, so the two Cholesky solves agree far below the tolerance)."""
    from oracle.pls_oracle import InducingPointBasisOracle

    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = torch.Generator().manual_seed(51)
        n, d, m, j = 1500, 2, 36, 70
        x = 4 * torch.rand(n, d, generator=g, dtype=torch.float64) - 2
        gx, gy = torch.meshgrid(torch.linspace(-2, 2, 6), torch.linspace(-2, 2, 6), indexing="ij")
        z = torch.stack([gx.reshape(-1), gy.reshape(-1)], dim=1).double()
        ls = torch.tensor([0.5, 0.6], dtype=torch.float64)
        y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
        y_induce = torch.sin(z.sum(1))
        orc_basis = InducingPointBasisOracle(RBFScaleKernel(ls, 1.1), z, y_induce, x)
        assert torch.linalg.cond(orc_basis.base_gram_induce) < 1e3
        orc = PLSOracle(orc_basis, Cost("student_t", y, Link("identity"), degrees_of_freedom=5.0, scale=0.8))
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=ls), outputscale=1.1)
        p = y_induce[:, None] + 0.3 * torch.randn(m, j, generator=g, dtype=torch.float64)
        zn = torch.randn(m, j, generator=g, dtype=torch.float64)
        want = orc.calculate_particle_update(p, 2e-3, noise=zn)
        e_want = orc.calculate_energy_potential(p)
        for gram_cache in (False, True):  # Gram tiles generated in the kernels / k(X, Z) kept in HBM and streamed
            basis = b200.InducingPointBasis(b200.PLSKernel(kernel, z), z, y_induce, x, dc_budget_bytes=512 * 70 * 8, gram_cache=gram_cache)
            pls = b200.PLS(basis, costs.StudentTCost(5.0, y, links.IdentityLinkFunction(), scale=0.8))
            assert rel_err(pls.calculate_particle_update(p.cuda(), 2e-3, noise=zn), want) < TOL
            assert len(basis.engine(j).chunks) > 1 and (basis.engine(j).gram is not None) == gram_cache
            assert abs(pls.calculate_energy_potential(p.cuda()) - e_want) <= TOL * abs(e_want)
            q = p.cuda().clone()
            basis.fused_particle_update(q, pls.cost, 2e-3, noise=zn, in_place=True)
            assert rel_err(q, p + want) < TOL
    finally:
        torch.set_default_dtype(torch.float32)


def test_ipb_fused_training_loop_and_philox(b200):
    """InducingPointBasis through train_pls: the fused epoch (energy and gradient from one forward, the prior term on the
    W = k(Z, Z)^{-1} P that pass formed) equals the reference-shaped loop over calculate_particle_update +
    calculate_energy_potential on the reference's noise stream; with the device-side Philox stream the run does not depend on
    how the particles are split (the coloured noise is a column-wise product of V sqrt(lambda) with keyed normals)."""
    from projected_langevin_sampling_b200.trainers import train_pls

    torch.set_default_dtype(torch.float64)
    try:
        costs, links = _costs_mod()
        g = torch.Generator().manual_seed(61)
        n, d, m, j = 900, 2, 25, 40
        x = 4 * torch.rand(n, d, generator=g) - 2
        gx, gy = torch.meshgrid(torch.linspace(-2, 2, 5), torch.linspace(-2, 2, 5), indexing="ij")
        z = torch.stack([gx.reshape(-1), gy.reshape(-1)], dim=1).double()
        y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g)
        y_induce = torch.sin(z.sum(1))
        kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=torch.tensor([0.6, 0.7])), outputscale=1.2)
        basis = b200.InducingPointBasis(b200.PLSKernel(kernel, z), z, y_induce, x, dc_budget_bytes=256 * 40 * 8)
        pls = b200.PLS(basis, costs.GaussianCost(observation_noise=0.3, y_train=y, link_function=links.IdentityLinkFunction()))
        p0 = pls.initialise_particles(number_of_particles=j, seed=2, noise_only=False)
        torch.manual_seed(9)
        p_ref, e_ref = p0.clone().cuda(), []
        for _ in range(5):
            p_ref += pls.calculate_particle_update(p_ref, 5e-4)
            e_ref.append(pls.calculate_energy_potential(p_ref))
        torch.manual_seed(9)
        p, energies = train_pls(pls, p0.clone().cuda(), 5, 5e-4, early_stopper_patience=10.0)
        assert np.allclose(energies, e_ref, rtol=1e-11)
        assert rel_err(p, p_ref) < 1e-11
        # Philox: whole particle set vs two particle shards advanced separately
        whole, e_whole = train_pls(pls, p0.clone().cuda(), 4, 5e-4, early_stopper_patience=10.0, philox_seed=123)
        parts = []
        for j0, j1 in ((0, 16), (16, j)):
            part, _ = train_pls(pls, p0[:, j0:j1].clone().cuda(), 4, 5e-4, early_stopper_patience=10.0, philox_seed=123, j_global_offset=j0)
            parts.append(part)
        assert rel_err(torch.cat(parts, dim=1), whole) < 1e-12 and len(e_whole) == 4 and all(np.isfinite(e_whole))
        assert not torch.equal(whole.cpu(), p.cpu())  # a different noise stream
        q = p0.clone().cuda()
        pls.step_(q, 5e-4, philox=(123, 0, 0))
        again = p0.clone().cuda()
        pls.step_(again, 5e-4, philox=(123, 0, 0))
        assert torch.equal(q, again) and bool(torch.isfinite(q).all())
    finally:
        torch.set_default_dtype(torch.float32)


def _mock_r_kernel(pkg, samples):
    """The reference tests' MockProjectedLangevinSamplingKernel (mockers/kernel.py:26-43): its forward is the base kernel."""
    from projected_langevin_sampling_b200.kernels import dense_gram

    class MockPLSKernel(pkg.PLSKernel):
        def forward(self, x1, x2, additional_approximation_samples=None, **kw):
            return dense_gram(self.base_kernel, x1, x2)

    return MockPLSKernel(pkg.LinearKernel(), samples)


def test_reference_predictive_noise_vectors(b200):
    """sample_predictive_noise of both bases against the reference's golden draws (tests/test_basis.py:522-635 ONB, :640-838
    IPB; seed 0, float32 default dtype as in the reference's test run) and IPB predict with the noise given (:862-1010)."""
    xs = torch.tensor([[3.0, 2.0, 3.2], [1.5, 6.5, 1.5]])
    onb_want = torch.tensor([[0.0851, -0.1569, -0.2067], [3.1662, 4.6236, -1.2954], [1.3697, 1.4171, 0.7368], [3.9759, 6.4164, -2.9854]])
    ipb_want = torch.tensor([[1.4442, 3.7593, -0.4158], [1.9489, 4.2264, -0.9286], [3.0377, 3.1129, -3.4442], [1.4840, 1.3103, 0.6729]])
    onb = b200.OrthonormalBasis(_mock_r_kernel(b200, Z2), Z2, X5, verbose=False)
    b200.set_seed(0)
    assert torch.allclose(onb.sample_predictive_noise(P23, xs).cpu().float(), onb_want, rtol=1e-3, atol=2e-4)
    ipb = b200.InducingPointBasis(_mock_r_kernel(b200, Z2), Z2, torch.tensor([2.1, 3.3]), X5)
    b200.set_seed(0)
    # rank-deficient 4 x 4 covariance: the round-off eigenvalue of the float32 golden run contributes ~1e-3 of noise
    assert torch.allclose(ipb.sample_predictive_noise(P23, xs).cpu().float(), ipb_want, rtol=1e-3, atol=1e-2)
    pred = ipb.predict_untransformed_samples(P23, xs, noise=ipb_want)
    assert torch.allclose(pred.cpu().float(), torch.tensor([[-4.4373, -3.6672, 2.9305], [-7.8718, -6.6582, 8.8616]]), rtol=2e-3, atol=2e-3)
