"""GPU parity at the shapes BASELINE.json names, through the public API / C ABI.

C2 (N=10k, M=64, J=1024, Bernoulli) is checked in full against the oracle, including the selector.  C3 (N=100k, M=256,
J=4096, Poisson f^2) and C4 (N=1M, D=8, M=1024, J=4096, Gaussian) are checked at FULL N and M on a slice of the particles
(every term of the step is column-wise, orthonormal.py:151-158, so a slice of columns is the same computation) with the
oracle's dense algebra evaluated in row chunks, plus size-independent properties on the full particle set.  Every config runs
with the Gram generated inside the kernels, cached in HBM, and staged chunk by chunk (OrthonormalBasis(gram_cache=...)).
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.pls_oracle import (  # noqa: E402
    Cost,
    Link,
    OrthonormalBasisOracle,
    PLSOracle,
    RBFScaleKernel,
    conditional_variance_select,
    set_seed as oracle_set_seed,
)

TOL = 1e-10


def rel_err(got, want):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    scale = want.abs().max().item()
    return (got - want).abs().max().item() / (scale if scale > 0 else 1.0)


@pytest.fixture(scope="module")
def b200():
    import projected_langevin_sampling_b200 as pkg
    from projected_langevin_sampling_b200 import _native

    _native.context()
    return pkg


def _mods():
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions

    return costs, link_functions


def _curve_inputs(n, kind, seed=0):
    """bench.py's synthetic inputs for the 1-D configs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(-3, 3, n, dtype=torch.float64).reshape(-1, 1)
    curve = 2.0 * torch.sin(1.5 * x.reshape(-1))
    y = torch.bernoulli(torch.sigmoid(curve), generator=g).double() if kind == "bernoulli" else torch.poisson(curve**2, generator=g).double()
    return x, y, g


@pytest.mark.parametrize("gram_cache", [False, True, "staged"], ids=["generated", "cached", "staged"])
def test_config2_bernoulli_full(b200, gram_cache):
    """C2: N=10 000, D=1, M=64, J=1024, BernoulliCost + Sigmoid.  The selector is compared for the first 24 pivots: a 1-D RBF
    Gram with lengthscale 0.5 on [-3, 3] has numerical rank ~30, beyond which every conditional variance is round-off of the
    1e-12 jitter and the pivot order is noise in the reference as well; the step uses bench.py's evenly spaced inducing points."""
    costs, links = _mods()
    n, m, j = 10_000, 64, 1024
    x, y, g = _curve_inputs(n, "bernoulli")
    orc_kernel = RBFScaleKernel(torch.tensor([0.5], dtype=torch.float64), 1.0)
    kernel = b200.ScaleKernel(b200.RBFKernel(lengthscale=0.5), outputscale=1.0)
    # selector: linspace inputs give exact ties; the selector resolves them as the reference does (numpy's default argsort on
    # the host, for the tied iterations only), so it is compared with the oracle's literal restatement of :105-109
    b200.set_seed(0)
    z_sel, idx = b200.ConditionalVarianceInducingPointSelector()(x=x, m=24, kernel=kernel)
    oracle_set_seed(0)
    z_orc, idx_orc, trace = conditional_variance_select(x, 24, orc_kernel, return_trace=True)
    assert trace["di"].max() > 1e-9  # still above the round-off floor after 24 pivots
    assert idx.tolist() == idx_orc.tolist() and torch.equal(z_sel, z_orc)
    z = x[torch.linspace(0, n - 1, m).long()].clone()
    eig = torch.linalg.eigh((1 / m) * orc_kernel(z, z))
    orc = PLSOracle(OrthonormalBasisOracle(orc_kernel, z, x, eigenvalue_threshold=1e-10, eig=eig), Cost("bernoulli", y, Link("sigmoid")))
    pls = b200.PLS(b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=1e-10, eigendecomposition=eig, verbose=False,
                                         gram_cache=gram_cache),
                   costs.BernoulliCost(y, links.SigmoidLinkFunction()))
    m_k = orc.basis.approximation_dimension
    assert pls.basis.approximation_dimension == m_k
    p = 0.3 * torch.randn(m_k, j, generator=g, dtype=torch.float64)
    xi = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    assert rel_err(pls.calculate_cost_derivative(p.cuda()), orc.calculate_cost_derivative(p)) < TOL
    want = orc.calculate_particle_update(p, 1e-4, noise=xi)
    assert rel_err(pls.calculate_particle_update(p.cuda(), 1e-4, noise=xi), want) < TOL
    e_want = orc.calculate_energy_potential(p)
    assert abs(pls.calculate_energy_potential(p.cuda()) - e_want) <= TOL * abs(e_want)


def _chunked_oracle_update(kernel, x, y, z, vt, lam, p, eta, xi, dcost, chunk=50_000):
    """The reference's update (orthonormal.py:151-158) with its dense N-sized products evaluated in row chunks on the CPU."""
    g = torch.zeros(z.shape[0], p.shape[1], dtype=torch.float64)
    w = vt @ p
    for r0 in range(0, x.shape[0], chunk):
        k = kernel(x[r0:r0 + chunk], z)  # (rows, M)
        f = k @ w
        g += k.T @ dcost(y[r0:r0 + chunk, None], f)
    return -eta * (vt.T @ g) - eta * (p / lam[:, None]) + math.sqrt(2 * eta) * xi, g


@pytest.mark.parametrize("gram_cache", [False, True, "staged"], ids=["generated", "cached", "staged"])
def test_config3_poisson_full_rows(b200, gram_cache):
    """C3: N=100 000, D=1, M=256, PoissonCost + Square: a 192-particle slice at full N and M against the oracle, and the
    full J=4096 step's slice against the slice's own step (column independence)."""
    costs, links = _mods()
    n, m, j_slice, j_full = 100_000, 256, 192, 4096
    x, y, g = _curve_inputs(n, "poisson")
    z = x[torch.linspace(0, n - 1, m).long()].clone()
    orc_kernel = RBFScaleKernel(torch.tensor([0.5], dtype=torch.float64), 1.0)
    kernel = b200.ScaleKernel(b200.RBFKernel(lengthscale=0.5), outputscale=1.0)
    lam, vec = torch.linalg.eigh((1 / m) * orc_kernel(z, z))
    keep = lam > 1e-12
    lam_k, vec_k = lam[keep], vec[:, keep]
    vt = vec_k / torch.sqrt(lam_k.shape[0] * lam_k)  # orthonormal.py:63-68
    basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=1e-12, eigendecomposition=(lam, vec), verbose=False,
                                  gram_cache=gram_cache)
    assert rel_err(basis.scaled_eigenvectors, vt) < 1e-13
    pls = b200.PLS(basis, costs.PoissonCost(y, links.SquareLinkFunction()))
    m_k = lam_k.shape[0]
    # particles whose predictions stay away from F = 0, where -2y/F is singular: F ~ 1.5 + small
    phi_rows = torch.randperm(n, generator=g)[:4000]
    phi = orc_kernel(x[phi_rows], z) @ vt
    p_full = torch.linalg.lstsq(phi, 1.5 + 0.05 * torch.randn(4000, j_full, generator=g, dtype=torch.float64)).solution.contiguous()
    p = p_full[:, :j_slice].contiguous()
    xi_full = torch.randn(m_k, j_full, generator=g, dtype=torch.float64)
    want, _ = _chunked_oracle_update(orc_kernel, x, y, z, vt, lam_k, p, 1e-6, xi_full[:, :j_slice],
                                     lambda yy, f: -2 * yy / f + 2 * f)  # costs/poisson.py:76-82
    got = pls.calculate_particle_update(p.cuda(), 1e-6, noise=xi_full[:, :j_slice].contiguous())
    assert rel_err(got, want) < TOL
    got_full = pls.calculate_particle_update(p_full.cuda(), 1e-6, noise=xi_full)
    assert rel_err(got_full[:, :j_slice], want) < TOL


@pytest.mark.parametrize("gram_cache", [False, True, "staged"], ids=["generated", "cached", "staged"])
def test_config4_gaussian_full_rows(b200, gram_cache):
    """C4: N=1 000 000, D=8 ARD, M=1024, GaussianCost: a 64-particle slice at full N and M against the oracle's dense algebra
    (row-chunked on the CPU), then properties of the full J=4096 step: its slice equals the slice's own step, the Philox
    noise does not depend on how J is sharded, and the fused energy equals the cost-only forward."""
    costs, links = _mods()
    n, d, m, j_slice, j_full = 1_000_000, 8, 1024, 64, 4096
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    ls = torch.tensor([math.sqrt(d) * (0.75 + 0.5 * k / (d - 1)) for k in range(d)], dtype=torch.float64)
    y = torch.sin(x.sum(1) / math.sqrt(d)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    z = x[:m].clone()
    orc_kernel = RBFScaleKernel(ls, 1.0)
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=ls), outputscale=1.0)
    lam, vec = torch.linalg.eigh((1 / m) * orc_kernel(z, z))
    keep = lam > 0.0
    lam_k, vec_k = lam[keep], vec[:, keep]
    vt = vec_k / torch.sqrt(lam_k.shape[0] * lam_k)
    basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigendecomposition=(lam, vec), verbose=False, gram_cache=gram_cache)
    pls = b200.PLS(basis, costs.GaussianCost(0.01, y, links.IdentityLinkFunction()))
    m_k = lam_k.shape[0]
    assert basis.approximation_dimension == m_k
    p = torch.randn(m_k, j_slice, generator=g, dtype=torch.float64)
    xi = torch.randn(m_k, j_slice, generator=g, dtype=torch.float64)
    eta = 1e-9
    want, g_want = _chunked_oracle_update(orc_kernel, x, y, z, vt, lam_k, p, eta, xi, lambda yy, f: (1 / 0.01) * (f - yy))
    eng = basis.engine(j_slice)
    g_got = eng.gradient(p.cuda(), pls.cost.native(), pls.cost.y_device()).clone()
    assert rel_err(g_got[:, :j_slice], g_want) < TOL
    got = pls.calculate_particle_update(p.cuda(), eta, noise=xi)
    assert rel_err(got, want) < TOL
    # full particle set: the slice's columns are the same computation
    p_full = torch.cat([p, torch.randn(m_k, j_full - j_slice, generator=g, dtype=torch.float64)], dim=1).cuda()
    xi_full = torch.cat([xi, torch.randn(m_k, j_full - j_slice, generator=g, dtype=torch.float64)], dim=1)
    got_full = pls.calculate_particle_update(p_full, eta, noise=xi_full)
    assert rel_err(got_full[:, :j_slice], want) < TOL
    # Philox stream: sharding J in two halves reproduces the unsharded step (same noise; the gradient may differ by
    # round-off because the number of N-splits depends on the particle count)
    q = p_full.clone()
    pls.step_(q, eta, philox=(11, 3, 0))
    lo, hi = p_full[:, : j_full // 2].contiguous(), p_full[:, j_full // 2:].contiguous()
    pls.step_(lo, eta, philox=(11, 3, 0))
    pls.step_(hi, eta, philox=(11, 3, j_full // 2))
    assert rel_err(lo, q[:, : j_full // 2]) < 1e-13 and rel_err(hi, q[:, j_full // 2:]) < 1e-13
    # fused energy (cost sums from the step's own forward) == cost-only forward
    eng = basis.engine(j_full)
    e_fused = eng.energy_and_gradient(p_full, pls.cost.native(), pls.cost.y_device()).mean().item()
    e_plain = pls.calculate_energy_potential(p_full)
    assert abs(e_fused - e_plain) <= 1e-12 * abs(e_plain)
    assert np.isfinite(e_plain)


@pytest.mark.parametrize("gram_cache", [False, True, "staged"], ids=["generated", "cached", "staged"])
def test_gaussian_normal_equations_shortcut(b200, gram_cache):
    """Opt-in Gaussian / identity shortcut (LangevinEngine._normal_equations: A' = k(Z,X)k(X,Z)/s and b' = k(Z,X)y/s formed once,
    every step in M x M algebra): update and energy against the oracle, then 25 in-place steps and the fused training epoch
    against the general path on a larger problem."""
    costs, links = _mods()
    from projected_langevin_sampling_b200.trainers import train_pls

    g = torch.Generator().manual_seed(5)
    n, d, m, j = 30_000, 4, 96, 300
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    z = x[:m].clone()
    ls = torch.tensor([1.6, 2.0, 2.4, 2.8], dtype=torch.float64)
    orc_kernel = RBFScaleKernel(ls, 1.2)
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=ls), outputscale=1.2)
    eig = torch.linalg.eigh((1 / m) * orc_kernel(z, z))

    def make(shortcut):
        basis = b200.OrthonormalBasis(b200.PLSKernel(kernel, z), z, x, eigenvalue_threshold=1e-10, eigendecomposition=eig, verbose=False,
                                      gram_cache=gram_cache, gaussian_normal_equations=shortcut)
        return b200.PLS(basis, costs.GaussianCost(0.05, y, links.IdentityLinkFunction()))

    fast, general = make(True), make(False)
    orc = PLSOracle(OrthonormalBasisOracle(orc_kernel, z, x, eigenvalue_threshold=1e-10, eig=eig), Cost("gaussian", y, Link("identity"), observation_noise=0.05))
    m_k = orc.basis.approximation_dimension
    p = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    xi = torch.randn(m_k, j, generator=g, dtype=torch.float64)
    want = orc.calculate_particle_update(p, 1e-6, noise=xi)
    assert rel_err(fast.calculate_particle_update(p.cuda(), 1e-6, noise=xi), want) < TOL
    assert fast.basis.engine(j)._neq is not None and general.basis.engine(j)._neq is None
    pf, pg = p.cuda().clone(), p.cuda().clone()
    for s in range(25):
        fast.step_(pf, 2e-6, philox=(3, s, 0))
        general.step_(pg, 2e-6, philox=(3, s, 0))
    assert rel_err(pf, pg) < TOL
    e_fast = fast.basis.engine(j).energy_and_gradient(pf, fast.cost.native(), fast.cost.y_device()).mean().item()
    e_want = orc.calculate_energy_potential(pg.cpu())
    assert abs(e_fast - e_want) <= TOL * abs(e_want)
    _, ef = train_pls(fast, pf.clone(), number_of_epochs=6, step_size=1e-6, early_stopper_patience=1e9, philox_seed=11)
    _, eg = train_pls(general, pf.clone(), number_of_epochs=6, step_size=1e-6, early_stopper_patience=1e9, philox_seed=11)
    assert len(ef) == len(eg) == 6 and max(abs(a - b) / abs(b) for a, b in zip(ef, eg)) < TOL


# ---- selector at C4-like scale against runs of the reference (tests/golden/make_golden.py::selector_scale) ----------------
def _scale_problem(n, d, seed):
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    ls = torch.tensor([math.sqrt(d) * (0.75 + 0.5 * k / max(d - 1, 1)) for k in range(d)], dtype=torch.float64)
    return x, ls


@pytest.mark.parametrize("tag", ["n100000_m256", "n1000000_m1024"])
def test_selector_at_scale_against_reference_run(b200, golden_dir, tag):
    """N = 100 000, M = 256 and the C4 shape N = 1 000 000, M = 1024 (D = 8 ARD): the indices of the unmodified reference run
    (numpy BLAS `np.dot(cj, ci[:i])` accumulation, torch CPU exp) against pls_cv_select_f64 (fixed-order FMA chain, CUDA exp) --
    bit-exact -- together with the smallest relative top-2 gap met on the way (how far the selection was from depending on
    round-off).  Then the row-sharded entry points (three uneven emulated ranks) on the same problem."""
    import os

    from projected_langevin_sampling_b200 import _native as nat, ops
    from projected_langevin_sampling_b200.kernels import kernel_spec

    g = np.load(os.path.join(golden_dir, "selector_scale.npz"))
    if tag + "_idx" not in g:
        pytest.skip(f"no reference run stored for {tag}")
    n, m = (int(v[1:]) for v in tag.split("_"))
    d = int(g[tag + "_d"])
    x, ls = _scale_problem(n, d, int(g[tag + "_data_seed"]))
    kernel = b200.ScaleKernel(b200.RBFKernel(ard_num_dims=d, lengthscale=ls), outputscale=float(g[tag + "_outputscale"]))
    b200.set_seed(int(g[tag + "_seed"]))
    sel = b200.ConditionalVarianceInducingPointSelector()
    z, idx = sel(x=x, m=m, kernel=kernel)
    want = g[tag + "_idx"]
    agree = int((idx.numpy() == want).sum())
    first_diff = int(np.argmax(idx.numpy() != want)) if agree < m else -1
    info = sel.last_run_info
    print(f"[selector {tag}] identical pivots {agree}/{m} (first difference at {first_diff}); min top-2 relative gap "
          f"{info['min_top2_rel_gap']:.3e}; tied picks {info['tied_picks']}; host tie calls {info['host_tie_calls']}")
    assert idx.tolist() == want.tolist()
    assert torch.equal(z, x[torch.from_numpy(want)])
    assert info["min_top2_rel_gap"] > 1e-12  # round-off of the two arithmetic orders is ~1e-15: the comparison is meaningful
    del sel, z
    torch.cuda.empty_cache()

    # row-sharded entry points, ranks emulated on this GPU: same permutation, three uneven shards
    ctx = nat.context()
    b200.set_seed(int(g[tag + "_seed"]))
    perm = np.random.permutation(n)
    spec = kernel_spec(kernel, d)
    centre = x.mean(dim=0).tolist()
    bounds = [0, n // 3 + 17, 2 * n // 3 - 5, n]
    states = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        xa = ops.prepare_points(ctx, spec.kernel_id, x[torch.from_numpy(perm[a:b])].cuda(), spec.inv_lengthscale, centre, 0.5 * spec.log_outputscale)
        states.append(ops.ShardedSelectorState(ctx, spec.kernel_id, xa, a, n, d, spec.outputscale, m, 1e-12, 0.0))
    local, nsel = ops.cv_select_sharded(states, lambda recs: torch.cat(recs), lambda sl: torch.cat(sl).cpu().numpy())
    assert nsel == m
    assert perm[local.cpu().numpy()].tolist() == want.tolist()
