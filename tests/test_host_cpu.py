"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/pls_b200.h declares, layout
helpers, host-side argument logic, the product never touches the oracle, and the multi-process partition logic under
gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pls_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from projected_langevin_sampling_b200 import _native

    names = _declared_functions()
    assert len(names) >= 18
    assert sorted(_native.SIGNATURES) == names  # the binding covers exactly the header
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert _native.load_library().pls_abi_version() == _native.ABI_VERSION


def test_layout_helpers():
    from projected_langevin_sampling_b200 import _native

    lib = _native.load_library()
    # smallest stride >= D + 2 that is 4 mod 8 (bank-conflict-free point tiles, 16-byte rows)
    assert [lib.pls_point_stride(d) for d in (1, 2, 3, 8, 10, 11, 16, 18, 19, 26)] == [4, 4, 12, 12, 12, 20, 20, 20, 28, 28]
    assert lib.pls_point_stride(0) == -1 and lib.pls_point_stride(27) == -1
    with pytest.raises(ValueError):
        _native.point_stride(40)
    # headline shape: 8 x 32 = 256 output tiles -> 37 splits = 9472 CTAs = 64 whole waves of 148
    s = lib.pls_backward_splits(None, 1_000_000, 1024, 4096)
    assert (256 * s) % 148 == 0 and s >= 4
    assert lib.pls_backward_splits(None, 100, 10, 100) == 1  # tiny problems are not split
    # header + pivot record (SP + m) + 5 doubles per 256-row block + one taken byte per row + a candidate record
    assert lib.pls_cv_scratch_doubles(1000, 8, 16) >= 16 + (12 + 16) + 5 * 4 + 125 + lib.pls_cv_candidate_doubles(8, 16)
    assert lib.pls_cv_scratch_doubles(1000, 0, 16) == 0 and lib.pls_cv_scratch_doubles(1000, 8, 1) == 0


def test_no_gpu_fails_loudly():
    from projected_langevin_sampling_b200 import _native

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeLibraryError):
        _native.context()
    lib = _native.load_library()
    h = ctypes.c_void_p()
    assert lib.pls_ctx_create(0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.pls_last_error(None)
    import projected_langevin_sampling_b200 as pkg

    with pytest.raises(_native.NativeLibraryError):  # no silent CPU fallback anywhere on the path
        pkg.OrthonormalBasis(pkg.PLSKernel(pkg.LinearKernel(), torch.ones(2, 2)), torch.ones(2, 2), torch.ones(3, 2))


def test_kernel_spec_duck_typing():
    from projected_langevin_sampling_b200 import _native, kernels

    k = kernels.ScaleKernel(kernels.RBFKernel(ard_num_dims=3, lengthscale=torch.tensor([[1.0, 2.0, 4.0]])), outputscale=2.5)
    spec = kernels.kernel_spec(k, 3)
    assert spec.kernel_id == _native.KERNEL_RBF and spec.lengthscale == [1.0, 2.0, 4.0] and spec.outputscale == 2.5
    assert spec.inv_lengthscale == [1.0, 0.5, 0.25]
    assert kernels.kernel_spec(kernels.ScaleKernel(kernels.RBFKernel(lengthscale=0.15), 3.0), 2).lengthscale == [0.15, 0.15]

    class RBFKernel:  # a gpytorch-like object: only the attribute names matter
        lengthscale = torch.tensor([[0.7]])

    class ScaleKernel:
        base_kernel = RBFKernel()
        outputscale = torch.tensor(1.9)

    spec = kernels.kernel_spec(ScaleKernel(), 1)
    assert spec.kernel_id == _native.KERNEL_RBF and abs(spec.outputscale - 1.9) < 1e-7

    class MockKernel:
        pass

    assert kernels.kernel_spec(MockKernel(), 2).kernel_id == _native.KERNEL_LINEAR

    class Matern:
        pass

    with pytest.raises(TypeError):
        kernels.kernel_spec(Matern(), 2)
    with pytest.raises(ValueError):
        kernels.kernel_spec(k, 5)


def test_cost_native_struct():
    from projected_langevin_sampling_b200 import _native as nat
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions as lf

    y = torch.zeros(3)
    c = costs.GaussianCost(0.5, y, lf.IdentityLinkFunction()).native()
    assert (c.cost_id, c.link_id, c.closed_form, c.observation_noise) == (nat.COST_GAUSSIAN, nat.LINK_IDENTITY, 1, 0.5)
    assert costs.GaussianCost(0.5, y, lf.SquareLinkFunction()).native().closed_form == 0
    assert costs.GaussianCost(0.5, y, lf.IdentityLinkFunction()).native(force_autograd=True).closed_form == 0
    b = costs.BernoulliCost(y, lf.SigmoidLinkFunction(jitter=1e-8))
    assert b.y_train.dtype == torch.double and b.native().link_jitter == 1e-8 and b.native().closed_form == 1
    mm = costs.MultiModalCost(0.4, 1.5, 0.3, y, lf.IdentityLinkFunction()).native()
    assert (mm.shift, mm.bernoulli_noise, mm.closed_form) == (1.5, 0.3, 0)
    st = costs.StudentTCost(4.0, y, lf.IdentityLinkFunction(), scale=0.7).native()
    assert (st.degrees_of_freedom, st.scale, st.closed_form) == (4.0, 0.7, 1)
    # the probit constant follows the default dtype at call time, as link_functions.py:42 does
    torch.set_default_dtype(torch.float64)
    try:
        assert abs(costs.BernoulliCost(y, lf.ProbitLinkFunction()).native().probit_divisor - 2.0**0.5) < 5e-16
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(costs.BernoulliCost(y, lf.ProbitLinkFunction()).native().probit_divisor - 1.4142135381698608) < 1e-15


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "projected_langevin_sampling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("pls_oracle", "oracle") or "import oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_shard_ranges_partition():
    from projected_langevin_sampling_b200.distributed import GridPlacement, shard_range

    for total, world, align in [(4096, 8, 2), (1000, 3, 2), (1_000_000, 4, 128), (7, 8, 1), (20_000_000, 2, 128)]:
        pieces = [shard_range(total, r, world, align) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == total
        for (a, b), (c, d) in zip(pieces, pieces[1:]):
            assert b == c and a <= b
        assert all(b % align == 0 for _, b in pieces[:-1])
    g = GridPlacement(rank=5, world=8, n_groups=2, j_groups=4)
    assert (g.n_index, g.j_index) == (1, 1) and g.row_group_ranks() == [1, 5]
    assert g.rows(20_000_000) == (10_000_000, 20_000_000) and g.particles(16384) == (4096, 8192)
    with pytest.raises(ValueError):
        GridPlacement(rank=0, world=8, n_groups=3, j_groups=2)


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PLS_ROOT"])
from projected_langevin_sampling_b200.distributed import GridPlacement, gradient_allreduce, make_row_group, shard_range
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# (a) particle sharding: every rank owns a slice, slices tile [0, J), no communication needed for the update itself
j0, j1 = shard_range(1001, rank, world, align=2)
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([j1 - j0]))
assert sum(int(s) for s in sizes) == 1001
# (b) row sharding: 2 x 1 grid, the (M, J_local) gradient is summed over the row group
place = GridPlacement(rank=rank, world=world, n_groups=2, j_groups=1)
hook = gradient_allreduce(make_row_group(place))
g = torch.full((4, 6), float(rank + 1), dtype=torch.float64)
hook(g)
assert torch.equal(g, torch.full((4, 6), 3.0, dtype=torch.float64)), g
r0, r1 = place.rows(1000)
assert (r0, r1) == ((0, 512) if rank == 0 else (512, 1000))
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, PLS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_early_stopper_follows_reference_rules():
    """experiments/early_stopper.py:4-24: non-finite -> stop; no improvement accumulates simulated time; improvement resets."""
    from projected_langevin_sampling_b200.early_stopper import EarlyStopper

    s = EarlyStopper(patience=0.25)
    assert not s.should_stop(10.0, 0.1)          # first loss becomes the minimum
    assert not s.should_stop(10.0, 0.1)          # equal counts as "no improvement": t = 0.1
    assert not s.should_stop(11.0, 0.1)          # t = 0.2
    assert not s.should_stop(9.0, 0.1)           # improvement: t = 0
    assert not s.should_stop(9.5, 0.1) and not s.should_stop(9.5, 0.1)  # t = 0.2
    assert s.should_stop(9.5, 0.1)               # t = 0.3 >= 0.25
    assert EarlyStopper().should_stop(float("nan"), 1e-3) and EarlyStopper().should_stop(float("inf"), 1e-3)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exactly one JSON line on stdout with the
    contract's keys, whatever the libraries print."""
    import json

    # (a 20 000-row slice: the full-N arm needs ~30 GB of host memory and minutes of CPU time)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--rows", "20000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    # "reference+stub" = the unmodified reference installed under oracle/_ref (build container / GPU box), "port" = the oracle's
    # restatement when that install is absent
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference+stub", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["J_global"] == 4096 and d["cpu_baseline"]["measured"]["j_chunk"] == 256
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d


def test_gram_mode_argument_and_environment(monkeypatch):
    """engine.gram_mode: the default is the generated Gram (nothing N x M in memory, BASELINE.json north_star); the other modes
    are opt-in by argument or by PLS_B200_GRAM_CACHE, and anything else is an error."""
    from projected_langevin_sampling_b200.engine import gram_mode

    monkeypatch.delenv("PLS_B200_GRAM_CACHE", raising=False)
    assert gram_mode(None) is False and gram_mode(False) is False and gram_mode("generated") is False
    assert gram_mode(True) is True and gram_mode("cached") is True
    assert gram_mode("auto") == "auto" and gram_mode("staged") == "staged"
    with pytest.raises(ValueError):
        gram_mode("sometimes")
    monkeypatch.setenv("PLS_B200_GRAM_CACHE", "staged")
    assert gram_mode(False) == "staged"
    monkeypatch.setenv("PLS_B200_GRAM_CACHE", "off")
    assert gram_mode(True) is False


def test_psd_safe_cholesky_jitter_retries():
    """InducingPointBasis factorises k(Z, Z) as gpytorch.solve does: plain Cholesky, then diagonal jitter 1e-8, 1e-7, 1e-6."""
    import warnings

    from projected_langevin_sampling_b200.projected_langevin_sampling.basis.inducing_point import psd_safe_cholesky

    g = torch.Generator().manual_seed(0)
    a = torch.randn(6, 6, generator=g, dtype=torch.float64)
    spd = a @ a.T + torch.eye(6, dtype=torch.float64)
    assert torch.equal(psd_safe_cholesky(spd), torch.linalg.cholesky(spd))  # untouched when positive definite
    u = torch.randn(6, 2, generator=g, dtype=torch.float64)
    near = u @ u.T  # rank 2: numerically indefinite
    near[0, 0] -= 1e-9
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        l = psd_safe_cholesky(near)
    assert w and "jitter" in str(w[0].message)
    assert (l @ l.T - near).abs().max() < 1e-5
    with pytest.raises(torch.linalg.LinAlgError):
        psd_safe_cholesky(-torch.eye(3, dtype=torch.float64))
    bad = spd.clone()
    bad[2, 1] = float("nan")  # the factorisation reads the lower triangle
    with pytest.raises(ValueError):
        psd_safe_cholesky(bad)


def test_gaussian_predict_is_a_light_diagonal_normal():
    """GaussianCost.predict: mean / variance / stddev / confidence_region as the reference's gpytorch object offers, without the
    dense Cholesky (and without failing on a zero variance, J = 1 gives NaN variance in the reference too)."""
    from projected_langevin_sampling_b200.projected_langevin_sampling import costs, link_functions as lf

    s = torch.tensor([[1.0, 3.0, 5.0], [2.0, 2.0, 2.0]], dtype=torch.float64)
    d = costs.GaussianCost(0.1, torch.zeros(2), lf.IdentityLinkFunction()).predict(s)
    assert torch.equal(d.mean, torch.tensor([3.0, 2.0], dtype=torch.float64))
    assert torch.equal(d.variance, torch.tensor([4.0, 0.0], dtype=torch.float64))
    lo, hi = d.confidence_region()
    assert torch.equal(lo, torch.tensor([-1.0, 2.0], dtype=torch.float64)) and torch.equal(hi, torch.tensor([7.0, 2.0], dtype=torch.float64))
    assert torch.equal(d.covariance_matrix, torch.diag(d.variance)) and d.sample().shape == (2,)


def test_bench_strong_scaling_placement_and_shared_config():
    """bench.py at N GPUs: the named J = 4096 is split over a 1 x N grid (strong scaling, equal even slices that tile the
    particle axis), and both arms print the same `config` dictionary for the same workload and world size."""
    sys.path.insert(0, ROOT)
    import bench
    from projected_langevin_sampling_b200.distributed import GridPlacement

    w = bench.WORKLOADS["c4"]
    for world in (1, 2, 4, 8):
        slices = [GridPlacement(rank=r, world=world, n_groups=1, j_groups=world).particles(w["j"]) for r in range(world)]
        assert slices[0][0] == 0 and slices[-1][1] == w["j"]
        assert all(a[1] == b[0] for a, b in zip(slices[:-1], slices[1:]))
        assert {b - a for a, b in slices} == {w["j"] // world}
        rows = [GridPlacement(rank=r, world=world, n_groups=1, j_groups=world).rows(w["n"]) for r in range(world)]
        assert all(r == (0, w["n"]) for r in rows)  # data replicated: no row sharding in the default grid
    cfg = bench.base_config(w, 8, "1x8")
    assert cfg == bench.base_config(dict(w), 8, "1x8") and cfg["J_global"] == 4096 and cfg["N"] == 1_000_000 and cfg["M"] == 1024
    # a 2 x 4 grid: rows split in two 128-aligned halves, particles in four slices; ranks of a row group share their particle slice
    g = [GridPlacement(rank=r, world=8, n_groups=2, j_groups=4) for r in range(8)]
    assert g[0].rows(w["n"])[1] == g[4].rows(w["n"])[0] and g[0].rows(w["n"])[1] % 128 == 0
    assert g[1].particles(w["j"]) == g[5].particles(w["j"]) == (1024, 2048) and g[1].row_group_ranks() == [1, 5]
