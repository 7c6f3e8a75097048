"""GPU unit tests of individual C-ABI entry points (called through the ctypes binding): the hot loop's exponent routine,
the small DGEMM, the split reduction, direct forward/backward calls with leading dimensions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from projected_langevin_sampling_b200 import _native

    return _native.context()


def _ulps(got: torch.Tensor, want: torch.Tensor) -> torch.Tensor:
    spacing = torch.from_numpy(np.spacing(want.cpu().numpy())).to(want.device)
    return (got - want).abs() / spacing


def test_hot_loop_exp_within_two_ulps(ctx):
    from projected_langevin_sampling_b200 import ops

    g = torch.Generator().manual_seed(0)
    x = torch.cat([
        -700.0 * torch.rand(2_000_000, generator=g, dtype=torch.float64),  # the Gram exponent's range
        10.0 * torch.rand(200_000, generator=g, dtype=torch.float64) - 5.0,
        torch.tensor([0.0, -1e-300, -1e-17, -0.5, -1.0, -36.04365338911715, -699.999, 1.0, 2.302585092994046]),
    ]).cuda()
    fast = ops.gram_exp(ctx, x, fast=True)
    lib = ops.gram_exp(ctx, x, fast=False)
    want = torch.exp(x.cpu().to(torch.float64)).cuda()  # torch CPU exp as the third opinion
    assert _ulps(lib, want).max().item() <= 1.0
    assert _ulps(fast, want).max().item() <= 2.0
    assert _ulps(fast, want).mean().item() < 0.5
    assert ops.gram_exp(ctx, torch.tensor([0.0], dtype=torch.float64).cuda(), fast=True).item() == 1.0
    # clamped below -700: tiny but finite, never NaN
    tail = ops.gram_exp(ctx, torch.tensor([-1e4, -1e8, -float("inf")], dtype=torch.float64).cuda(), fast=True)
    assert torch.isfinite(tail).all() and (tail < 1e-300).all() and (tail >= 0).all()


@pytest.mark.parametrize("rows,k,j,trans", [(64, 64, 64, False), (100, 37, 130, False), (33, 200, 77, True), (1, 5, 3, True)])
def test_small_gemm_matches_torch(ctx, rows, k, j, trans):
    from projected_langevin_sampling_b200 import ops

    g = torch.Generator().manual_seed(rows + k + j)
    a = torch.randn((k, rows) if trans else (rows, k), generator=g, dtype=torch.float64).cuda()
    b = torch.randn(k, j, generator=g, dtype=torch.float64).cuda()
    out = torch.zeros(rows, j + 3, dtype=torch.float64).cuda()  # leading dimension > j
    ops.gemm(ctx, a, b, out[:, :j], trans_a=trans)
    want = (a.T if trans else a).cpu() @ b.cpu()
    assert (out[:, :j].cpu() - want).abs().max().item() <= 1e-12 * max(1.0, want.abs().max().item())
    assert (out[:, j:] == 0).all()  # nothing written outside the logical matrix


def test_forward_backward_direct_calls_with_leading_dimensions(ctx):
    """pls_forward_f64 / pls_backward_f64 / pls_reduce_splits_f64 called directly, against dense torch algebra on the
    Gram matrix the library itself produces with pls_gram_f64."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(11)
    n, m, d, j = 700, 90, 5, 150
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
    inv_ls = [0.5, 0.4, 0.3, 0.6, 0.7]
    centre = z.mean(0).tolist()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.0)
    za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, inv_ls, centre, float(np.log(1.7)))
    k_xz = ops.gram(ctx, nat.KERNEL_RBF, xa, za, d)  # (n, m), includes the outputscale
    dense = 1.7 * torch.exp(-0.5 * (((x[:, None, :] - z[None, :, :]) * torch.tensor(inv_ls, dtype=torch.float64).cuda()) ** 2).sum(-1))
    assert (k_xz - dense).abs().max().item() < 1e-13
    ldw = 160
    w_store = torch.randn(m, ldw, generator=g, dtype=torch.float64).cuda()
    f_store = torch.full((n, 156), -7.0, dtype=torch.float64).cuda()
    ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w_store, j, nat.EPI_PREDICTION, f_store)
    want_f = k_xz @ w_store[:, :j]
    assert (f_store[:, :j] - want_f).abs().max().item() < 1e-11
    assert (f_store[:, j:] == -7.0).all()
    dc_store = torch.randn(n, 152, generator=g, dtype=torch.float64).cuda()
    for splits in (1, 3, 7):
        gp = torch.full((splits, m, 154), 5.0, dtype=torch.float64).cuda()
        ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc_store, j, gp, splits, accumulate=False)
        out = torch.empty(m, 154, dtype=torch.float64).cuda()
        ops.reduce_splits(ctx, gp, j, out)
        want_g = k_xz.T @ dc_store[:, :j]
        assert (out[:, :j] - want_g).abs().max().item() < 1e-10 * want_g.abs().max().item()
        ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc_store, j, gp, splits, accumulate=True)  # += doubles it
        ops.reduce_splits(ctx, gp, j, out)
        assert (out[:, :j] - 2 * want_g).abs().max().item() < 1e-10 * want_g.abs().max().item()
        assert (gp[:, :, j:] == 5.0).all()


def test_argument_errors_are_reported_not_thrown(ctx):
    from projected_langevin_sampling_b200 import _native as nat

    w = torch.zeros(4, 5, dtype=torch.float64).cuda()  # odd leading dimension: rejected by the streamed-operand contract
    out = torch.zeros(8, 6, dtype=torch.float64).cuda()
    xa = torch.zeros(8, 4, dtype=torch.float64).cuda()
    za = torch.zeros(4, 4, dtype=torch.float64).cuda()
    rc = ctx.lib.pls_forward_f64(ctx.handle, nat.KERNEL_RBF, xa.data_ptr(), 8, za.data_ptr(), 4, 1, w.data_ptr(), 5, 5,
                                 nat.EPI_PREDICTION, None, None, out.data_ptr(), 6, ctx.stream())
    assert rc != 0 and b"even ldw" in ctx.lib.pls_last_error(ctx.handle)
    rc = ctx.lib.pls_forward_f64(ctx.handle, 9, xa.data_ptr(), 8, za.data_ptr(), 4, 1, w.data_ptr(), 6, 5,
                                 nat.EPI_PREDICTION, None, None, out.data_ptr(), 6, ctx.stream())
    assert rc != 0 and b"unknown kernel_id" in ctx.lib.pls_last_error(ctx.handle)
    with pytest.raises(nat.NativeLibraryError):
        ctx.check(rc)


@pytest.mark.parametrize("rt", [1, 2])
@pytest.mark.parametrize(
    "n,m,d,j,ldw,lddc",
    [
        (130, 70, 26, 1, 2, 2),          # the largest supported D, a single particle
        (64, 32, 4, 256, 256, 256),      # exactly one wide tile, one pipeline stage
        (1000, 257, 7, 513, 530, 514),   # ragged everything; leading dimensions that are not multiples of 16 (2-D TMA path)
        (333, 97, 1, 300, 304, 320),     # D = 1; leading dimensions that are multiples of 16 (3-D TMA path)
        (40, 5, 3, 20, 32, 20),          # fewer reduction points than one 8-point group
    ],
)
def test_generated_operand_gemm_tile_shapes(ctx, rt, n, m, d, j, ldw, lddc):
    """Both CTA tile shapes (64 x 256 and 128 x 128), both TMA descriptor paths, ragged edges: forward, backward with
    several split counts and the cost-sum epilogue against dense torch algebra on pls_gram_f64's matrix."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    ctx.lib.pls_set_tile_shape(ctx.handle, rt)
    try:
        g = torch.Generator().manual_seed(1000 * n + j)
        x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
        z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
        inv_ls = [1.0 / (1.0 + 0.1 * k + d ** 0.5) for k in range(d)]
        centre = z.mean(0).tolist()
        xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.0)
        za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, inv_ls, centre, float(np.log(0.8)))
        k_xz = ops.gram(ctx, nat.KERNEL_RBF, xa, za, d)
        w_store = torch.randn(m, ldw, generator=g, dtype=torch.float64).cuda()
        f_store = torch.full((n, ldw), 3.0, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w_store, j, nat.EPI_PREDICTION, f_store)
        want_f = k_xz @ w_store[:, :j]
        scale = max(1.0, want_f.abs().max().item())
        assert (f_store[:, :j] - want_f).abs().max().item() < 1e-12 * scale
        assert (f_store[:, j:] == 3.0).all()
        # cost-sum epilogue: per-row-tile partial sums of the Gaussian cost
        y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
        cost = nat.PlsCost()
        cost.cost_id, cost.link_id, cost.closed_form, cost.observation_noise = nat.COST_GAUSSIAN, nat.LINK_IDENTITY, 1, 0.3
        tr = ops.forward_tile_rows(ctx, j)
        assert tr == (64 if rt == 1 else 128)
        tiles = (n + tr - 1) // tr
        part = torch.zeros(tiles, ldw, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w_store, j, nat.EPI_COST, part, cost=cost, y=y)
        want_c = ((want_f - y[:, None]) ** 2 / (2 * 0.3)).sum(0)
        assert (part[:, :j].sum(0) - want_c).abs().max().item() < 1e-11 * max(1.0, want_c.abs().max().item())
        # fused epilogue (pls_forward_step_f64): the cost derivative and the cost sums from one pass over the same F tile
        dc_out = torch.full((n, ldw), 9.0, dtype=torch.float64).cuda()
        part2 = torch.zeros(tiles, ldw, dtype=torch.float64).cuda()
        ops.forward_step(ctx, nat.KERNEL_RBF, xa, za, d, w_store, j, cost, y, dc_out, part2)
        want_dc = (want_f - y[:, None]) / 0.3
        assert (dc_out[:, :j] - want_dc).abs().max().item() < 1e-11 * max(1.0, want_dc.abs().max().item())
        assert (dc_out[:, j:] == 9.0).all()
        assert (part2[:, :j].sum(0) - want_c).abs().max().item() < 1e-11 * max(1.0, want_c.abs().max().item())
        dc_store = torch.randn(n, lddc, generator=g, dtype=torch.float64).cuda()
        want_g = k_xz.T @ dc_store[:, :j]
        for splits in (1, 2, 5):
            gp = torch.full((splits, m, lddc), -2.0, dtype=torch.float64).cuda()
            ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc_store, j, gp, splits, accumulate=False)
            out = torch.empty(m, lddc, dtype=torch.float64).cuda()
            ops.reduce_splits(ctx, gp, j, out)
            assert (out[:, :j] - want_g).abs().max().item() < 1e-11 * max(1.0, want_g.abs().max().item())
            assert (gp[:, :, j:] == -2.0).all()
    finally:
        ctx.lib.pls_set_tile_shape(ctx.handle, 0)


def test_non_finite_cost_derivative_stays_in_its_column(ctx):
    """An Inf in Dc (Poisson at F = 0 produces one, costs/poisson.py:76-82 has no guard) must only affect its own particle
    column, as in the reference's dense product -- also when it sits in a row past the end of another split."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(5)
    n, m, d, j = 500, 40, 2, 64
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = x[:m].clone()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, [1.0, 1.0], [0.0, 0.0], 0.0)
    za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, [1.0, 1.0], [0.0, 0.0], 0.0)
    dc = torch.randn(n, j, generator=g, dtype=torch.float64).cuda()
    dc[321, 7] = float("inf")
    splits = 3
    gp = torch.zeros(splits, m, j, dtype=torch.float64).cuda()
    ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc, j, gp, splits, accumulate=False)
    out = torch.empty(m, j, dtype=torch.float64).cuda()
    ops.reduce_splits(ctx, gp, j, out)
    finite_cols = torch.isfinite(out).all(0)
    assert not finite_cols[7] and finite_cols[torch.arange(j).cuda() != 7].all()


@pytest.mark.parametrize("bounds", [[0, 1700, 3000], [0, 1, 1, 2999, 3000], [0, 3000]])
@pytest.mark.parametrize("threshold", [0.0, 2400.0])
@pytest.mark.parametrize("rule", ["numpy", "stable"])
def test_row_sharded_selector_equals_unsharded(ctx, bounds, threshold, rule):
    """pls_cv_shard_* with the ranks emulated on one GPU (the all-gather is a concatenation): uneven shards, a one-row
    shard and an EMPTY shard; the indices and the early-stop count must equal pls_cv_select_f64's, under both tie rules
    (duplicated points straddling the shards force ties: "numpy" gathers d and asks the host, "stable" decides by GLOBAL index)."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(17)
    n, d, m = 3000, 5, 40
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    x[:1000] = x[1700:2700]  # exact duplicates straddling the shards: whenever one of them is the maximum it is a 2-way tie
    x[2999] = x[1]
    inv_ls = [0.5, 0.6, 0.7, 0.4, 0.45]
    centre = x.mean(0).tolist()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.5 * float(np.log(1.3)))
    info = {}
    want_idx, want_n = ops.cv_select(ctx, nat.KERNEL_RBF, xa, d, 1.3, m, 1e-12, threshold, tie_rule=rule, info=info)
    states = [ops.ShardedSelectorState(ctx, nat.KERNEL_RBF, xa[a:b].contiguous(), a, n, d, 1.3, m, 1e-12, threshold, tie_rule=rule)
              for a, b in zip(bounds[:-1], bounds[1:])]
    gathered = []

    def gather_d(slices):
        gathered.append(1)
        return torch.cat(slices).cpu().numpy()

    got_idx, got_n = ops.cv_select_sharded(states, lambda recs: torch.cat(recs), gather_d)
    assert got_n == want_n and (threshold == 0.0) == (want_n == m)
    assert torch.equal(got_idx, want_idx)
    assert len(gathered) == info["host_tie_calls"] and (rule == "stable") == (info["host_tie_calls"] == 0)
    assert info["host_tie_calls"] + info["tied_picks"] >= 3  # the construction did produce ties
    for s in states[1:]:  # every rank ends with the same answer
        assert torch.equal(s.indices, got_idx)


def test_selector_numpy_rule_against_cpu_restatement(ctx):
    """Many-way exact ties (points on a 1-D grid with a short lengthscale, as the README demo) at a size where several batches
    and several host decisions interleave: the indices equal a literal numpy restatement of conditional_variance.py:91-109 fed
    with the kernel's own d (so only the tie handling is under test here, not round-off)."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    n, m = 5000, 48
    perm = np.random.default_rng(5).permutation(n)
    x = torch.linspace(-4, 4, n, dtype=torch.float64)[perm][:, None].cuda()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, [1 / 0.07], [0.0], 0.0)
    seen = []

    def spy(dh, chosen):  # the reference's rule, recording what it was asked
        seen.append((dh.copy(), chosen.copy()))
        return ops.numpy_tie_rule(dh, chosen)

    info = {}
    idx, nsel = ops.cv_select(ctx, nat.KERNEL_RBF, xa, 1, 1.0, m, 1e-12, 0.0, tie_rule=spy, info=info)
    idx = idx.cpu().numpy()
    assert nsel == m and len(set(idx.tolist())) == m and len(seen) >= 5 and info["host_tie_calls"] == len(seen)
    for dh, chosen in seen:  # each call was a genuine tie, and the pivot stored for that slot is the rule's answer
        rest = np.delete(dh, chosen)
        assert (rest == rest.max()).sum() > 1
        assert idx[len(chosen)] == ops.numpy_tie_rule(dh, chosen)
        assert idx[: len(chosen)].tolist() == chosen.tolist()


def test_branch_free_epilogue_arithmetic_in_ulps(ctx):
    """FlatMath (csrc/pls_cost.cuh): the division, log and exp the register epilogue evaluates without branches, against
    float64 CPU results: <= 1 ulp (div), <= 2 ulp (log, exp); IEEE results for zero / infinite / NaN operands."""
    from projected_langevin_sampling_b200 import ops

    g = torch.Generator().manual_seed(2)
    n = 1_000_000
    a = (torch.randn(n, generator=g, dtype=torch.float64) * torch.exp(20 * torch.randn(n, generator=g, dtype=torch.float64)))
    b = (torch.randn(n, generator=g, dtype=torch.float64) * torch.exp(20 * torch.randn(n, generator=g, dtype=torch.float64)))
    q = ops.flat_math(ctx, 0, a.cuda(), b.cuda())
    assert _ulps(q, (a / b).cuda()).max().item() <= 1.0
    x = torch.cat([torch.exp(40 * torch.randn(n, generator=g, dtype=torch.float64)), 1.0 + 1e-3 * torch.randn(n, generator=g, dtype=torch.float64),
                   torch.tensor([1.0, 0.5, 2.0, 2.0 ** 0.5, 0.7071067811865476, 5e-324, 1e-310, 1.7976931348623157e308])])
    x = x[torch.isfinite(x) & (x > 0)]
    lg = ops.flat_math(ctx, 1, x.cuda())
    want = torch.log(x).cuda()
    nz = want != 0
    assert _ulps(lg[nz], want[nz]).max().item() <= 2.0
    assert (lg[~nz] == 0).all()  # log 1 = 0 exactly
    ex = 700 * (2 * torch.rand(n, generator=g, dtype=torch.float64) - 1)
    assert _ulps(ops.flat_math(ctx, 2, ex.cuda()), torch.exp(ex).cuda()).max().item() <= 2.0
    inf, nan = float("inf"), float("nan")
    sa = torch.tensor([1.0, -1.0, 0.0, 1.0, -2.0, inf, nan, 3.0], dtype=torch.float64)
    sb = torch.tensor([0.0, 0.0, 0.0, inf, -inf, inf, 1.0, nan], dtype=torch.float64)
    got, want = ops.flat_math(ctx, 0, sa.cuda(), sb.cuda()).cpu(), sa / sb
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.equal(got[~torch.isnan(want)], want[~torch.isnan(want)])
    sl = torch.tensor([0.0, -1.0, inf, nan, 1.0], dtype=torch.float64)
    got, want = ops.flat_math(ctx, 1, sl.cuda()).cpu(), torch.log(sl)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.equal(got[~torch.isnan(want)], want[~torch.isnan(want)])
    se = torch.tensor([710.0, 1e4, inf, -1e4, nan, 0.0], dtype=torch.float64)
    got = ops.flat_math(ctx, 2, se.cuda()).cpu()
    assert got[0] == inf and got[1] == inf and got[2] == inf and 0 <= got[3] < 1e-300 and torch.isnan(got[4]) and got[5] == 1.0


def _random_shapes(count, seed):
    rng = np.random.default_rng(seed)
    shapes = []
    for _ in range(count):
        n = int(rng.integers(1, 1400))
        m = int(rng.integers(2, 420))
        d = int(rng.integers(1, 27))
        j = int(rng.integers(1, 700))
        pad = int(rng.choice([0, 2, 6, 14]))
        shapes.append((n, m, d, j, j + (j & 1) + pad))
    return shapes


@pytest.mark.parametrize("n,m,d,j,ld", _random_shapes(14, seed=2024))
def test_generated_operand_gemm_random_shapes(ctx, n, m, d, j, ld):
    """Seeded random (N, M, D, J, leading dimension) draws -- ragged tiles, every exponent depth, one- and many-stage
    reductions, persistent CTAs with zero, one or several tiles: all four forward epilogues and the backward role, with
    generated and with cached Gram values, against dense float64 torch algebra on pls_gram_f64's matrix."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(n * 7919 + m * 31 + j)
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
    inv_ls = [1.0 / (2.0 + d ** 0.5 + 0.05 * k) for k in range(d)]
    centre = z.mean(0).tolist()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.0)
    za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, inv_ls, centre, float(np.log(1.4)))
    k_xz = ops.gram(ctx, nat.KERNEL_RBF, xa, za, d)
    w = torch.randn(m, ld, generator=g, dtype=torch.float64).cuda()
    y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
    want_f = k_xz @ w[:, :j]
    scale = max(1.0, want_f.abs().max().item())
    cost = nat.PlsCost()
    cost.cost_id, cost.link_id, cost.closed_form = nat.COST_STUDENT_T, nat.LINK_IDENTITY, 1
    cost.degrees_of_freedom, cost.scale, cost.link_jitter, cost.probit_divisor = 4.0, 0.7, 1e-10, 2.0 ** 0.5
    e = want_f - y[:, None]
    want_dc = 5.0 * e / (4.0 * 0.49 + e * e)
    want_c = (2.5 * torch.log(1.0 + e * e / (4.0 * 0.49))).sum(0)
    # generated Gram values, then the same launches with the Gram cached in HBM (pls_*_cached_f64) -- offset into a larger cache
    # so that the row padding the kernels may read holds OTHER rows' (finite, non-zero) values, as in the engine's row chunks
    wrap = torch.arange(256, device=xa.device) % n
    big = ops.gram_cache(ctx, nat.KERNEL_RBF, torch.cat([xa[wrap], xa, xa[wrap]]), za, d)
    assert big.shape[1] == (m + 127) // 128 * 128 and big.shape[0] % 128 == 0 and torch.equal(big[256:256 + n, :m], k_xz)
    big[:, m:] = 0.37  # column padding: read, multiplied by zero rows of the streamed matrix, must only be finite
    for gram in (None, big[256:]):
        f = torch.full((n, ld), 3.0, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_PREDICTION, f, gram=gram)
        assert (f[:, :j] - want_f).abs().max().item() < 1e-12 * scale and (f[:, j:] == 3.0).all()
        dc = torch.full((n, ld), 3.0, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_COST_DERIVATIVE, dc, cost=cost, y=y, gram=gram)
        assert (dc[:, :j] - want_dc).abs().max().item() < 1e-11 * max(1.0, want_dc.abs().max().item()) and (dc[:, j:] == 3.0).all()
        tiles = (n + ops.forward_tile_rows(ctx, j) - 1) // ops.forward_tile_rows(ctx, j)
        part = torch.zeros(tiles, ld, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_COST, part, cost=cost, y=y, gram=gram)
        assert (part[:, :j].sum(0) - want_c).abs().max().item() < 1e-11 * max(1.0, want_c.abs().max().item())
        dc2 = torch.full((n, ld), 4.0, dtype=torch.float64).cuda()
        part2 = torch.zeros(tiles, ld, dtype=torch.float64).cuda()
        ops.forward_step(ctx, nat.KERNEL_RBF, xa, za, d, w, j, cost, y, dc2, part2, gram=gram)
        assert (dc2[:, :j] - want_dc).abs().max().item() < 1e-11 * max(1.0, want_dc.abs().max().item()) and (dc2[:, j:] == 4.0).all()
        assert (part2[:, :j].sum(0) - want_c).abs().max().item() < 1e-11 * max(1.0, want_c.abs().max().item())
        dcin = torch.randn(n, ld, generator=g, dtype=torch.float64).cuda()
        want_g = k_xz.T @ dcin[:, :j]
        splits = ops.backward_splits(ctx, n, m, j)
        gp = torch.full((splits, m, ld), -2.0, dtype=torch.float64).cuda()
        ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dcin, j, gp, splits, accumulate=False, gram=gram)
        out = torch.empty(m, ld, dtype=torch.float64).cuda()
        ops.reduce_splits(ctx, gp, j, out)
        assert (out[:, :j] - want_g).abs().max().item() < 1e-11 * max(1.0, want_g.abs().max().item()) and (gp[:, :, j:] == -2.0).all()


@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (37, 129, 3), (1000, 300, 8), (4100, 64, 16), (513, 1000, 26)])
def test_gram_fill_matches_gram(ctx, n, m, d):
    """pls_gram_fill_f64 (the stream-speed fill of the cached / staged Gram buffers, table exp) against pls_gram_f64 (library
    exp): a few ulps apart, nothing written outside rows [0, n) x columns [0, m)."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(n + m + d)
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
    inv_ls = [1.0 / (1.0 + d ** 0.5 + 0.1 * k) for k in range(d)]
    centre = z.mean(0).tolist()
    for kid in (nat.KERNEL_RBF, nat.KERNEL_LINEAR):
        xa = ops.prepare_points(ctx, kid, x, inv_ls, centre if kid == nat.KERNEL_RBF else [0.0] * d, 0.0)
        za = ops.prepare_points(ctx, kid, z, inv_ls, centre if kid == nat.KERNEL_RBF else [0.0] * d, float(np.log(1.7)))
        want = ops.gram(ctx, kid, xa, za, d)
        ld = int(ctx.lib.pls_gram_cache_ld(m))
        buf = torch.full((int(ctx.lib.pls_gram_cache_rows(n)) + 5, ld), -7.0, dtype=torch.float64).cuda()
        ops.gram_fill(ctx, kid, xa, za, d, buf)
        got = buf[:n, :m]
        tol = 8 * np.finfo(np.float64).eps
        assert ((got - want).abs() <= tol * want.abs().clamp_min(1e-300) + (1e-15 if kid == nat.KERNEL_LINEAR else 0.0)).all()
        assert (buf[n:] == -7.0).all() and (buf[:, m:] == -7.0).all()


@pytest.mark.parametrize(
    "n,m,d,j,ld,cost_name",
    [
        (1000, 200, 8, 512, 512, "gaussian"),     # one 64 x 512 tile per row tile, several chunks
        (777, 130, 3, 900, 912, "poisson"),       # 4 column blocks of 256, the last one partial (388 valid columns); ragged rows / points
        (300, 33, 1, 1024, 1024, "bernoulli"),    # two 512-column tiles, 2 chunks (odd / even chunk counts park different halves)
        (200, 32, 16, 512, 520, "gaussian"),      # exactly one chunk: the parked half is never reloaded inside the main loop; 2-D TMA path
        (5000, 96, 5, 2048, 2048, "student_t"),   # persistent forward: several tiles per CTA
    ],
)
def test_parked_accumulators_are_bit_identical(ctx, n, m, d, j, ld, cost_name):
    """NS = 2 (64 x 512 CTA tiles, the second accumulator set parked in tensor memory and swapped once per chunk) accumulates every
    column over the chunks in the same order as NS = 1, so every entry point must return the SAME BITS either way: forward
    (prediction, cost derivative, cost sums, fused derivative + sums) and backward with several split counts."""
    from projected_langevin_sampling_b200 import _native as nat, ops

    g = torch.Generator().manual_seed(n + m + j)
    x = torch.randn(n, d, generator=g, dtype=torch.float64).cuda()
    z = torch.randn(m, d, generator=g, dtype=torch.float64).cuda()
    inv_ls = [1.0 / (1.0 + 0.1 * k + d ** 0.5) for k in range(d)]
    centre = z.mean(0).tolist()
    xa = ops.prepare_points(ctx, nat.KERNEL_RBF, x, inv_ls, centre, 0.0)
    za = ops.prepare_points(ctx, nat.KERNEL_RBF, z, inv_ls, centre, float(np.log(1.1)))
    w = (0.3 * torch.randn(m, ld, generator=g, dtype=torch.float64)).cuda()
    dc = torch.randn(n, ld, generator=g, dtype=torch.float64).cuda()
    cost = nat.PlsCost()
    cost.closed_form, cost.observation_noise, cost.link_jitter = 1, 0.3, 1e-10
    cost.degrees_of_freedom, cost.scale = 4.0, 1.0
    if cost_name == "gaussian":
        cost.cost_id, cost.link_id = nat.COST_GAUSSIAN, nat.LINK_IDENTITY
        y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
    elif cost_name == "poisson":
        cost.cost_id, cost.link_id = nat.COST_POISSON, nat.LINK_SQUARE
        y = torch.poisson(2.0 * torch.ones(n), generator=g).double().cuda()
    elif cost_name == "bernoulli":
        cost.cost_id, cost.link_id = nat.COST_BERNOULLI, nat.LINK_SIGMOID
        y = torch.bernoulli(0.5 * torch.ones(n), generator=g).double().cuda()
    else:
        cost.cost_id, cost.link_id = nat.COST_STUDENT_T, nat.LINK_IDENTITY
        y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
    tiles = (n + 63) // 64

    def run_all():
        out = {}
        for name, epi in (("f", nat.EPI_PREDICTION), ("dc", nat.EPI_COST_DERIVATIVE)):
            buf = torch.full((n, ld), 7.0, dtype=torch.float64).cuda()
            ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, epi, buf, cost=cost, y=y)
            out[name] = buf
        part = torch.full((tiles, ld), 7.0, dtype=torch.float64).cuda()
        ops.forward(ctx, nat.KERNEL_RBF, xa, za, d, w, j, nat.EPI_COST, part, cost=cost, y=y)
        out["sums"] = part
        dc2 = torch.full((n, ld), 7.0, dtype=torch.float64).cuda()
        part2 = torch.full((tiles, ld), 7.0, dtype=torch.float64).cuda()
        ops.forward_step(ctx, nat.KERNEL_RBF, xa, za, d, w, j, cost, y, dc2, part2)
        out["dc2"], out["sums2"] = dc2, part2
        for splits in (1, 3):
            gp = torch.full((splits, m, ld), 7.0, dtype=torch.float64).cuda()
            ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc, j, gp, splits, accumulate=False)
            ops.backward(ctx, nat.KERNEL_RBF, za, xa, d, dc, j, gp, splits, accumulate=True)
            out[f"gp{splits}"] = gp
        torch.cuda.synchronize()
        return out

    try:
        ctx.lib.pls_set_tile_sets(ctx.handle, 1)  # never
        want = run_all()
        ctx.lib.pls_set_tile_sets(ctx.handle, 2)  # whenever the shape allows (the default rule would skip these small launches)
        ctx.lib.pls_set_tile_cluster(ctx.handle, 1)
        got = run_all()
        # ... and with pairs of CTAs sharing the generated Gram values through distributed shared memory (1024-column cluster tiles:
        # J = 900, 1024 and 2048 here; the cost-sum epilogues and the other widths keep the single-CTA kernel)
        ctx.lib.pls_set_tile_cluster(ctx.handle, 2)
        paired = run_all()
    finally:
        ctx.lib.pls_set_tile_sets(ctx.handle, 0)
        ctx.lib.pls_set_tile_cluster(ctx.handle, 0)
    for name in got:
        same = (paired[name] == got[name]) | (torch.isnan(paired[name]) & torch.isnan(got[name]))
        assert bool(same.all()), f"{name}: {int((~same).sum())} entries differ between CTA pairs (CL = 2) and single CTAs"
    k_xz = ops.gram(ctx, nat.KERNEL_RBF, xa, za, d)
    ref_f = k_xz @ w[:, :j]
    assert (want["f"][:, :j] - ref_f).abs().max().item() < 1e-12 * max(1.0, ref_f.abs().max().item())
    # with at most 3 chunks per tile the NS = 1 kernel runs the cost-sum epilogues (COST and the fused derivative + cost) through its
    # shared-memory staged path -- another row order for the sums, IEEE division / exp instead of the branch-free routines for the
    # derivative -- while NS = 2 streams twice as many blocks and qualifies for the register epilogue: round-off agreement only
    staged_sums = (m + 31) // 32 <= 3
    for name in want:
        a, b = got[name], want[name]
        if staged_sums and name in ("sums", "sums2", "dc2"):
            assert (a[:, :j] - b[:, :j]).abs().max().item() <= 1e-13 * max(1.0, b[:, :j].abs().max().item())
            continue
        same = (a == b) | (torch.isnan(a) & torch.isnan(b))
        assert bool(same.all()), f"{name}: {int((~same).sum())} entries differ between NS = 2 and NS = 1"
        assert (a[..., j:] == 7.0).all(), f"{name}: wrote outside the logical matrix"
