"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/pls_b200.h).

Every function takes float64 CUDA tensors, passes raw pointers + leading dimensions, enqueues on torch's current
stream and returns tensors that it allocated with torch (device memory management is torch's job, the arithmetic
is the library's).  `ctx.launches` counts the kernels enqueued (bench.py reports it as `gpu_launches`).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _native as nat

F64 = torch.float64


def even(n: int) -> int:
    """Leading dimensions of streamed matrices must be even (16-byte rows for the bulk copies)."""
    return n + (n & 1)


def as_device_f64(t: torch.Tensor, device: Optional[torch.device] = None) -> torch.Tensor:
    dev = device or torch.device("cuda", torch.cuda.current_device())
    return t.detach().to(device=dev, dtype=F64).contiguous()


def alloc_matrix(rows: int, cols: int, device) -> Tuple[torch.Tensor, int]:
    """rows x even(cols) storage; returns (storage, ld).  Use storage[:, :cols] as the logical matrix."""
    ld = even(cols)
    return torch.empty((rows, ld), dtype=F64, device=device), ld


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "matrix must be row-major with unit column stride"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _dbl_array(values: Sequence[float]):
    arr = (C.c_double * len(values))(*[float(v) for v in values])
    return arr


# ---- setup ------------------------------------------------------------------------------------------------------
def prepare_points(ctx: nat.Context, kernel_id: int, x: torch.Tensor, inv_lengthscale: Sequence[float],
                   centre: Sequence[float], c_extra: float) -> torch.Tensor:
    """x (n x D) -> augmented layout (n x SP), see include/pls_b200.h."""
    nat.require_cuda_tensor(x, "x")
    n, d = x.shape
    sp = nat.point_stride(d)
    out = torch.empty((n, sp), dtype=F64, device=x.device)
    ctx.check(ctx.lib.pls_prepare_points_f64(ctx.handle, kernel_id, x.data_ptr(), n, d, _ld(x), _dbl_array(inv_lengthscale),
                                             _dbl_array(centre), float(c_extra), out.data_ptr(), ctx.stream()))
    ctx.launches += 1
    return out


def gram(ctx: nat.Context, kernel_id: int, rows_aug: torch.Tensor, cols_aug: torch.Tensor, d: int) -> torch.Tensor:
    out = torch.empty((rows_aug.shape[0], cols_aug.shape[0]), dtype=F64, device=rows_aug.device)
    if out.numel():
        ctx.check(ctx.lib.pls_gram_f64(ctx.handle, kernel_id, rows_aug.data_ptr(), rows_aug.shape[0], cols_aug.data_ptr(),
                                       cols_aug.shape[0], d, out.data_ptr(), max(out.shape[1], 1), ctx.stream()))
        ctx.launches += 1
    return out


# ---- step pieces ---------------------------------------------------------------------------------------------------
def gemm(ctx: nat.Context, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, trans_a: bool = False) -> torch.Tensor:
    """out[:, :j] = op(a) @ b[:, :j];  a, b, out row-major 2-D (views with a row stride are fine)."""
    rows = a.shape[1] if trans_a else a.shape[0]
    k = a.shape[0] if trans_a else a.shape[1]
    j = b.shape[1]
    assert b.shape[0] == k and out.shape[0] == rows and out.shape[1] >= j
    ctx.check(ctx.lib.pls_gemm_f64(ctx.handle, int(trans_a), a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), out.data_ptr(),
                                   _ld(out), rows, j, k, ctx.stream()))
    ctx.launches += 1
    return out


def gram_cache(ctx: nat.Context, kernel_id: int, rows_aug: torch.Tensor, cols_aug: torch.Tensor, d: int, chunk: int = 262144) -> torch.Tensor:
    """k(X, Z) kept resident for the pls_*_cached_f64 entry points: (pls_gram_cache_rows(N), pls_gram_cache_ld(M)) float64,
    zero padding, filled by pls_gram_f64 in row chunks.  Pass `cache[r0:]` as `gram=` for the rows starting at r0."""
    n, m = rows_aug.shape[0], cols_aug.shape[0]
    k = torch.zeros((int(ctx.lib.pls_gram_cache_rows(n)), int(ctx.lib.pls_gram_cache_ld(m))), dtype=F64, device=rows_aug.device)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        ctx.check(ctx.lib.pls_gram_f64(ctx.handle, kernel_id, rows_aug[r0:r1].data_ptr(), r1 - r0, cols_aug.data_ptr(), m, d,
                                       k[r0:].data_ptr(), k.shape[1], ctx.stream()))
        ctx.launches += 1
    return k


def gram_fill(ctx: nat.Context, kernel_id: int, rows_aug: torch.Tensor, cols_aug: torch.Tensor, d: int, out: torch.Tensor) -> torch.Tensor:
    """Rows [0, len(rows_aug)) of `out` (a gram_cache-shaped buffer) = k(rows, cols) at stream speed (pls_gram_fill_f64): the
    per-chunk staging buffer of the engine's "staged" Gram mode."""
    if rows_aug.shape[0] and cols_aug.shape[0]:
        ctx.check(ctx.lib.pls_gram_fill_f64(ctx.handle, kernel_id, rows_aug.data_ptr(), rows_aug.shape[0], cols_aug.data_ptr(),
                                            cols_aug.shape[0], d, out.data_ptr(), _ld(out), ctx.stream()))
        ctx.launches += 1
    return out


def forward(ctx: nat.Context, kernel_id: int, xa: torch.Tensor, za: torch.Tensor, d: int, w: torch.Tensor, j: int,
            epilogue: int, out: torch.Tensor, cost: Optional[nat.PlsCost] = None, y: Optional[torch.Tensor] = None,
            gram: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gram: rows [r0, ...) of a gram_cache() for the rows of xa -- the Gram values are then loaded, not generated."""
    costp = C.byref(cost) if cost is not None else None
    if gram is not None:
        ctx.check(ctx.lib.pls_forward_cached_f64(ctx.handle, gram.data_ptr(), _ld(gram), xa.shape[0], za.shape[0], w.data_ptr(), _ld(w),
                                                 j, epilogue, costp, nat.ptr(y), out.data_ptr(), _ld(out), ctx.stream()))
    else:
        ctx.check(ctx.lib.pls_forward_f64(ctx.handle, kernel_id, xa.data_ptr(), xa.shape[0], za.data_ptr(), za.shape[0], d,
                                          w.data_ptr(), _ld(w), j, epilogue, costp, nat.ptr(y), out.data_ptr(), _ld(out), ctx.stream()))
    ctx.launches += 1
    return out


def forward_step(ctx: nat.Context, kernel_id: int, xa: torch.Tensor, za: torch.Tensor, d: int, w: torch.Tensor, j: int,
                 cost: nat.PlsCost, y: torch.Tensor, dc: torch.Tensor, cost_partial: torch.Tensor,
                 gram: Optional[torch.Tensor] = None) -> None:
    """Forward with the cost derivative AND the per-row-tile cost sums from the same F tile (pls_forward_step_f64)."""
    if gram is not None:
        ctx.check(ctx.lib.pls_forward_step_cached_f64(ctx.handle, gram.data_ptr(), _ld(gram), xa.shape[0], za.shape[0], w.data_ptr(),
                                                      _ld(w), j, C.byref(cost), y.data_ptr(), dc.data_ptr(), _ld(dc),
                                                      cost_partial.data_ptr(), _ld(cost_partial), ctx.stream()))
    else:
        ctx.check(ctx.lib.pls_forward_step_f64(ctx.handle, kernel_id, xa.data_ptr(), xa.shape[0], za.data_ptr(), za.shape[0], d,
                                               w.data_ptr(), _ld(w), j, C.byref(cost), y.data_ptr(), dc.data_ptr(), _ld(dc),
                                               cost_partial.data_ptr(), _ld(cost_partial), ctx.stream()))
    ctx.launches += 1


def backward(ctx: nat.Context, kernel_id: int, za: torch.Tensor, xa: torch.Tensor, d: int, dc: torch.Tensor, j: int,
             gp: torch.Tensor, splits: int, accumulate: bool, gram: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gp: (splits, M, ld) storage.  gram: as in forward()."""
    assert gp.dim() == 3 and gp.shape[0] == splits and gp.shape[1] == za.shape[0] and gp.is_contiguous()
    if gram is not None:
        ctx.check(ctx.lib.pls_backward_cached_f64(ctx.handle, gram.data_ptr(), _ld(gram), za.shape[0], xa.shape[0], dc.data_ptr(),
                                                  _ld(dc), j, gp.data_ptr(), gp.shape[2], splits, int(accumulate), ctx.stream()))
    else:
        ctx.check(ctx.lib.pls_backward_f64(ctx.handle, kernel_id, za.data_ptr(), za.shape[0], xa.data_ptr(), xa.shape[0], d,
                                           dc.data_ptr(), _ld(dc), j, gp.data_ptr(), gp.shape[2], splits, int(accumulate),
                                           ctx.stream()))
    ctx.launches += 1
    return gp


def forward_tile_rows(ctx: nat.Context, j: int) -> int:
    """Rows per forward tile for a launch over j particle columns (granularity of the PLS_EPI_COST partial sums)."""
    return int(ctx.lib.pls_forward_tile_rows(ctx.handle, j))


def backward_splits(ctx: nat.Context, n_rows: int, m: int, j: int) -> int:
    return int(ctx.lib.pls_backward_splits(ctx.handle, n_rows, m, j))


def reduce_splits(ctx: nat.Context, gp: torch.Tensor, j: int, out: torch.Tensor) -> torch.Tensor:
    ctx.check(ctx.lib.pls_reduce_splits_f64(ctx.handle, gp.data_ptr(), gp.shape[0], gp.shape[1], j, gp.shape[2],
                                            out.data_ptr(), _ld(out), ctx.stream()))
    ctx.launches += 1
    return out


def project_update(ctx: nat.Context, vt: torch.Tensor, gm: torch.Tensor, p: torch.Tensor, j: int, inv_lambda: torch.Tensor,
                   eta: float, out: torch.Tensor, noise_mode: int = nat.NOISE_NONE, xi: Optional[torch.Tensor] = None,
                   seed: int = 0, step: int = 0, j_global_offset: int = 0, in_place: bool = False) -> torch.Tensor:
    m, m_k = vt.shape
    ctx.check(ctx.lib.pls_project_update_f64(ctx.handle, vt.data_ptr(), _ld(vt), m, m_k, gm.data_ptr(), _ld(gm), p.data_ptr(),
                                             _ld(p), j, inv_lambda.data_ptr(), float(eta), noise_mode, nat.ptr(xi),
                                             _ld(xi) if xi is not None else 0, seed & (2**64 - 1), step & (2**64 - 1),
                                             j_global_offset, int(in_place), out.data_ptr(), _ld(out), ctx.stream()))
    ctx.launches += 1
    return out


def cost_derivative(ctx: nat.Context, cost: nat.PlsCost, y: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    n, j = f.shape
    out = torch.empty((n, j), dtype=F64, device=f.device)
    if out.numel():
        ctx.check(ctx.lib.pls_cost_derivative_f64(ctx.handle, C.byref(cost), y.data_ptr(), f.data_ptr(), _ld(f), n, j,
                                                  out.data_ptr(), max(j, 1), ctx.stream()))
        ctx.launches += 1
    return out


def cost_value(ctx: nat.Context, cost: nat.PlsCost, y: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    n, j = f.shape
    tiles = (n + nat.COST_VALUE_TILE_ROWS - 1) // nat.COST_VALUE_TILE_ROWS
    partial = torch.empty((max(tiles, 1), j), dtype=F64, device=f.device)
    out = torch.empty((j,), dtype=F64, device=f.device)
    ctx.check(ctx.lib.pls_cost_value_f64(ctx.handle, C.byref(cost), y.data_ptr(), f.data_ptr(), _ld(f), n, j,
                                         partial.data_ptr(), out.data_ptr(), ctx.stream()))
    ctx.launches += 2
    return out


def energy_terms(ctx: nat.Context, partial: torch.Tensor, j: int, p: Optional[torch.Tensor],
                 inv_lambda: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty((j,), dtype=F64, device=partial.device)
    ctx.check(ctx.lib.pls_energy_terms_f64(ctx.handle, partial.data_ptr(), partial.shape[0], _ld(partial), nat.ptr(p),
                                           _ld(p) if p is not None else 0, p.shape[0] if p is not None else 0,
                                           nat.ptr(inv_lambda), j, out.data_ptr(), ctx.stream()))
    ctx.launches += 1
    return out


def philox_normal(ctx: nat.Context, seed: int, step: int, rows: int, j: int, j_global_offset: int = 0,
                  device=None) -> torch.Tensor:
    out = torch.empty((rows, j), dtype=F64, device=device or torch.device("cuda", ctx.device_index))
    ctx.check(ctx.lib.pls_philox_normal_f64(ctx.handle, seed & (2**64 - 1), step & (2**64 - 1), rows, j, j_global_offset,
                                            out.data_ptr(), max(j, 1), ctx.stream()))
    ctx.launches += 1
    return out


def gram_exp(ctx: nat.Context, x: torch.Tensor, fast: bool) -> torch.Tensor:
    """exp(x) with the hot loop's table-driven routine (fast=True) or the CUDA library exp (fast=False)."""
    out = torch.empty_like(x)
    ctx.check(ctx.lib.pls_gram_exp_f64(ctx.handle, x.data_ptr(), x.numel(), int(fast), out.data_ptr(), ctx.stream()))
    ctx.launches += 1
    return out


def lincomb3(ctx: nat.Context, a: float, x: torch.Tensor, b: float, y: torch.Tensor, c: float, z: torch.Tensor, j: int,
             out: torch.Tensor, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[:, :j] = (base[:, :j] if base is given) + a x + b y + c z."""
    ctx.check(ctx.lib.pls_lincomb3_f64(ctx.handle, x.shape[0], j, float(a), x.data_ptr(), _ld(x), float(b), y.data_ptr(), _ld(y),
                                       float(c), z.data_ptr(), _ld(z), nat.ptr(base), _ld(base) if base is not None else 0,
                                       out.data_ptr(), _ld(out), ctx.stream()))
    ctx.launches += 1
    return out


def flat_math(ctx: nat.Context, op: int, a: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a / b (op 0), log a (op 1), exp a (op 2) with the register epilogue's branch-free routines (tests)."""
    out = torch.empty_like(a)
    ctx.check(ctx.lib.pls_flat_math_f64(ctx.handle, op, a.data_ptr(), nat.ptr(b), a.numel(), out.data_ptr(), ctx.stream()))
    ctx.launches += 1
    return out


def numpy_tie_rule(di, chosen) -> int:
    """The reference's choice among exactly tied conditional variances, verbatim: the last entry of numpy's DEFAULT (unstable)
    argsort of d that has not been chosen yet (src/inducing_point_selectors/conditional_variance.py:105-109).  Runs on the host
    copy of d the selector hands over when -- and only when -- the maximum is attained by several points."""
    import numpy as np

    taken = set(int(c) for c in chosen)
    for next_idx in reversed(np.argsort(di)):
        if int(next_idx) not in taken:
            return int(next_idx)
    return -1


TIE_RULES = ("numpy", "stable")


def _tie_mode(tie_rule) -> int:
    if callable(tie_rule) or tie_rule == "numpy":
        return nat.CV_TIES_HOST
    if tie_rule == "stable":
        return nat.CV_TIES_HIGHEST_INDEX
    raise ValueError(f"tie_rule must be one of {TIE_RULES} or a callable (d, chosen) -> index, got {tie_rule!r}")


def cv_select(ctx: nat.Context, kernel_id: int, xp_aug: torch.Tensor, d: int, kdiag: float, m: int, jitter: float,
              threshold: Optional[float], tie_rule="numpy", info: Optional[dict] = None) -> Tuple[torch.Tensor, int]:
    """Returns (indices into the permuted order (m,), number selected).
    tie_rule: "numpy" (default: exact ties are resolved on the host by numpy_tie_rule, i.e. as the reference does), "stable"
    (highest permuted index, decided on the device) or a callable (d_host: np.ndarray, chosen: np.ndarray) -> index.
    info (optional dict) receives min_top2_rel_gap, tied_picks and host_tie_calls."""
    import numpy as np

    n = xp_aug.shape[0]
    dev = xp_aug.device
    ci = torch.empty((m - 1, n), dtype=F64, device=dev)
    di = torch.empty((n,), dtype=F64, device=dev)
    scratch = torch.zeros((int(ctx.lib.pls_cv_scratch_doubles(n, d, m)),), dtype=F64, device=dev)
    indices = torch.full((m,), n, dtype=torch.int64, device=dev)  # sentinel N, as conditional_variance.py:63
    nsel = C.c_int(0)
    mode = _tie_mode(tie_rule)
    rule = tie_rule if callable(tie_rule) else numpy_tie_rule
    calls, failure = [0], []

    def on_tie(_user, d_host, n_host, chosen, n_chosen):
        try:
            calls[0] += 1
            d_arr = np.ctypeslib.as_array(d_host, shape=(int(n_host),))
            c_arr = np.ctypeslib.as_array(chosen, shape=(int(n_chosen),)) if n_chosen else np.zeros((0,), dtype=np.int64)
            return int(rule(d_arr, c_arr))
        except BaseException as exc:  # an exception must not unwind through the C frames
            failure.append(exc)
            return -1

    callback = nat.CV_TIE_FN(on_tie)
    rc = ctx.lib.pls_cv_select_f64(ctx.handle, kernel_id, xp_aug.data_ptr(), n, d, float(kdiag), m, float(jitter),
                                   float(threshold) if threshold is not None else 0.0, int(threshold is not None), mode, callback,
                                   None, ci.data_ptr(), di.data_ptr(), scratch.data_ptr(), indices.data_ptr(), C.byref(nsel),
                                   ctx.stream())
    if failure:
        raise failure[0]
    ctx.check(rc)
    ctx.launches += 2 * m
    if info is not None:
        hdr = scratch[: nat.CV_HEADER_DOUBLES].cpu()
        info.update(min_top2_rel_gap=float(hdr[8]), tied_picks=int(hdr.view(torch.int64)[9]), host_tie_calls=calls[0],
                    trace=float(hdr[4]))
    return indices, int(nsel.value)


class ShardedSelectorState:
    """One rank's workspaces of the row-sharded ConditionalVariance selector (pls_cv_shard_*)."""

    def __init__(self, ctx: nat.Context, kernel_id: int, xa_local: torch.Tensor, n_offset: int, n_total: int, d: int, kdiag: float,
                 m: int, jitter: float, threshold: Optional[float], tie_rule="numpy"):
        self.ctx, self.kernel_id, self.xa, self.n_offset, self.d, self.m = ctx, kernel_id, xa_local, int(n_offset), d, m
        self.kdiag, self.jitter = float(kdiag), float(jitter)
        self.threshold, self.has_threshold = (float(threshold), 1) if threshold is not None else (0.0, 0)
        self.tie_mode = _tie_mode(tie_rule)
        self.tie_rule = tie_rule if callable(tie_rule) else numpy_tie_rule
        self.n_local = xa_local.shape[0]
        self.n_total = int(n_total)
        dev = xa_local.device
        self.ci = torch.empty((m - 1, max(self.n_local, 1)), dtype=F64, device=dev)
        self.di = torch.empty((max(self.n_local, 1),), dtype=F64, device=dev)
        self.scratch = torch.zeros((int(ctx.lib.pls_cv_shard_scratch_doubles(self.n_local, d, m)),), dtype=F64, device=dev)
        self.record = int(ctx.lib.pls_cv_candidate_doubles(d, m))
        self.candidate = torch.zeros((self.record,), dtype=F64, device=dev)
        self.indices = torch.full((m,), int(n_total), dtype=torch.int64, device=dev)  # sentinel N (conditional_variance.py:63)

    def begin(self) -> torch.Tensor:
        c = self.ctx
        c.check(c.lib.pls_cv_shard_begin_f64(c.handle, self.kernel_id, self.xa.data_ptr(), self.n_local, self.n_offset, self.d,
                                             self.kdiag, self.m, self.jitter, self.di.data_ptr(), self.scratch.data_ptr(),
                                             self.candidate.data_ptr(), c.stream()))
        c.launches += 2
        return self.candidate

    def pick(self, candidates: torch.Tensor, slot: int, forced: bool = False) -> None:
        c = self.ctx
        world = candidates.numel() // self.record
        c.check(c.lib.pls_cv_shard_pick_f64(c.handle, candidates.data_ptr(), world, slot, self.d, self.m, self.threshold,
                                            self.has_threshold, self.tie_mode, int(forced), self.n_local, self.n_offset,
                                            self.scratch.data_ptr(), self.indices.data_ptr(), c.stream()))
        c.launches += 1

    def update(self, iteration: int) -> torch.Tensor:
        c = self.ctx
        c.check(c.lib.pls_cv_shard_update_f64(c.handle, self.kernel_id, self.xa.data_ptr(), self.n_local, self.n_offset, self.d,
                                              iteration, self.m, self.jitter, self.ci.data_ptr(), self.di.data_ptr(),
                                              self.scratch.data_ptr(), self.candidate.data_ptr(), c.stream()))
        c.launches += 2
        return self.candidate

    def force(self, slot: int, pivot: int) -> torch.Tensor:
        """The candidate record of the point with GLOBAL permuted index `pivot` (an empty record on the ranks that do not hold it)."""
        c = self.ctx
        c.check(c.lib.pls_cv_shard_force_f64(c.handle, self.xa.data_ptr(), self.n_local, self.n_offset, self.d, self.m, slot, int(pivot),
                                             self.ci.data_ptr(), self.di.data_ptr(), self.scratch.data_ptr(), self.candidate.data_ptr(),
                                             c.stream()))
        c.launches += 1
        return self.candidate

    def status(self) -> Tuple[int, bool, bool, int]:
        """(n_selected, stopped, tie pending, slot of the tie); synchronises."""
        c = self.ctx
        st = (C.c_int64 * 4)()
        c.check(c.lib.pls_cv_shard_status(c.handle, self.scratch.data_ptr(), st, c.stream()))
        return int(st[0]), bool(st[1]), bool(st[2]), int(st[3])

    def finish(self) -> int:
        c = self.ctx
        nsel = C.c_int(0)
        c.check(c.lib.pls_cv_shard_finish(c.handle, self.scratch.data_ptr(), C.byref(nsel), c.stream()))
        return int(nsel.value)


def cv_select_sharded(states: Sequence[ShardedSelectorState], gather, gather_d=None) -> Tuple[torch.Tensor, int]:
    """Drives the sharded selector.  `states` are the ranks handled by THIS process (one in production; several when a
    test emulates the ranks on one GPU); `gather(list of this process's candidate records)` returns all ranks' records
    concatenated in rank order (torch.distributed.all_gather_into_tensor in production).  `gather_d(list of this process's
    slices of d)` returns the WHOLE d (all ranks, rank order) as a host numpy array; it is called only when the maximum is
    attained by several points and the tie rule is the host's (every rank then applies the same rule to the same array)."""
    m = states[0].m
    host_ties = states[0].tie_mode == nat.CV_TIES_HOST
    cands = gather([s.begin() for s in states])
    for s in states:
        s.pick(cands, 0)
    i, batch = 0, (8 if host_ties else m)
    while i < m - 1:
        end = min(m - 1, i + batch)
        while i < end:
            cands = gather([s.update(i) for s in states])
            for s in states:
                s.pick(cands, i + 1)
            i += 1
        if not host_ties:
            break
        _, stopped, tie, slot = states[0].status()
        if tie:
            if gather_d is None:
                raise RuntimeError("cv_select_sharded: a tie has to be resolved on the host but no gather_d was given")
            d_host = gather_d([s.di[: s.n_local] for s in states])
            chosen = states[0].indices[:slot].cpu().numpy()
            pivot = int(states[0].tie_rule(d_host, chosen))
            if pivot < 0 or pivot >= states[0].n_total or pivot in set(chosen.tolist()):
                raise RuntimeError(f"cv_select_sharded: the tie rule returned {pivot}, which is out of range or already chosen")
            cands = gather([s.force(slot, pivot) for s in states])
            for s in states:
                s.pick(cands, slot, forced=True)
            if stopped:
                break
            i, batch = slot, 1
        else:
            if stopped:
                break
            batch = min(2 * batch, 64)
    nsel = [s.finish() for s in states]
    return states[0].indices, nsel[0]
