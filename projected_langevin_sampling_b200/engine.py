"""Launch sequence and workspaces of the fused Langevin step on one GPU.

One step (reference: PLS.calculate_particle_update, src/projected_langevin_sampling/projected_langevin_sampling.py:107-123
-> orthonormal.py:98-108, costs/*.py, orthonormal.py:128-159) becomes

    W  = V~ P                                  pls_gemm_f64            (M x J; 2 M M_k J flops)
    for each chunk of training rows:
        Dc = d_2 c(y, k(X_c, Z) W)             pls_forward[_cached]_f64   (Gram tiles generated on the fly or loaded, cost in registers)
        Gp += k(Z, X_c) Dc                     pls_backward[_cached]_f64  (likewise, split over rows)
    G' = sum_s Gp[s]                           pls_reduce_splits_f64   (deterministic order)
    [all-reduce G' over the N-shard group]     torch.distributed / NCCL, only when the training rows are sharded
    P += -eta V~^T G' - eta P / lambda + sqrt(2 eta) xi      pls_project_update_f64

The N x J prediction matrix is never materialised; the only N x J intermediate is the Dc chunk (`dc_budget_bytes`, default
32 GiB or 40 % of the free device memory), written and read once per step (~2 % of the step time at the headline shape).  The N x M Gram is generated inside the
kernels from the points (default: nothing N x M in memory) or -- opt-in `gram_cache`, when N x M doubles fit (the reference
keeps k(Z, X) for the whole run, orthonormal.py:36-41) -- computed once and streamed, which takes the exponent work off the
FP64 pipe (+9 % at the headline shape for 8.2 GB).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Tuple, Union

import torch

from . import _native as nat
from . import ops

# Bound on the Dc chunk, the only N-sized intermediate.  32 GiB holds the whole headline problem (N = 1M rows x J = 4096 particles
# = 30.5 GiB) in ONE forward + ONE backward launch per step: measured 8056 -> 8089 -> 8113 particle-updates/s at 8 / 16 / 33 GiB
# (4 / 2 / 1 chunks per step; tools/dc_budget_compare.sh) -- fewer launch tails, longer backward splits.  Clamped to 40 % of the free
# device memory when the workspace is planned (a B200 has 180 GB).
DEFAULT_DC_BUDGET = int(os.environ.get("PLS_B200_DC_BUDGET_GIB", "32")) << 30
ROW_ALIGN = 128
DEFAULT_GRAM_CACHE_BYTES = 24 << 30


def gram_mode(mode: Union[bool, str, None]) -> Union[bool, str]:
    """Normalised Gram mode: False (generated, the default), True (cached), "auto" or "staged"; the environment variable
    PLS_B200_GRAM_CACHE (off / on / auto / staged) overrides the argument."""
    mode = os.environ.get("PLS_B200_GRAM_CACHE", mode if mode is not None else False)
    if mode in (True, "on", "1", "true", "cached"):
        return True
    if mode in (False, "off", "0", "false", "generated"):
        return False
    if mode in ("auto", "staged"):
        return mode
    raise ValueError(f"unknown Gram mode {mode!r} (False / True / 'auto' / 'staged')")


def want_gram_cache(mode: Union[bool, str, None], ctx: nat.Context, n: int, m: int, device: torch.device) -> bool:
    """Policy for keeping k(X, Z) resident.  mode: False / "off" (the default: Gram tiles are regenerated inside the kernels and
    nothing N x M is ever in memory, as BASELINE.json's north_star specifies), True / "on", or "auto": cache when it takes at
    most PLS_B200_GRAM_CACHE_BYTES (24 GiB) and a third of the free device memory.  The environment variable
    PLS_B200_GRAM_CACHE overrides the argument."""
    mode = gram_mode(mode)
    if mode is True or mode is False:
        return mode
    if mode == "staged":
        return False
    nbytes = int(ctx.lib.pls_gram_cache_rows(n)) * int(ctx.lib.pls_gram_cache_ld(m)) * 8
    free, _ = torch.cuda.mem_get_info(device)
    return n > 0 and m > 0 and nbytes <= int(os.environ.get("PLS_B200_GRAM_CACHE_BYTES", DEFAULT_GRAM_CACHE_BYTES)) and nbytes <= free // 3


class LangevinEngine:
    def __init__(self, ctx: nat.Context, kernel_id: int, d: int, xa: torch.Tensor, za: torch.Tensor, vt: torch.Tensor,
                 inv_lambda: torch.Tensor, j: int, dc_budget_bytes: int = DEFAULT_DC_BUDGET,
                 gradient_reduce: Optional[Callable[[torch.Tensor], None]] = None,
                 weights_fn: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None, gram: Optional[torch.Tensor] = None,
                 gaussian_normal_equations: bool = False, gram_staged: bool = False):
        self.ctx, self.kernel_id, self.d = ctx, kernel_id, d
        self.gram = gram  # ops.gram_cache(...) of (xa, za) or None: Gram values loaded instead of generated
        # "staged" Gram: nothing N x M is kept -- every step re-forms k(X_c, Z) for the row chunk in flight into ONE chunk-sized
        # buffer (pls_gram_fill_f64, ~1 % of the step) which that chunk's forward and backward launches then stream, instead of
        # each of their J/256 column tiles regenerating it
        self.kstage: Optional[torch.Tensor] = None
        self._gram_staged = gram_staged and gram is None
        self.gaussian_normal_equations = gaussian_normal_equations  # opt-in, see _normal_equations()
        self._neq = None  # (key, A' in Gram-cache layout, b', y^T y / (2 s))
        self._neq_y: Optional[torch.Tensor] = None
        self.xa, self.za, self.vt, self.inv_lambda = xa, za, vt, inv_lambda
        self.n, self.m, self.m_k, self.j = xa.shape[0], za.shape[0], vt.shape[1], j
        self.gradient_reduce = gradient_reduce
        self.weights_fn = weights_fn  # fills W (M x J) from the particles; default W = V~ P (OrthonormalBasis)
        dev = xa.device
        self.dc_budget_bytes = int(dc_budget_bytes)
        self.workspace: Optional[torch.Tensor] = None
        self._gram_code = nat.GRAM_CACHED if gram is not None else (nat.GRAM_STAGED if self._gram_staged and self.n > 0 else nat.GRAM_GENERATED)
        self._plan_workspace(with_cost=False)
        self._neq_cost: Optional[torch.Tensor] = None
        self._zeros: Optional[torch.Tensor] = None
        self._cost_sums: Optional[torch.Tensor] = None

    def _plan_workspace(self, with_cost: bool) -> None:
        """pls_step_plan_f64 fixes the row chunking and the workspace layout; ONE torch allocation backs W, G', the Dc chunk,
        the split partials, the cost-sum region (only once an energy has been asked for: 0.5 GB at the headline shape) and the
        Gram staging chunk.  The tensors below are views into it (the piecewise entry points and the tests use them)."""
        ctx, dev = self.ctx, self.xa.device
        plan = nat.StepPlan()
        if self.dc_budget_bytes >= DEFAULT_DC_BUDGET:  # the default (or more): never more than 40 % of what the device has free
            free, _ = torch.cuda.mem_get_info(dev)
            if self.workspace is not None:
                free += self.workspace.numel() * 8  # re-planning (a cost region is being added): the old workspace is released below
            self.dc_budget_bytes = max(min(self.dc_budget_bytes, int(0.4 * free)), 1 << 28)
        self.workspace = None
        if ctx.lib.pls_step_plan_f64(ctx.handle, self.n, self.m, self.m_k, self.j, self.dc_budget_bytes, self._gram_code, int(with_cost),
                                     C.byref(plan)) != 0:
            raise ValueError(f"pls_step_plan_f64 rejected the shape N={self.n} M={self.m} M_k={self.m_k} J={self.j}")
        self.plan = plan
        self.ldj = int(plan.ldj)
        self.chunk_rows = int(plan.chunk_rows)
        self.chunks: List[Tuple[int, int]] = [(r, min(r + self.chunk_rows, self.n)) for r in range(0, self.n, max(self.chunk_rows, 1))]
        self.splits = int(plan.splits)
        self.tile_rows = int(plan.tile_rows)
        self.workspace = torch.zeros((int(plan.workspace_bytes) // 8,), dtype=torch.float64, device=dev)
        assert self.workspace.data_ptr() % 256 == 0

        def view(offset: int, rows: int, cols: int) -> torch.Tensor:
            return self.workspace[offset // 8: offset // 8 + rows * cols].view(rows, cols)

        self.w = view(plan.off_w, self.m, self.ldj)
        self.gm = view(plan.off_gm, self.m, self.ldj)
        self.dc = view(plan.off_dc, max(self.chunk_rows, 1), self.ldj)[: self.chunk_rows]
        self.gp = view(plan.off_gp, self.splits * self.m, self.ldj).view(self.splits, self.m, self.ldj)
        self.cost_partial = view(plan.off_cost_partial, max(int(plan.cost_tiles), 1), self.ldj) if with_cost else None
        self.kstage = None
        if self._gram_code == nat.GRAM_STAGED:
            self.kstage = view(plan.off_kstage, int(ctx.lib.pls_gram_cache_rows(self.chunk_rows)), int(ctx.lib.pls_gram_cache_ld(self.m)))

    def _launches_per_gradient(self, with_cost: bool) -> int:
        per_chunk = 2 + (1 if self._gram_code == nat.GRAM_STAGED else 0)
        return (1 if self.weights_fn is None else 0) + (per_chunk * len(self.chunks) + 1 + (1 if with_cost else 0) if self.n > 0 else 1)

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def _gram(self, r0: int = 0) -> Optional[torch.Tensor]:
        return None if self.gram is None else self.gram[r0:]

    def _weights(self, particles: torch.Tensor) -> torch.Tensor:
        if self.weights_fn is not None:
            self.weights_fn(particles, self.w)  # e.g. W = k(Z, Z)^{-1} P for the InducingPointBasis
            return self.w
        return ops.gemm(self.ctx, self.vt, particles, self.w)  # W = V~ P

    def gradient(self, particles: torch.Tensor, cost: nat.PlsCost, y: torch.Tensor, with_cost: bool = False) -> torch.Tensor:
        """G' = k(Z, X) d_2 c(y, k(X, Z) V~ P)  -> (M, ldj) workspace view.  with_cost also leaves the per-row-tile cost
        sums of the SAME forward pass in self.cost_partial (the energy potential costs no second forward)."""
        if self.gaussian_normal_equations and self._is_gaussian_identity(cost):
            self._weights(particles)
            return self._gradient_normal_equations(cost, y, with_cost)
        self._neq_cost = None
        if with_cost and self.cost_partial is None:
            self._plan_workspace(with_cost=True)
        ctx = self.ctx
        if self.weights_fn is not None:
            self.weights_fn(particles, self.w)  # e.g. W = k(Z, Z)^{-1} P for the InducingPointBasis
            vt_ptr, ldv, p_src = None, 0, self.w
        else:
            vt_ptr, ldv, p_src = self.vt.data_ptr(), ops._ld(self.vt), particles
        if with_cost and self._cost_sums is None:
            self._cost_sums = torch.zeros((self.j,), dtype=torch.float64, device=self.xa.device)
        # one call: W = V~ P, the row-chunk loop (forward with the cost epilogue, backward), the split reduction (pls_grad_f64)
        ctx.check(ctx.lib.pls_grad_f64(ctx.handle, C.byref(self.plan), self.kernel_id, self.d, self.xa.data_ptr(), self.za.data_ptr(),
                                       vt_ptr, ldv, p_src.data_ptr(), ops._ld(p_src), C.byref(cost), nat.ptr(y),
                                       nat.ptr(self.gram), ops._ld(self.gram) if self.gram is not None else 0,
                                       self.workspace.data_ptr(), self._cost_sums.data_ptr() if with_cost else None, ctx.stream()))
        ctx.launches += self._launches_per_gradient(with_cost)
        if self.gradient_reduce is not None:
            self.gradient_reduce(self.gm)
        return self.gm

    # ---- Gaussian / identity shortcut (opt-in) ---------------------------------------------------------------------------
    @staticmethod
    def _is_gaussian_identity(cost: nat.PlsCost) -> bool:
        return cost.cost_id == nat.COST_GAUSSIAN and cost.link_id == nat.LINK_IDENTITY and cost.closed_form != 0

    def _normal_equations(self, cost: nat.PlsCost, y: torch.Tensor):
        """For the Gaussian cost with the identity link the gradient is LINEAR in W:
               G' = k(Z,X) (k(X,Z) W - y) / s = A' W - b' 1^T,   A' = k(Z,X) k(X,Z) / s  (M x M),  b' = k(Z,X) y / s  (M),
        and the cost sums are quadratic: sum_n (F_nj - y_n)^2 / (2 s) = 1/2 w_j^T A' w_j - b'^T w_j + y^T y / (2 s).
        A' and b' are formed ONCE (one backward contraction with the Gram as the streamed matrix: 2 N M^2 flops), after which a
        step costs 2 M^2 J flops instead of 4 N M J.  This is a re-association of the reference's algebra
        (orthonormal.py:98-108,151-158 with costs/gaussian.py:75-88), exact up to round-off, NOT the general path: bench.py's
        headline never uses it."""
        # keyed on the label tensor's identity AND its version counter (an in-place edit of y bumps it); the entry keeps a
        # reference to y so the allocator cannot hand the same address to a different label tensor while the entry lives
        key = (y.data_ptr(), y.shape[0], y._version, float(cost.observation_noise))
        if self._neq is not None and self._neq[0] == key and self._neq_y is y:
            return self._neq
        self._neq_y = y
        ctx, m, dev = self.ctx, self.m, self.xa.device
        s_obs = float(cost.observation_noise)
        a_pad = torch.zeros((int(ctx.lib.pls_gram_cache_rows(m)), int(ctx.lib.pls_gram_cache_ld(m))), dtype=torch.float64, device=dev)
        ldm = a_pad.shape[1]
        rows = 65536
        splits = ops.backward_splits(ctx, min(rows, self.n), m, m)
        gp_a = torch.zeros((splits, m, ldm), dtype=torch.float64, device=dev)
        gp_b = torch.zeros((splits, m, 16), dtype=torch.float64, device=dev)
        y2 = torch.zeros((self.n, 16), dtype=torch.float64, device=dev)  # y as column 0 of a one-block-wide streamed matrix
        y2[:, 0] = y
        for ci, r0 in enumerate(range(0, self.n, rows)):
            r1 = min(self.n, r0 + rows)
            k = self.gram[r0:] if self.gram is not None else ops.gram_cache(ctx, self.kernel_id, self.xa[r0:r1], self.za, self.d)
            # the Gram chunk is both the (loaded) left operand and the streamed matrix: Gp += k(Z, X_c) k(X_c, Z)
            ops.backward(ctx, self.kernel_id, self.za, self.xa[r0:r1], self.d, k[: r1 - r0], m, gp_a, splits, accumulate=ci > 0, gram=k)
            ops.backward(ctx, self.kernel_id, self.za, self.xa[r0:r1], self.d, y2[r0:r1], 1, gp_b, splits, accumulate=ci > 0, gram=k)
        a = torch.zeros((m, ldm), dtype=torch.float64, device=dev)
        b2 = torch.zeros((m, 16), dtype=torch.float64, device=dev)
        ops.reduce_splits(ctx, gp_a, m, a)
        ops.reduce_splits(ctx, gp_b, 1, b2)
        b = b2[:, 0].contiguous()
        yy = (y * y).sum().reshape(1)
        if self.gradient_reduce is not None:  # rows sharded: A', b' and y^T y are sums over the row group -- reduced once, not per step
            self.gradient_reduce(a)
            self.gradient_reduce(b)
            self.gradient_reduce(yy)
        a_pad[:m, :m] = a[:, :m] / s_obs
        self._neq = (key, a_pad, (b / s_obs).contiguous(), yy / (2.0 * s_obs))
        return self._neq

    def _gradient_normal_equations(self, cost: nat.PlsCost, y: torch.Tensor, with_cost: bool) -> torch.Tensor:
        _, a_pad, b, yy = self._normal_equations(cost, y)
        # G' = A' W through the cached-Gram forward kernel (A' plays the Gram: M "training rows" x M inducing points)
        ops.forward(self.ctx, self.kernel_id, self.za, self.za, self.d, self.w, self.j, nat.EPI_PREDICTION, self.gm, gram=a_pad)
        g = self.gm[:, : self.j]
        w = self.w[:, : self.j]
        self._neq_cost = (0.5 * (w * g).sum(0) - b @ w + yy) if with_cost else None  # before b' is subtracted: g = A' W
        g.sub_(b[:, None])
        return self.gm

    def step(self, particles: torch.Tensor, eta: float, cost: nat.PlsCost, y: torch.Tensor, out: torch.Tensor,
             noise_mode: int, xi: Optional[torch.Tensor] = None, seed: int = 0, step_index: int = 0,
             j_global_offset: int = 0, in_place: bool = False) -> torch.Tensor:
        if self.gradient_reduce is None and self.weights_fn is None and not (self.gaussian_normal_equations and self._is_gaussian_identity(cost)):
            # single GPU / particle-sharded: the whole step is ONE C call (pls_step_f64)
            ctx = self.ctx
            ctx.check(ctx.lib.pls_step_f64(ctx.handle, C.byref(self.plan), self.kernel_id, self.d, self.xa.data_ptr(), self.za.data_ptr(),
                                           self.vt.data_ptr(), ops._ld(self.vt), self.inv_lambda.data_ptr(), particles.data_ptr(),
                                           ops._ld(particles), C.byref(cost), nat.ptr(y), nat.ptr(self.gram),
                                           ops._ld(self.gram) if self.gram is not None else 0, float(eta), noise_mode, nat.ptr(xi),
                                           ops._ld(xi) if xi is not None else 0, seed & (2**64 - 1), step_index & (2**64 - 1),
                                           j_global_offset, int(in_place), out.data_ptr(), ops._ld(out), None,
                                           self.workspace.data_ptr(), ctx.stream()))
            ctx.launches += self._launches_per_gradient(False) + 1
            self._neq_cost = None
            return out
        gm = self.gradient(particles, cost, y)
        return ops.project_update(self.ctx, self.vt, gm, particles, self.j, self.inv_lambda, eta, out, noise_mode=noise_mode,
                                  xi=xi, seed=seed, step=step_index, j_global_offset=j_global_offset, in_place=in_place)

    def energy_and_gradient(self, particles: torch.Tensor, cost: nat.PlsCost, y: torch.Tensor) -> torch.Tensor:
        """One forward + backward: leaves G' in self.gm and returns the per-particle energy c_j + 1/2 sum_m P_mj^2 / lambda_m
        of `particles` (J,) (PLS.calculate_energy_potential before its mean, orthonormal.py:110-126).  With a `weights_fn`
        (InducingPointBasis) the prior term is taken on W = k(Z, Z)^{-1} P, which the gradient has just left in self.w:
        c_j + M/2 sum_m W_mj^2 (inducing_point.py:97-119; `inv_lambda` then holds the constant M)."""
        self.gradient(particles, cost, y, with_cost=True)
        prior_of = self.w[:, : self.j] if self.weights_fn is not None else particles
        if self._neq_cost is not None:  # Gaussian shortcut: the cost sums are complete on every rank already
            return ops.energy_terms(self.ctx, self._neq_cost.reshape(1, -1).contiguous(), self.j, prior_of, self.inv_lambda)
        c = self._cost_sums  # sum_n c(y_n, F[n][j]) of the SAME forward pass, written by pls_grad_f64
        if self.gradient_reduce is not None:  # rows are sharded: the cost sums are partial over this rank's rows, the prior term is not
            c = c.clone()
            self.gradient_reduce(c)
        return ops.energy_terms(self.ctx, c.reshape(1, -1), self.j, prior_of, self.inv_lambda)

    def _zero_row(self) -> torch.Tensor:
        if self._zeros is None:
            self._zeros = torch.zeros((1, self.ldj), dtype=torch.float64, device=self.xa.device)
        return self._zeros

    def apply_update(self, particles: torch.Tensor, eta: float, out: torch.Tensor, noise_mode: int,
                     xi: Optional[torch.Tensor] = None, seed: int = 0, step_index: int = 0, j_global_offset: int = 0,
                     in_place: bool = False) -> torch.Tensor:
        """The Langevin update from the gradient already in self.gm (second half of `step`)."""
        return ops.project_update(self.ctx, self.vt, self.gm, particles, self.j, self.inv_lambda, eta, out, noise_mode=noise_mode,
                                  xi=xi, seed=seed, step=step_index, j_global_offset=j_global_offset, in_place=in_place)

    def prediction(self, particles: torch.Tensor) -> torch.Tensor:
        """F = k(X, Z) V~ P  -> (N, J)  (materialised: API parity / small problems only)."""
        self._weights(particles)
        out, _ = ops.alloc_matrix(self.n, self.j, self.xa.device)
        ops.forward(self.ctx, self.kernel_id, self.xa, self.za, self.d, self.w, self.j, nat.EPI_PREDICTION, out, gram=self._gram())
        return out[:, : self.j]

    def cost_derivative(self, particles: torch.Tensor, cost: nat.PlsCost, y: torch.Tensor) -> torch.Tensor:
        self._weights(particles)
        out, _ = ops.alloc_matrix(self.n, self.j, self.xa.device)
        ops.forward(self.ctx, self.kernel_id, self.xa, self.za, self.d, self.w, self.j, nat.EPI_COST_DERIVATIVE, out, cost=cost, y=y,
                    gram=self._gram())
        return out[:, : self.j]

    def cost_partials(self, particles: torch.Tensor, cost: nat.PlsCost, y: torch.Tensor) -> torch.Tensor:
        """per-row-tile column sums of c(y, F) -> (tiles, J) without materialising F."""
        self._weights(particles)
        tile_rows = ops.forward_tile_rows(self.ctx, self.j)
        tiles = (self.n + tile_rows - 1) // tile_rows
        part = torch.empty((max(tiles, 1), self.ldj), dtype=torch.float64, device=self.xa.device)
        ops.forward(self.ctx, self.kernel_id, self.xa, self.za, self.d, self.w, self.j, nat.EPI_COST, part, cost=cost, y=y, gram=self._gram())
        return part

    def cost(self, particles: torch.Tensor, cost: nat.PlsCost, y: torch.Tensor) -> torch.Tensor:
        c = ops.energy_terms(self.ctx, self.cost_partials(particles, cost, y), self.j, None, None)
        if self.gradient_reduce is not None:  # rows sharded: sum the per-particle cost over the row group
            self.gradient_reduce(c)
        return c

    def backproject_update(self, particles: torch.Tensor, cost_derivative: torch.Tensor, eta: float, out: torch.Tensor,
                           noise_mode: int, xi: Optional[torch.Tensor]) -> torch.Tensor:
        """Update from a caller-provided Dc (N x J): OrthonormalBasis._calculate_particle_update."""
        dc = cost_derivative
        if dc.stride(1) != 1 or (dc.stride(0) & 1) or (dc.data_ptr() & 15):
            buf, _ = ops.alloc_matrix(self.n, self.j, dc.device)
            buf[:, : self.j].copy_(dc)
            dc = buf[:, : self.j]
        splits = ops.backward_splits(self.ctx, self.n, self.m, self.j)
        gp = torch.empty((splits, self.m, self.ldj), dtype=torch.float64, device=dc.device)
        ops.backward(self.ctx, self.kernel_id, self.za, self.xa, self.d, dc, self.j, gp, splits, accumulate=False, gram=self._gram())
        ops.reduce_splits(self.ctx, gp, self.j, self.gm)
        if self.gradient_reduce is not None:
            self.gradient_reduce(self.gm)
        return ops.project_update(self.ctx, self.vt, self.gm, particles, self.j, self.inv_lambda, eta, out,
                                  noise_mode=noise_mode, xi=xi)
