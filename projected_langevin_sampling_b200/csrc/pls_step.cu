// One C-ABI call per Langevin step: pls_step_plan_f64 / pls_grad_f64 / pls_step_f64 (include/pls_b200.h).
//
// Host code only: the launch sequence of one step over caller-owned workspaces, built from the library's own entry points
//     W  = V~ P                                  pls_gemm_f64
//     for each chunk of training rows:
//         [K_c = k(X_c, Z)                       pls_gram_fill_f64, staged Gram only]
//         Dc = d_2 c(y, k(X_c, Z) W)             pls_forward[_step][_cached]_f64
//         Gp (+)= k(Z, X_c) Dc                   pls_backward[_cached]_f64
//     G' = sum_s Gp[s]                           pls_reduce_splits_f64
//     P += -eta V~^T G' - eta P / lambda + sqrt(2 eta) xi      pls_project_update_f64   (pls_step_f64 only)
// i.e. PLS.calculate_particle_update (src/projected_langevin_sampling/projected_langevin_sampling.py:107-123 ->
// basis/orthonormal.py:98-108, costs/*.py, orthonormal.py:128-159) without a host round trip between the pieces.
#include <cstring>

#include "pls_aux.h"
#include "pls_common.cuh"

namespace {
constexpr int64_t ROW_ALIGN = 128;
inline int64_t even64(int64_t v) { return v + (v & 1); }
inline int64_t align256(int64_t bytes) { return (bytes + 255) / 256 * 256; }
int fail_plan(pls_ctx* ctx, const char* msg) {
  if (ctx) ctx->error = msg;
  return 1;
}
}  // namespace

extern "C" {

int pls_step_plan_f64(const pls_ctx* ctx, int64_t n, int64_t m, int64_t m_k, int64_t j, int64_t dc_budget_bytes, int gram_mode,
                      int with_cost, pls_step_plan* plan) {
  if (!plan || n < 0 || m < 1 || m_k < 0 || j < 1 || gram_mode < PLS_GRAM_GENERATED || gram_mode > PLS_GRAM_CACHED) return 1;
  std::memset(plan, 0, sizeof(*plan));
  if (dc_budget_bytes <= 0) dc_budget_bytes = 32LL << 30;
  plan->n = n;
  plan->m = m;
  plan->m_k = m_k;
  plan->j = j;
  plan->ldj = even64(j);
  plan->gram_mode = gram_mode;
  plan->with_cost = with_cost != 0;
  int64_t rows = (dc_budget_bytes / (plan->ldj * 8)) / ROW_ALIGN * ROW_ALIGN;
  if (gram_mode == PLS_GRAM_STAGED) {
    // the staging buffer must stay chunk-sized ("nothing N x M kept"): at most 2 GiB of Gram values per chunk
    const int64_t staged_rows = ((2LL << 30) / (pls_gram_cache_ld(m) * 8)) / ROW_ALIGN * ROW_ALIGN;
    if (rows > staged_rows) rows = staged_rows;
  }
  if (rows < ROW_ALIGN) rows = ROW_ALIGN;
  plan->chunk_rows = n < rows ? n : rows;
  plan->n_chunks = plan->chunk_rows > 0 ? (int32_t)((n + plan->chunk_rows - 1) / plan->chunk_rows) : 0;
  plan->splits = pls_backward_splits(ctx, plan->chunk_rows, m, j);
  plan->tile_rows = pls_forward_tile_rows(ctx, j);
  int64_t tiles = 0;
  for (int64_t r0 = 0; r0 < n; r0 += plan->chunk_rows) {
    const int64_t r = (n - r0 < plan->chunk_rows) ? (n - r0) : plan->chunk_rows;
    tiles += (r + plan->tile_rows - 1) / plan->tile_rows;
  }
  plan->cost_tiles = tiles;
  int64_t off = 0;
  plan->off_w = off;
  off += align256(m * plan->ldj * 8);
  plan->off_gm = off;
  off += align256(m * plan->ldj * 8);
  plan->off_dc = off;
  off += align256((plan->chunk_rows > 0 ? plan->chunk_rows : 1) * plan->ldj * 8);
  plan->off_gp = off;
  off += align256((int64_t)plan->splits * m * plan->ldj * 8);
  plan->off_cost_partial = off;
  off += align256((with_cost ? (tiles > 0 ? tiles : 1) : 0) * plan->ldj * 8);
  plan->off_kstage = off;
  if (gram_mode == PLS_GRAM_STAGED && n > 0) off += align256(pls_gram_cache_rows(plan->chunk_rows) * pls_gram_cache_ld(m) * 8);
  plan->workspace_bytes = off;
  return 0;
}

int pls_grad_f64(pls_ctx* ctx, const pls_step_plan* plan, int kernel_id, int d, const double* xa, const double* za, const double* vt,
                 int64_t ldv, const double* p, int64_t ldp, const pls_cost* cost, const double* y, const double* gram, int64_t ldk,
                 void* workspace, double* cost_sums, void* stream) {
  if (!ctx) return 1;
  if (!plan || !workspace || !p || !cost || (!y && plan->n > 0)) return fail_plan(ctx, "pls_grad_f64: NULL argument");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail_plan(ctx, "pls_grad_f64: the workspace must be 256-byte aligned");
  if (cost_sums && !plan->with_cost) return fail_plan(ctx, "pls_grad_f64: cost sums need a plan made with with_cost != 0");
  if (plan->gram_mode == PLS_GRAM_CACHED && !gram) return fail_plan(ctx, "pls_grad_f64: PLS_GRAM_CACHED needs the Gram cache");
  char* ws = static_cast<char*>(workspace);
  double* w = reinterpret_cast<double*>(ws + plan->off_w);
  double* gm = reinterpret_cast<double*>(ws + plan->off_gm);
  double* dc = reinterpret_cast<double*>(ws + plan->off_dc);
  double* gp = reinterpret_cast<double*>(ws + plan->off_gp);
  double* cpart = reinterpret_cast<double*>(ws + plan->off_cost_partial);
  double* kstage = reinterpret_cast<double*>(ws + plan->off_kstage);
  const int64_t n = plan->n, m = plan->m, j = plan->j, ldj = plan->ldj;
  const int sp = pls_point_stride(d);
  const double* wsrc = w;
  int64_t ldw = ldj;
  if (vt) {  // W = V~ P (orthonormal.py:106-108, re-associated)
    if (pls_gemm_f64(ctx, 0, vt, ldv, p, ldp, w, ldj, m, j, plan->m_k, stream)) return 1;
  } else {  // the caller's `p` already is W (M x J): InducingPointBasis passes k(Z,Z)^{-1} P
    wsrc = p;
    ldw = ldp;
  }
  if (n == 0) {  // an empty row shard contributes a zero gradient
    cudaError_t e = cudaMemsetAsync(gm, 0, sizeof(double) * (size_t)m * (size_t)ldj, (cudaStream_t)stream);
    if (e == cudaSuccess && cost_sums) e = cudaMemsetAsync(cost_sums, 0, sizeof(double) * (size_t)j, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_plan(ctx, cudaGetErrorString(e));
    return 0;
  }
  const int64_t ldks = pls_gram_cache_ld(m);
  int64_t t0 = 0;
  int chunk = 0;
  for (int64_t r0 = 0; r0 < n; r0 += plan->chunk_rows, ++chunk) {
    const int64_t rows = (n - r0 < plan->chunk_rows) ? (n - r0) : plan->chunk_rows;
    const double* xc = xa ? xa + r0 * sp : nullptr;
    const double* k = nullptr;
    int64_t kld = 0;
    if (plan->gram_mode == PLS_GRAM_STAGED) {
      if (pls_gram_fill_f64(ctx, kernel_id, xc, rows, za, m, d, kstage, ldks, stream)) return 1;
      k = kstage;
      kld = ldks;
    } else if (plan->gram_mode == PLS_GRAM_CACHED) {
      k = gram + r0 * ldk;
      kld = ldk;
    }
    const int64_t tiles = (rows + plan->tile_rows - 1) / plan->tile_rows;
    int rc;
    if (cost_sums) {
      double* cp = cpart + t0 * ldj;
      rc = k ? pls_forward_step_cached_f64(ctx, k, kld, rows, m, wsrc, ldw, j, cost, y + r0, dc, ldj, cp, ldj, stream)
             : pls_forward_step_f64(ctx, kernel_id, xc, rows, za, m, d, wsrc, ldw, j, cost, y + r0, dc, ldj, cp, ldj, stream);
    } else {
      rc = k ? pls_forward_cached_f64(ctx, k, kld, rows, m, wsrc, ldw, j, PLS_EPI_COST_DERIVATIVE, cost, y + r0, dc, ldj, stream)
             : pls_forward_f64(ctx, kernel_id, xc, rows, za, m, d, wsrc, ldw, j, PLS_EPI_COST_DERIVATIVE, cost, y + r0, dc, ldj, stream);
    }
    if (rc) return 1;
    t0 += tiles;
    rc = k ? pls_backward_cached_f64(ctx, k, kld, m, rows, dc, ldj, j, gp, ldj, plan->splits, chunk > 0, stream)
           : pls_backward_f64(ctx, kernel_id, za, m, xc, rows, d, dc, ldj, j, gp, ldj, plan->splits, chunk > 0, stream);
    if (rc) return 1;
  }
  if (pls_reduce_splits_f64(ctx, gp, plan->splits, m, j, ldj, gm, ldj, stream)) return 1;
  if (cost_sums && pls_energy_terms_f64(ctx, cpart, plan->cost_tiles, ldj, nullptr, 0, 0, nullptr, j, cost_sums, stream)) return 1;
  return 0;
}

int pls_step_f64(pls_ctx* ctx, const pls_step_plan* plan, int kernel_id, int d, const double* xa, const double* za, const double* vt,
                 int64_t ldv, const double* inv_lambda, double* p, int64_t ldp, const pls_cost* cost, const double* y,
                 const double* gram, int64_t ldk, double eta, int noise_mode, const double* xi, int64_t ldxi, uint64_t seed,
                 uint64_t step, int64_t j_global_offset, int in_place, double* out, int64_t ldo, double* energy_out, void* workspace,
                 void* stream) {
  if (!ctx) return 1;
  if (!plan || !vt || !inv_lambda || !out) return fail_plan(ctx, "pls_step_f64: NULL argument");
  char* ws = static_cast<char*>(workspace);
  if (pls_grad_f64(ctx, plan, kernel_id, d, xa, za, vt, ldv, p, ldp, cost, y, gram, ldk, workspace, energy_out, stream)) return 1;
  if (energy_out) {  // energy of the INPUT particles: cost sums + 1/2 sum_m P_mj^2 / lambda_m (orthonormal.py:110-126 before the mean)
    if (pls_energy_terms_f64(ctx, energy_out, 1, plan->ldj > plan->j ? plan->j : plan->ldj, p, ldp, plan->m_k, inv_lambda, plan->j,
                             energy_out, stream))
      return 1;
  }
  const double* gm = reinterpret_cast<const double*>(ws + plan->off_gm);
  return pls_project_update_f64(ctx, vt, ldv, plan->m, plan->m_k, gm, plan->ldj, p, ldp, plan->j, inv_lambda, eta, noise_mode, xi, ldxi,
                                seed, step, j_global_offset, in_place, out, ldo, stream);
}

}  // extern "C"
