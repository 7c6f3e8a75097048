// Internal (non-ABI) declarations shared by the csrc translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pls_b200.h"

struct CUtensorMap_st;  // <cuda.h>

struct pls_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  const uint64_t* step_counter = nullptr;  // device counter added to pls_project_update_f64's `step` (pls_set_step_counter)
  int tile_rt = 0;  // 0 = choose per launch from the particle count; 1 / 2 force a tile shape (PLS_B200_TILE_RT, tests)
  int fused_functor = 1;  // fused training epilogue: one call for cost values + derivatives (0: two calls; PLS_B200_FUSED_FUNCTOR, A/B runs)
  int cluster = 0;  // pairs of CTAs sharing the Gram generation: 0 = rule of choose_cluster (backward role only), 1 = never, 2 = both roles (PLS_B200_CLUSTER)
  int tile_ns = 0;  // 0 = park a second accumulator set in tensor memory whenever the shape allows; 1 = never (PLS_B200_TILE_NS, tests)
  std::string error;
  // pls_profile_begin / pls_profile_end: CUDA-event pairs around every launch of the hot kernel (role 0 forward, 1 backward)
  struct ProfileRecord {
    int role;
    double flops;
    cudaEvent_t e0, e1;
  };
  bool profiling = false;
  std::vector<ProfileRecord> profile;
};

namespace pls {

// Parameters of the generated-operand GEMM  C[r][j] (+)= sum_k kappa(row_r, red_k) * B[k][j]
struct GenGemmParams {
  const double* rows_aug;  // row-side augmented points (n_rows x sp)
  int64_t n_rows;
  const double* red_aug;  // reduction-side augmented points (red_total x sp)
  int64_t red_total;
  const double* gram;  // optional cached Gram (training rows x inducing columns, ldk): replaces the generated values
  int64_t ldk;
  const double* b;  // streamed matrix (red_total x ldb)
  int64_t ldb;
  int64_t j;  // valid columns
  int sp;
  int d;
  int kernel_id;
  int epilogue;  // PLS_EPI_* for the forward role; ignored by the backward role
  int splits;    // backward role: number of reduction splits (gridDim = tiles * splits)
  int accumulate;
  int rt;  // tile shape: row tiles per warp (1: 64 x 256 CTA tile, 2: 128 x 128)
  int both;             // set by the launcher: the fused derivative + cost epilogue evaluates both in one functor call
  int wbuf_ok;          // set by the launcher: shared memory has room for the per-warp cost sums of the register epilogue
  int tma3d;            // set by the launcher: the 3-D tensor map of the streamed matrix is usable
  int64_t full_blocks;  // set by the launcher: complete 16-column blocks per row of the streamed matrix (ldb / 16)
  double* out;
  int64_t ldo;
  double* out2;  // forward, PLS_EPI_COST_DERIVATIVE_AND_COST: per-row-tile cost sums (tiles x ldo2)
  int64_t ldo2;
  const double* y;
  pls_cost cost;
};

// tile shape for a launch over j particle columns: wide tiles unless the slice is narrow
inline int choose_tile_rt(const pls_ctx* ctx, int64_t j) {
  if (ctx && (ctx->tile_rt == 1 || ctx->tile_rt == 2)) return ctx->tile_rt;
  return (j <= 128) ? 2 : 1;
}

// accumulator sets per CTA tile of the generated-Gram kernels: 2 (64 x 512 tile, the second set parked in tensor memory) when the
// particle slice is an even number of 256-column tiles AND the launch is large enough for the wider tile to pay -- a tile of at
// least 8 chunks (forward: M >= 256) and enough 64 x 512 tiles (times the splits the reduction allows) to keep every SM busy;
// measured: C2 (M = 64, J = 1024) 4.66 M particle-updates/s with 256-column tiles against 3.35 M with 512-column ones, C3
// (M = 256) and larger the other way round.  out_rows x j is the output, red_len the reduction length.  The cached-Gram kernels
// generate nothing and stay at 1.  ctx->tile_ns: 0 = this rule, 1 = never, 2 = whenever the shape allows it (tests).
inline int choose_tile_ns(const pls_ctx* ctx, int64_t j, bool cached, int64_t out_rows, int64_t red_len, bool backward) {
  if (cached || choose_tile_rt(ctx, j) != 1 || (ctx && ctx->tile_ns == 1)) return 1;
  if (((j + 255) / 256) % 2 != 0) return 1;
  if (ctx && ctx->tile_ns == 2) return 2;
  const int64_t sms = (ctx && ctx->sm_count > 0) ? ctx->sm_count : 148;
  const int64_t chunks = (red_len + 31) / 32;
  const int64_t tiles = ((out_rows + 63) / 64) * ((j + 511) / 512);
  if (!backward) return (chunks >= 8 && tiles >= 2 * sms) ? 2 : 1;
  const int64_t max_splits = chunks / 8 > 1 ? chunks / 8 : 1;
  return (tiles * max_splits >= 8 * sms) ? 2 : 1;
}

// CTAs per cluster of the generated-Gram kernels (on top of NS = 2): 2 = pairs of CTAs on adjacent 512-column tiles share the
// generation of the Gram values through distributed shared memory.  Needs a whole number of 1024-column cluster tiles.  Measured at
// C4 (tools/cluster_compare.sh): backward 33.62 -> 33.83 TFLOP/s, persistent forward 32.73 -> 32.47 (tying warp w of one CTA to warp
// w of its peer costs the forward the free drift of its epilogues), so the default rule (ctx->cluster == 0) pairs the BACKWARD
// role only, and only when there are enough cluster tiles x splits for 8 waves; 1 = never, 2 = both roles whenever possible.
inline int choose_cluster(const pls_ctx* ctx, int64_t j, int64_t out_rows, int64_t red_len, bool backward) {
  if (!ctx || ctx->cluster == 1 || ((j + 255) / 256) % 4 != 0) return 1;
  if (ctx->cluster == 2) return 2;
  if (!backward) return 1;
  const int64_t sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int64_t chunks = (red_len + 31) / 32;
  const int64_t max_splits = chunks / 8 > 1 ? chunks / 8 : 1;
  return (((out_rows + 63) / 64) * ((j + 1023) / 1024) * max_splits >= 8 * sms) ? 2 : 1;
}

// Tensor maps of the streamed matrix b (rows x ldb doubles, 16-byte aligned, ldb even) for the hot kernel's stages of 32 rows
// x (16 * blocks) columns, 128-byte swizzle: tm3 = [ldb/16 column blocks][rows][16] (one copy per stage), tm2 = [rows][ldb]
// with a 32 x 16 box (one copy per column block; handles a partial last block).  *tma3d_ok = 0 if the driver rejects tm3.
cudaError_t make_stream_maps(const pls_ctx* ctx, const double* b, int64_t rows, int64_t ldb, int blocks, ::CUtensorMap_st* tm3,
                             ::CUtensorMap_st* tm2, int* tma3d_ok);

cudaError_t launch_gen_gemm_forward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream);
cudaError_t launch_gen_gemm_backward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream);

}  // namespace pls
