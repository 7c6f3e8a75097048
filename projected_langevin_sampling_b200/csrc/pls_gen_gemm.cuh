// The hot kernel of the PLS Langevin step: a GEMM whose A operand is GENERATED, never stored.
//
//     C[r][j] (+)= sum_k  kappa(row_r, red_k) * B[k][j]
//
//   forward role   rows = training points x_n, reduction = inducing points z_m, B = W = V~ P  (M x J)
//                  -> F = k(X,Z) W, then the cost epilogue in registers            (orthonormal.py:98-108 + costs/*.py)
//   backward role  rows = inducing points z_m, reduction = training points x_n, B = d_2 c  (N x J), split over n
//                  -> G'[m][j] = sum_n k(z_m, x_n) Dc[n][j]                        (orthonormal.py:151-155)
//
// Design (B200, sm_100a):
//   * FP64 has no tcgen05 kind; the FP64 tensor instruction is DMMA.8x8x4 (every wider PTX f64 mma shape lowers to it)
//     and it shares one 64-FMA/clk/SM pipe with DFMA (measured, profiles/fp64_microbench_r01.txt), so every FP64 op
//     spent on generating K is taken from the GEMM.  A generated K element costs ~4*ceil(D/4) FMA slots of exponent
//     DMMA + ~10 slots of exp and is reused across the BJ columns of the tile, so the tile is as WIDE in J as the
//     register file allows: 64 fp64 accumulators per thread (half the SM's registers) arranged as RT row tiles x
//     32/RT column tiles per warp.  RT = 1 -> CTA tile 64 rows x 256 particles (generation overhead ~18/256),
//     RT = 2 -> 128 x 128 (~18/128; kept for narrow particle slices).  Warp w owns rows [8 RT w, 8 RT (w+1)) and ALL
//     columns, so the A fragments it needs are the ones it generates, in registers: no shared-memory round trip and
//     no __syncthreads in the main loop.
//   * the exponent tile itself is a DMMA: S = c_row + c_point + x~ . z~ (= -|x~ - z~|^2/2 + log sigma^2) comes out in
//     C-fragment layout; the thread that holds S[g][2t], S[g][2t+1] uses exp of them as the A fragments of two k4
//     steps whose k index t maps to reduction points 2t and 2t+1 (the B rows are addressed accordingly): no shuffle.
//   * B (W or Dc rows) and the reduction-point rows are staged by the TMA engine (cp.async.bulk -> UBLKCP) through a
//     3-stage mbarrier pipeline; every warp's lane 0 issues a share, all 8 warps consume; rows are padded by 2 doubles
//     so the LDS.128 B-fragment reads are bank-conflict free.
//   * the loop over 8-point groups is flat and branch-free (kernel kind is a template parameter) so ptxas can interleave
//     the exponent chain of the NEXT k4 step with the 32 DMMAs of the current one.
#pragma once
#include "pls_cost.cuh"
#include "pls_internal.h"

namespace pls {

namespace {

constexpr int SMEM_HEADER = 128 + 64 * 8;  // mbarriers + exp table
constexpr int NWARPS = NTHREADS / 32;

template <int RT>
struct Tile {
  static constexpr int NT = 32 / RT;          // n8 column tiles per warp
  static constexpr int BR = tile_rows(RT);    // rows per CTA
  static constexpr int BJ = tile_cols(RT);    // columns per CTA
  static constexpr int SB = BJ + 2;           // smem row stride of the streamed tile
  static constexpr int NPR = NT / 2;          // column pairs (one LDS.128 each)
};

template <int RT>
__host__ __device__ inline size_t gen_gemm_smem_bytes(int sp) {
  // pipeline buffers; the forward epilogue re-uses them to stage the BR x BJ output tile
  const size_t pipeline = (size_t)(STAGES * BK * Tile<RT>::SB + STAGES * BK * sp);
  const size_t staging = (size_t)(Tile<RT>::BR * Tile<RT>::SB);
  return SMEM_HEADER + sizeof(double) * (pipeline > staging ? pipeline : staging);
}

template <int NKD, bool BACKWARD, bool RBF, int RT>
__global__ void __launch_bounds__(NTHREADS, 1) gen_gemm_kernel(const GenGemmParams p) {
  using T = Tile<RT>;
  constexpr int NT = T::NT, BR = T::BR, BJ = T::BJ, SB = T::SB, NPR = T::NPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  double* sExp = reinterpret_cast<double*>(smem_raw + 128);        // 2^(j/64)
  double* sB = reinterpret_cast<double*>(smem_raw + SMEM_HEADER);  // [STAGES][BK][SB]
  double* sP = sB + STAGES * BK * SB;                              // [STAGES][BK][sp]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int g = lane >> 2;  // DMMA group id
  const int t = lane & 3;   // DMMA thread-in-group
  const int sp = p.sp;

  // ---- which tile / which slice of the reduction ---------------------------------------------------------------
  const int64_t n_row_tiles = (p.n_rows + BR - 1) / BR;
  const int64_t n_col_tiles = (p.j + BJ - 1) / BJ;
  int64_t bid = blockIdx.x;
  int64_t rt, ct;
  int split = 0;
  if (!BACKWARD) {  // particles fastest: neighbouring CTAs share the same training rows
    ct = bid % n_col_tiles;
    rt = bid / n_col_tiles;
  } else {  // inducing-row tiles fastest: the CTAs that stream the same Dc slab run together and share it in L2
    rt = bid % n_row_tiles;
    bid /= n_row_tiles;
    ct = bid % n_col_tiles;
    split = (int)(bid / n_col_tiles);
  }
  const int64_t row0 = rt * BR;
  const int64_t j0 = ct * BJ;

  int64_t begin = 0, end = p.red_total;
  if (BACKWARD) {
    const int64_t total_chunks = (p.red_total + BK - 1) / BK;
    const int64_t per = (total_chunks + p.splits - 1) / p.splits;
    begin = (int64_t)split * per * BK;
    end = begin + per * BK;
    if (end > p.red_total) end = p.red_total;
    if (begin > end) begin = end;
  }
  const int red_len = (int)(end - begin);
  const int nchunks = (red_len + BK - 1) / BK;
  const int total_groups = (red_len + 7) >> 3;  // BK is a multiple of 8: groups never straddle chunks

  // ---- one-time setup -------------------------------------------------------------------------------------------
  // zero the staging buffers: rows never written by a copy (reduction tail) must hold finite values
  for (int i = tid; i < STAGES * BK * SB + STAGES * BK * sp; i += NTHREADS) sB[i] = 0.0;
  if (tid < 64) sExp[tid] = kExp2Table[tid];
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], NWARPS);
      mbar_init(&empty[s], NWARPS);
    }
    fence_mbar_init();
  }
  fence_proxy_async();  // generic-proxy zero fill ordered before the async-proxy bulk copies
  __syncthreads();

  // row-side exponent fragments (A operand of the S DMMA): row g of each of this warp's RT row tiles.  Coordinates only;
  // the c entry (column d of the augmented row) seeds the accumulator together with the point's c.
  double a2[RT][NKD];
  double crow[RT];
#pragma unroll
  for (int h = 0; h < RT; ++h) {
    const int64_t r = row0 + (warp * RT + h) * 8 + g;
    const bool rv = r < p.n_rows;
#pragma unroll
    for (int kd = 0; kd < NKD; ++kd) {
      const int dd = t + 4 * kd;
      a2[h][kd] = (rv && dd < p.d) ? p.rows_aug[r * sp + dd] : 0.0;
    }
    crow[h] = rv ? p.rows_aug[r * sp + p.d] : 0.0;
  }

  const int64_t cw = (p.ldb - j0 < BJ) ? (p.ldb - j0) : BJ;  // columns copied per row (ldb even => 16-byte multiple)
  // Every warp's lane 0 issues its share of a stage (rows warp, warp + 8, ... of the streamed tile; warp 0 adds the point
  // rows) and posts the bytes it issued on the stage's full barrier (8 arrivals), so no single warp carries the copy
  // issue cost on its critical path.
  auto issue = [&](int c) {
    const int stage = c % STAGES;
    const int64_t k0 = begin + (int64_t)c * BK;
    const int kc = (int)((end - k0 < BK) ? (end - k0) : BK);
    uint64_t* bar = &full[stage];
    const int my_rows = (kc > warp) ? ((kc - warp + NWARPS - 1) / NWARPS) : 0;
    uint32_t bytes = (uint32_t)(my_rows * cw * 8);
    if (warp == 0) bytes += (uint32_t)(kc * sp * 8);
    mbar_expect_tx(bar, bytes);
    double* dst = sB + stage * BK * SB;
    const double* src = p.b + k0 * p.ldb + j0;
    for (int r = warp; r < kc; r += NWARPS) bulk_g2s(dst + r * SB, src + (int64_t)r * p.ldb, (uint32_t)(cw * 8), bar);
    if (warp == 0) bulk_g2s(sP + stage * BK * sp, p.red_aug + k0 * sp, (uint32_t)(kc * sp * 8), bar);
  };

  double acc[RT][NT][2];
#pragma unroll
  for (int h = 0; h < RT; ++h)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      acc[h][nt][0] = 0.0;
      acc[h][nt][1] = 0.0;
    }

  if (lane == 0) {
    for (int c = 0; c < STAGES - 1 && c < nchunks; ++c) issue(c);
  }

  // ---- main loop over groups of 8 reduction points ------------------------------------------------------------------
  // Software-pipelined at k4-step granularity: while the DMMAs of one k4 step (one reduction point per thread) are
  // issued, the Gram values of the NEXT k4 step are generated (S DMMAs + exp), so the latency of the exponent chain
  // hides behind this warp's own tensor work.
  // exponents of group `grp` of the point tile Pt: S[h] (8 rows x 8 points) in C-fragment layout
  auto exponent_tile = [&](const double* Pt, int grp, double (&s)[RT][2]) {
    const double* prow = Pt + (grp * 8 + g) * sp;          // B fragment: point g of the group, coordinate t + 4 kd
    const double* pc = Pt + (grp * 8 + 2 * t) * sp + p.d;  // c entries of this thread's two points
    const double c0 = pc[0], c1 = pc[sp];
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      s[h][0] = crow[h] + c0;
      s[h][1] = crow[h] + c1;
    }
#pragma unroll
    for (int kd = 0; kd < NKD; ++kd) {
      const double b2 = prow[t + 4 * kd];
#pragma unroll
      for (int h = 0; h < RT; ++h) dmma(s[h][0], s[h][1], a2[h][kd], b2);
    }
  };
  auto gram_value = [&](double sv, bool valid) {
    const double v = RBF ? gram_exp_fast(sv, sExp) : sv;
    return valid ? v : 0.0;
  };
  auto mma_step = [&](const double* brow, const double (&ka)[RT]) {
#pragma unroll
    for (int pr = 0; pr < NPR; ++pr) {
      const double2 bv = *reinterpret_cast<const double2*>(brow + 16 * pr);
#pragma unroll
      for (int h = 0; h < RT; ++h) {
        dmma(acc[h][2 * pr][0], acc[h][2 * pr][1], ka[h], bv.x);
        dmma(acc[h][2 * pr + 1][0], acc[h][2 * pr + 1][1], ka[h], bv.y);
      }
    }
  };

  double s[RT][2];  // exponents of the group in flight
  double k0[RT];    // Gram values of k4 step 0 (point 2t)
#pragma unroll
  for (int h = 0; h < RT; ++h) k0[h] = 0.0;
  if (nchunks > 0) {
    mbar_wait(&full[0], 0u);
    exponent_tile(sP, 0, s);
#pragma unroll
    for (int h = 0; h < RT; ++h) k0[h] = gram_value(s[h][0], 2 * t < red_len);
  }
#pragma unroll 1
  for (int gi = 0; gi < total_groups; ++gi) {
    const int c = gi >> 2;  // BK / 8 = 4 groups per chunk
    const int grp = gi & 3;
    const int stage = c % STAGES;
    // the group after this one: the next chunk's stage must have landed before its exponents are formed
    const int gn = (gi + 1 < total_groups) ? gi + 1 : gi;
    const int cn = gn >> 2;
    const int nstage = cn % STAGES;
    if (cn != c) mbar_wait(&full[nstage], ((uint32_t)(cn / STAGES)) & 1u);

    const int p0 = grp * 8 + 2 * t;  // this thread's two reduction points: p0 (k4 step 0) and p0 + 1 (k4 step 1)
    const int pg = gi * 8 + 2 * t;   // ... counted from `begin`
    const double* b0 = sB + (stage * BK + p0) * SB + 2 * g;
    // k4 step 0 with k0; meanwhile the Gram values of step 1
    double k1[RT];
#pragma unroll
    for (int h = 0; h < RT; ++h) k1[h] = gram_value(s[h][1], pg + 1 < red_len);
    mma_step(b0, k0);
    // k4 step 1 with k1; meanwhile exponents + step-0 Gram values of the next group
    exponent_tile(sP + nstage * BK * sp, gn & 3, s);
    const bool nvalid = (gn != gi) && (gn * 8 + 2 * t < red_len);
#pragma unroll
    for (int h = 0; h < RT; ++h) k0[h] = gram_value(s[h][0], nvalid);
    mma_step(b0 + SB, k1);

    if (grp == 3 || gi + 1 == total_groups) {  // chunk c fully consumed by this warp
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&empty[stage]);
        // refill the stage consumed in chunk c-1 with chunk c + STAGES - 1: every warp has long left that stage, so the
        // wait does not stall, and the copies have one full chunk of compute to land
        const int cf = c + STAGES - 1;
        if (cf < nchunks) {
          mbar_wait(&empty[cf % STAGES], (((uint32_t)(cf / STAGES)) & 1u) ^ 1u);
          issue(cf);
        }
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------
  // Column map: thread (g,t) holds, for column pair pr, the 4 consecutive columns j0 + 16 pr + 4 t + {0,1,2,3} as
  // acc[h][2pr][0], acc[h][2pr+1][0], acc[h][2pr][1], acc[h][2pr+1][1]; rows row0 + 8 (RT warp + h) + g.
  if (BACKWARD) {
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      const int64_t r = row0 + (warp * RT + h) * 8 + g;
      if (r >= p.n_rows) continue;
      double* orow = p.out + ((int64_t)split * p.n_rows + r) * p.ldo;
#pragma unroll
      for (int pr = 0; pr < NPR; ++pr) {
        const int64_t col = j0 + 16 * pr + 4 * t;
        double v[4] = {acc[h][2 * pr][0], acc[h][2 * pr + 1][0], acc[h][2 * pr][1], acc[h][2 * pr + 1][1]};
        if (col + 3 < p.j) {
          double2* dst = reinterpret_cast<double2*>(orow + col);
          if (p.accumulate) {
            const double2 o0 = dst[0], o1 = dst[1];
            v[0] += o0.x;
            v[1] += o0.y;
            v[2] += o1.x;
            v[3] += o1.y;
          }
          dst[0] = make_double2(v[0], v[1]);
          dst[1] = make_double2(v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < p.j) orow[col + e] = p.accumulate ? orow[col + e] + v[e] : v[e];
        }
      }
    }
    return;
  }

  // Forward role: stage the F tile in shared memory (the pipeline buffers are idle: every issued copy has landed and been
  // consumed), then one coalesced pass applies the cost functor -- inlined once, in a rolled loop with independent
  // evaluations in flight -- and writes whole 128-byte lines.
  __syncthreads();
  double* sF = sB;  // [BR][SB]
#pragma unroll
  for (int h = 0; h < RT; ++h) {
    double* frow = sF + ((warp * RT + h) * 8 + g) * SB + 4 * t;
#pragma unroll
    for (int pr = 0; pr < NPR; ++pr) {
      double2* dst = reinterpret_cast<double2*>(frow + 16 * pr);
      dst[0] = make_double2(acc[h][2 * pr][0], acc[h][2 * pr + 1][0]);
      dst[1] = make_double2(acc[h][2 * pr][1], acc[h][2 * pr + 1][1]);
    }
  }
  __syncthreads();
  const int64_t rows_here = (p.n_rows - row0 < BR) ? (p.n_rows - row0) : BR;
  const int64_t cols_here = (p.j - j0 < BJ) ? (p.j - j0) : BJ;

  if (p.epilogue == PLS_EPI_COST) {
    // per-column sum over this tile's rows (increasing row order) of c(y_n, F[n][j]) -> out[rt][j]
    if (tid < cols_here) {
      double v = 0.0;
      for (int r = 0; r < (int)rows_here; ++r) v += cost_value(p.cost, p.y[row0 + r], sF[r * SB + tid]);
      p.out[rt * p.ldo + j0 + tid] = v;
    }
    return;
  }

  const bool dcost = (p.epilogue == PLS_EPI_COST_DERIVATIVE);
  constexpr int PAIRS = BJ / 2;  // double2 per row
#pragma unroll 2
  for (int idx = tid; idx < BR * PAIRS; idx += NTHREADS) {
    const int r = idx / PAIRS;
    const int col = 2 * (idx - r * PAIRS);
    if (r >= rows_here || col >= cols_here) continue;
    double2 v = *reinterpret_cast<const double2*>(sF + r * SB + col);
    if (dcost) {
      const double yv = p.y[row0 + r];
      v.x = cost_derivative(p.cost, yv, v.x);
      v.y = cost_derivative(p.cost, yv, v.y);
    }
    double* dst = p.out + (row0 + r) * p.ldo + j0 + col;
    if (col + 1 < cols_here) *reinterpret_cast<double2*>(dst) = v;
    else dst[0] = v.x;
  }
}

template <int NKD, bool BACKWARD, bool RBF, int RT>
cudaError_t launch_one(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  using T = Tile<RT>;
  int64_t grid = ((p.n_rows + T::BR - 1) / T::BR) * ((p.j + T::BJ - 1) / T::BJ);
  if (BACKWARD) grid *= p.splits;
  if (grid <= 0) return cudaSuccess;
  if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
  const size_t smem = gen_gemm_smem_bytes<RT>(p.sp);
  if ((int64_t)smem > ctx->max_smem_optin) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(gen_gemm_kernel<NKD, BACKWARD, RBF, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gen_gemm_kernel<NKD, BACKWARD, RBF, RT><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int NKD, bool BACKWARD>
cudaError_t launch_kind(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  const bool rbf = (p.kernel_id == PLS_KERNEL_RBF);
  if (p.rt == 1) return rbf ? launch_one<NKD, BACKWARD, true, 1>(ctx, p, stream) : launch_one<NKD, BACKWARD, false, 1>(ctx, p, stream);
  return rbf ? launch_one<NKD, BACKWARD, true, 2>(ctx, p, stream) : launch_one<NKD, BACKWARD, false, 2>(ctx, p, stream);
}

}  // namespace

}  // namespace pls
