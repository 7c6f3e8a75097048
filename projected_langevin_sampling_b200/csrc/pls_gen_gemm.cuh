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
//   * FP64 has no tcgen05 kind; the FP64 tensor instruction is DMMA.8x8x4 and it shares one 64-FMA/clk/SM pipe with
//     DFMA (measured, profiles/fp64_microbench_r01.txt), so every FP64 op spent on generating K is taken from the
//     GEMM.  The tile is therefore as wide in J as the register file allows (128 x 128 fp64 accumulators = half the
//     SM's registers) and each K element is generated exactly once per CTA: warp w owns rows [16w, 16w+16) and ALL
//     128 columns, so the A fragments it needs are the ones it generates, in registers, with no shared-memory round
//     trip and no __syncthreads in the main loop.
//   * the exponent tile itself is a DMMA: rows and reduction points are stored "augmented"
//     ([x~ | c | 1] . [z~ | 1 | c']), so S = A2 * B2^T gives -|x~ - z~|^2/2 + log(sigma^2) directly in C-fragment
//     layout; the thread that holds S[g][2t], S[g][2t+1] uses them as the A fragments of two k4 steps whose k index
//     t maps to reduction points 2t and 2t+1 (the B rows are addressed accordingly), so no shuffle is needed either.
//   * B (W or Dc rows) and the reduction-point rows are staged by the TMA engine (cp.async.bulk -> UBLKCP) through a
//     3-stage mbarrier pipeline; one elected thread issues, all 8 warps consume; rows are padded to 130 doubles so the
//     LDS.128 B-fragment reads are bank-conflict free.
#pragma once
#include "pls_cost.cuh"
#include "pls_internal.h"

namespace pls {

namespace {

constexpr int SMEM_HEADER = 128 + 64 * 8;  // mbarriers + exp table
constexpr int NWARPS = NTHREADS / 32;

// The cost functors are called (not inlined) from the tile epilogue: 64 calls per thread per 128 x 128 x K tile is
// noise next to the main loop, and it keeps the kernel's code small.
static __device__ __noinline__ double cost_derivative_call(const pls_cost& c, double y, double f) { return cost_derivative(c, y, f); }
static __device__ __noinline__ double cost_value_call(const pls_cost& c, double y, double f) { return cost_value(c, y, f); }

__host__ __device__ inline size_t gen_gemm_smem_bytes(int sp) {
  return SMEM_HEADER + sizeof(double) * (size_t)(STAGES * BK * SB + STAGES * BK * sp);
}

template <int NKD, bool BACKWARD>
__global__ void __launch_bounds__(NTHREADS, 1) gen_gemm_kernel(const GenGemmParams p) {
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
  const int nchunks = (int)((end - begin + BK - 1) / BK);

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

  // row-side exponent fragments (A operand of the S DMMA): rows g and g+8 of this warp's 16 rows
  double a2[2][NKD];
  int pcol[NKD];
#pragma unroll
  for (int kd = 0; kd < NKD; ++kd) {
    const int dd = t + 4 * kd;
    pcol[kd] = (dd == p.d) ? p.d + 1 : ((dd == p.d + 1) ? p.d : dd);  // reduction side swaps the c / 1 entries
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = row0 + warp * 16 + g + 8 * h;
      a2[h][kd] = (r < p.n_rows) ? p.rows_aug[r * sp + dd] : 0.0;
    }
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

  double acc[2][16][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      acc[h][nt][0] = 0.0;
      acc[h][nt][1] = 0.0;
    }

  if (lane == 0) {
    for (int c = 0; c < STAGES - 1 && c < nchunks; ++c) issue(c);
  }

  const bool rbf = (p.kernel_id == PLS_KERNEL_RBF);

  // ---- main loop over reduction chunks ----------------------------------------------------------------------------
  // Software-pipelined at half-group granularity: while the 32 DMMAs of one k4 step (one reduction point per thread)
  // are issued, the Gram values of the NEXT k4 step are generated (S DMMAs + exp), so the latency of the exponent chain
  // hides behind this warp's own tensor work instead of relying on the other warp of the scheduler.
  auto exponent_tile = [&](const double* Pt, int grp, double& s00, double& s01, double& s10, double& s11) {
    // S (16 rows x 8 points) = rows . points^T over the augmented coordinates
    s00 = 0.0; s01 = 0.0; s10 = 0.0; s11 = 0.0;
    const double* prow = Pt + (grp * 8 + g) * sp;
#pragma unroll
    for (int kd = 0; kd < NKD; ++kd) {
      const double b2 = prow[pcol[kd]];
      dmma(s00, s01, a2[0][kd], b2);
      dmma(s10, s11, a2[1][kd], b2);
    }
  };
  auto gram_value = [&](double s, bool valid) {
    const double v = rbf ? gram_exp_fast(s, sExp) : s;
    return valid ? v : 0.0;
  };
  auto mma_step = [&](const double* brow, double ka, double kb) {
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
      const double2 bv = *reinterpret_cast<const double2*>(brow + 16 * pr);
      dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], ka, bv.x);
      dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], kb, bv.x);
      dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], ka, bv.y);
      dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], kb, bv.y);
    }
  };

  double s00, s01, s10, s11;  // exponents of the group in flight
  double k0a = 0.0, k0b = 0.0;  // Gram values of k4 step 0 (point 2t) for rows g, g+8
  if (nchunks > 0) {
    mbar_wait(&full[0], 0u);
    const int kc0 = (int)((end - begin < BK) ? (end - begin) : BK);
    exponent_tile(sP, 0, s00, s01, s10, s11);
    k0a = gram_value(s00, 2 * t < kc0);
    k0b = gram_value(s10, 2 * t < kc0);
  }
  for (int c = 0; c < nchunks; ++c) {
    const int stage = c % STAGES;
    const int64_t k0 = begin + (int64_t)c * BK;
    const int kc = (int)((end - k0 < BK) ? (end - k0) : BK);
    const int ngroups = (kc + 7) >> 3;
    const double* Bt = sB + stage * BK * SB;
    const double* Pt = sP + stage * BK * sp;

#pragma unroll 1
    for (int grp = 0; grp < ngroups; ++grp) {
      const int p0 = grp * 8 + 2 * t;  // this thread's two reduction points: p0 (k4 step 0) and p0 + 1 (k4 step 1)
      const double* b0 = Bt + p0 * SB + 2 * g;
      // k4 step 0 with (k0a, k0b); meanwhile the Gram values of step 1
      const double k1a = gram_value(s01, p0 + 1 < kc);
      const double k1b = gram_value(s11, p0 + 1 < kc);
      mma_step(b0, k0a, k0b);
      // k4 step 1 with (k1a, k1b); meanwhile exponents + step-0 Gram values of the next group (possibly next chunk)
      if (grp + 1 < ngroups) {
        exponent_tile(Pt, grp + 1, s00, s01, s10, s11);
        k0a = gram_value(s00, p0 + 8 < kc);
        k0b = gram_value(s10, p0 + 8 < kc);
      } else if (c + 1 < nchunks) {
        const int nstage = (c + 1) % STAGES;
        mbar_wait(&full[nstage], ((uint32_t)((c + 1) / STAGES)) & 1u);
        const int64_t nk0 = k0 + BK;
        const int nkc = (int)((end - nk0 < BK) ? (end - nk0) : BK);
        exponent_tile(sP + nstage * BK * sp, 0, s00, s01, s10, s11);
        k0a = gram_value(s00, 2 * t < nkc);
        k0b = gram_value(s10, 2 * t < nkc);
      }
      mma_step(b0 + SB, k1a, k1b);
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&empty[stage]);
      // refill the stage consumed in iteration c-1 with chunk c + STAGES - 1: every warp has long left that stage, so the
      // wait does not stall, and the copies have one full chunk of compute to land
      const int cn = c + STAGES - 1;
      if (cn < nchunks) {
        mbar_wait(&empty[cn % STAGES], (((uint32_t)(cn / STAGES)) & 1u) ^ 1u);
        issue(cn);
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------
  // Column map: thread (g,t) holds, for column pair pr, the 4 consecutive columns j0 + 16 pr + 4 t + {0,1,2,3} as
  // acc[h][2pr][0], acc[h][2pr+1][0], acc[h][2pr][1], acc[h][2pr+1][1]; rows row0 + 16 warp + g + 8 h.
  if (BACKWARD) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = row0 + warp * 16 + g + 8 * h;
      if (r >= p.n_rows) continue;
      double* orow = p.out + ((int64_t)split * p.n_rows + r) * p.ldo;
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {
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

  if (p.epilogue == PLS_EPI_COST) {
    // per-column sum over this tile's rows of c(y_n, F[n][j]) -> out[rt][j]
    double ysel[2];
    bool rvalid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = row0 + warp * 16 + g + 8 * h;
      rvalid[h] = r < p.n_rows;
      ysel[h] = rvalid[h] ? p.y[r] : 0.0;
    }
    __syncthreads();  // every warp is done with the staging buffers; reuse stage 0 as reduction scratch
    double* sred = sB;  // [8 warps][BJ]
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int nt = 2 * pr + (e & 1);
        const int ce = e >> 1;
        double v = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (rvalid[h]) v += cost_value_call(p.cost, ysel[h], acc[h][nt][ce]);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) sred[warp * BJ + 16 * pr + 4 * t + e] = v;
      }
    }
    __syncthreads();
    if (tid < BJ && j0 + tid < p.j) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NTHREADS / 32; ++w) v += sred[w * BJ + tid];
      p.out[rt * p.ldo + j0 + tid] = v;
    }
    return;
  }

  const bool dcost = (p.epilogue == PLS_EPI_COST_DERIVATIVE);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t r = row0 + warp * 16 + g + 8 * h;
    if (r >= p.n_rows) continue;
    const double yv = dcost ? p.y[r] : 0.0;
    double* orow = p.out + r * p.ldo;
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
      const int64_t col = j0 + 16 * pr + 4 * t;
      double v[4] = {acc[h][2 * pr][0], acc[h][2 * pr + 1][0], acc[h][2 * pr][1], acc[h][2 * pr + 1][1]};
      if (dcost) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = cost_derivative_call(p.cost, yv, v[e]);
      }
      if (col + 3 < p.j) {
        double2* dst = reinterpret_cast<double2*>(orow + col);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < p.j) orow[col + e] = v[e];
      }
    }
  }
}

template <int NKD, bool BACKWARD>
cudaError_t launch_one(const pls_ctx* ctx, const GenGemmParams& p, int64_t grid, cudaStream_t stream) {
  const size_t smem = gen_gemm_smem_bytes(p.sp);
  if ((int64_t)smem > ctx->max_smem_optin) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(gen_gemm_kernel<NKD, BACKWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gen_gemm_kernel<NKD, BACKWARD><<<(unsigned)grid, NTHREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

}  // namespace pls
