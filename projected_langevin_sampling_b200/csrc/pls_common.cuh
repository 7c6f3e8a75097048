// Shared device helpers for the PLS sm_100a kernels: FP64 tensor-core MMA (DMMA.8x8x4, the only FP64 tensor shape
// sm_100a has), mbarrier + bulk-copy (TMA engine, UBLKCP) staging, tiling constants.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pls_b200.h"

namespace pls {

// ---- tiling of the generated-operand GEMM (pls_gen_gemm.cu) -----------------------------------------------------
// CTA tile: 8 warps, each RT m8 row tiles x 32/RT n8 column tiles (64 fp64 accumulators per thread either way).
//   RT = 1: 64 rows x 256 particles (default: a generated Gram element is reused across 256 columns)
//   RT = 2: 128 rows x 128 particles (particle slices of <= 128 columns)
__host__ __device__ constexpr int tile_rows(int rt) { return 64 * rt; }
__host__ __device__ constexpr int tile_cols(int rt) { return 256 / rt; }
constexpr int BK = 32;         // reduction points per pipeline stage (4 groups of 8)
constexpr int STAGES = 3;      // bulk-copy pipeline depth
constexpr int NTHREADS = 256;  // 8 warps, 1 CTA / SM
constexpr int MAX_NKD = 7;     // ceil(D / 4) with D <= 26

__host__ __device__ inline int point_stride(int d) {
  // smallest SP >= d + 2 with SP % 8 == 4: rows of SP doubles make the B-fragment reads of the point tile
  // (lane (g,t) reads row g, column t + 4*kd) bank-conflict free and every row a multiple of 16 bytes.
  int need = d + 2;
  int sp = 4;
  while (sp < need) sp += 8;
  return sp;
}
__host__ __device__ inline int point_ksteps(int d) { return (d + 3) / 4; }  // exponent DMMAs per 8 x 8 tile: coordinates only

// ---- DMMA ------------------------------------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8), fp64.  Fragment layout (lane = 4*g + t): a = A[g][t], b = B[t][g],
// c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier / bulk copy ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy on the TMA engine (SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- exp -------------------------------------------------------------------------------------------------------
// exp(x) of the Gram exponent outside the hot loop (dense Gram, selector): CUDA's double exp, <= 1 ulp.
__device__ __forceinline__ double gram_exp(double x) { return exp(x); }

// 2^(j/64), j = 0..63, correctly rounded.  Copied to shared memory by the hot kernel.
__device__ __constant__ double kExp2Table[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// exp(x) inside the hot loop.  FP64 DFMA shares the 64-FMA/clk/SM pipe with DMMA, so every FP64 op here is taken from
// the GEMM: table-driven reduction x = (64 e + j) ln2/64 + r, |r| <= ln2/128, degree-5 polynomial -> 10 FP64 ops and a
// dependent chain of ~8 (CUDA's exp: ~20 ops, chain ~20), branch-free so it can be interleaved with the DMMA stream.
// Max error ~1 ulp (table entry 0.5 ulp + polynomial/rounding), checked against exp() in tests/test_gpu_kernels.py.
// Inputs below -700 are clamped (result ~1e-304 instead of an underflowed 0: irrelevant at any Gram scale).
__device__ __forceinline__ double gram_exp_fast(double x, const double* __restrict__ tbl) {
  // clamp on the ALU (compare of the high word), not on the FP64 pipe
  if ((unsigned)__double2hiint(x) >= 0xC085E000u) x = -700.0;
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);  // 64/ln2, 1.5 * 2^52: integer part lands in the low word
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -0.010830424695996044, x);   // ln2/64, high 34 bits (exact product)
  r = fma(kf, -2.5310172166650877e-13, r);        // ln2/64, low part
  double q = fma(r, 8.3333333333333333e-3, 4.1666666666666664e-2);
  q = fma(q, r, 1.6666666666666666e-1);
  q = fma(q, r, 0.5);
  const double r2 = r * r;
  const double p = fma(q, r2, r);  // expm1(r)
  const double tj = tbl[k & 63];
  const double res = fma(tj, p, tj);
  return __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
}

}  // namespace pls
