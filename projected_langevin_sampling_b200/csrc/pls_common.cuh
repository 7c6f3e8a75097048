// Shared device helpers for the PLS sm_100a kernels: FP64 tensor-core MMA (DMMA.8x8x4, the only FP64 tensor shape
// sm_100a has), mbarrier + bulk-copy (TMA engine, UBLKCP) staging, tiling constants.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pls_b200.h"

namespace pls {

// ---- tiling of the generated-operand GEMM (pls_gen_gemm.cu) -----------------------------------------------------
constexpr int BR = 128;        // output rows per CTA: 8 warps x 16 rows (two m8 DMMA row tiles per warp)
constexpr int BJ = 128;        // output columns (particles) per CTA: 16 n8 DMMA column tiles per warp
constexpr int BK = 32;         // reduction points per pipeline stage (4 groups of 8)
constexpr int STAGES = 3;      // bulk-copy pipeline depth
constexpr int SB = BJ + 2;     // smem row stride (doubles) of the streamed tile: conflict-free LDS.128 fragments
constexpr int NTHREADS = 256;  // 8 warps, 1 CTA / SM (64 fp64 accumulators per thread)
constexpr int MAX_NKD = 7;     // D + 2 <= 28

__host__ __device__ inline int point_stride(int d) {
  // smallest SP >= d + 2 with SP % 8 == 4: rows of SP doubles make the B-fragment reads of the point tile
  // (lane (g,t) reads row g, column t + 4*kd) bank-conflict free and every row a multiple of 16 bytes.
  int need = d + 2;
  int sp = 4;
  while (sp < need) sp += 8;
  return sp;
}
__host__ __device__ inline int point_ksteps(int d) { return (d + 2 + 3) / 4; }

// ---- DMMA ------------------------------------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8), fp64.  Fragment layout (lane = 4*g + t): a = A[g][t], b = B[t][g],
// c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// exp(x) for the Gram exponent.  CUDA's double exp is <= 1 ulp; kept behind one name so the table-driven variant can be
// swapped in (see DESIGN.md "K generation cost").
__device__ __forceinline__ double gram_exp(double x) { return exp(x); }

}  // namespace pls
