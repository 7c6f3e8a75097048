// The hot kernel of the PLS Langevin step: a GEMM whose A operand is GENERATED, never stored (default), or streamed from a
// caller-kept Gram (opt-in).
//
//     C[r][j] (+)= sum_k  kappa(row_r, red_k) * B[k][j]
//
//   forward role   rows = training points x_n, reduction = inducing points z_m, B = W = V~ P  (M x J)
//                  -> F = k(X,Z) W, then the cost epilogue                           (orthonormal.py:98-108 + costs/*.py)
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
//   * B (W or Dc rows) is staged by ONE tensor-map TMA instruction per 32-row stage (cp.async.bulk.tensor.3d ->
//     UTMALDG): the matrix is described as [column block of 16][row][16 doubles] and lands in shared memory as
//     BJ/16 blocks of 32 rows x 128 bytes with the 128-byte swizzle, which makes the LDS.128 B-fragment reads
//     bank-conflict free without padding (lane (g,t) reads 16-byte chunk g ^ (row & 7) of row 2t [+1]).  The
//     reduction-point rows follow with one plain bulk copy.  3-stage full-barrier pipeline; a stage is refilled by the
//     LAST warp to release it (atomic counter), so no warp ever waits to issue a copy.
//   * the loop over 8-point groups is flat and branch-free (the Gram source is a template parameter) and the Gram values of
//     a whole 32-point stage are produced one stage ahead, in two register sets that swap roles (never copied), so ptxas
//     interleaves the 8 independent exponent chains of the NEXT stage with the DMMAs of the current one.
//   * NS = 2 (default when the particle slice is an even number of 256-column tiles and the launch is large enough for the wider
//     tile to pay -- choose_tile_ns, pls_internal.h): the CTA tile is 64 x 512.  The 64
//     accumulators of the SECOND 256-column half are parked in tensor memory (tcgen05.st / tcgen05.ld, 2 x 128 columns per
//     warp, ping-pong) and swapped with the register set once per 32-point chunk; the streamed blocks come in the order
//     A0 B0 | B1 A1 | A2 B2 ... so a chunk costs ONE swap (137 clk, tools/tmem_swap_microbench.cu) and every generated Gram
//     value feeds 512 columns: the exponent + exp share of the FP64 pipe halves (7.4 % -> 3.7 % at D = 8).  Each column is
//     still accumulated over the chunks in order, so the results are bit-identical to NS = 1.  The reduction points of chunk
//     c + 1 travel with the SECOND block of chunk c (its Gram values are formed during that block), so no warp waits on a block
//     ahead of the one it multiplies; the stage refill (one lane, while its warp waits) is free of divisions.
//   * CL = 2 (backward role by default, choose_cluster in pls_internal.h): a cluster of two CTAs takes two adjacent 512-column tiles
//     of the same rows and reduction range; the Gram values of chunk c are formed by the CTA of rank c mod 2 only and mailed lane to
//     lane to the same warp of its peer through distributed shared memory (st.async + mbarrier transaction bytes; a remote arrive
//     hands the mailbox back), so one generated value feeds 1024 columns.  Same bits; backward +0.6 %, the persistent forward -0.8 %
//     (its warps lose the free drift of their epilogues), hence backward only.
//   * optional KSRC_CACHED variant (pls_*_cached_f64): the caller keeps k(X, Z) in HBM and the kernels LOAD their fragment
//     values (L2::evict_last, one stage ahead) instead of generating them -- the FP64 pipe then runs nothing but the
//     contraction.  The default path generates: nothing N x M is ever in memory.
#pragma once
#include <cuda.h>

#include "pls_cost.cuh"
#include "pls_internal.h"
#include "pls_tmem.cuh"

namespace pls {

namespace {

constexpr int NWARPS = NTHREADS / 32;
constexpr int BLOCK_BYTES = BK * 128;  // one 16-column block of a stage: 32 rows x 128 bytes

// Four cost derivatives per call (one accumulator column group), arguments and results in registers.  Called, not inlined,
// from the register epilogue: 16 calls per thread and tile keep the kernel small while the four evaluations inside give
// the transcendental functors instruction-level parallelism.  `c` points to a shared-memory copy of the cost.
struct D4 {
  double a, b, c, d;
};
// One (cost, link, closed-form) combination with its identifiers as compile-time constants: the functor's switches fold
// away and the four evaluations become straight-line code whose dependent FP64 chains (divisions, exp) interleave.  Left to
// the run-time switches, each evaluation is a chain of basic blocks and the four run one after the other, every dependent
// op queueing behind the other warp's DMMAs (measured: ~1200 clk per call).
template <int CID, int LID, int CF>
static __device__ __forceinline__ D4 cost_derivative4_as(const pls_cost& c, double y, const D4& f, const double* exp_table) {
  pls_cost cc = c;
  cc.cost_id = CID;
  cc.link_id = LID;
  cc.closed_form = CF;
  const FlatMath m{exp_table};  // branch-free division and exp: the four chains interleave
  D4 r;
  r.a = cost_derivative(cc, y, f.a, m);
  r.b = cost_derivative(cc, y, f.b, m);
  r.c = cost_derivative(cc, y, f.c, m);
  r.d = cost_derivative(cc, y, f.d, m);
  return r;
}
// `c->reserved` carries the case index cost_id * 8 + link_id * 2 + closed_form (set by the launcher): the two closed forms the
// reference's experiments run at scale are tested for FIRST, before the cost struct is copied and the jump table is taken --
// all inside this routine, so the kernels that call it are unchanged (a Poisson branch in the kernel's own epilogue cost the
// Gaussian forward 0.9 %).
constexpr int kCasePoissonSquareCF = PLS_COST_POISSON * 8 + PLS_LINK_SQUARE * 2 + 1;
constexpr int kCaseBernoulliSigmoidCF = PLS_COST_BERNOULLI * 8 + PLS_LINK_SIGMOID * 2 + 1;
static __device__ __noinline__ D4 cost_derivative4(const pls_cost* c, const double* exp_table, double y, D4 f) {
  const int fast_case = c->reserved;
  if (fast_case == kCasePoissonSquareCF) {  // -2 y / F + 2 F (poisson.py:68-82): no parameter of the struct is needed
    const FlatMath m{exp_table};
    D4 r;
    r.a = -2.0 * m.div(y, f.a) + 2.0 * f.a;
    r.b = -2.0 * m.div(y, f.b) + 2.0 * f.b;
    r.c = -2.0 * m.div(y, f.c) + 2.0 * f.c;
    r.d = -2.0 * m.div(y, f.d) + 2.0 * f.d;
    return r;
  }
  if (fast_case == kCaseBernoulliSigmoidCF) {  // only the link's jitter is read
    pls_cost cb;
    cb.link_jitter = c->link_jitter;
    return cost_derivative4_as<PLS_COST_BERNOULLI, PLS_LINK_SIGMOID, 1>(cb, y, f, exp_table);
  }
  const pls_cost cc = *c;
#define PLS_CASE(CID, LID)                                                   \
  case (CID * 8 + LID * 2 + 0): return cost_derivative4_as<CID, LID, 0>(cc, y, f, exp_table); \
  case (CID * 8 + LID * 2 + 1): return cost_derivative4_as<CID, LID, 1>(cc, y, f, exp_table);
#define PLS_CASES(CID) PLS_CASE(CID, 0) PLS_CASE(CID, 1) PLS_CASE(CID, 2) PLS_CASE(CID, 3)
  switch (cc.cost_id * 8 + cc.link_id * 2 + (cc.closed_form != 0)) {
    PLS_CASES(0) PLS_CASES(1) PLS_CASES(2) PLS_CASES(3) PLS_CASES(4)
  }
#undef PLS_CASES
#undef PLS_CASE
  return f;
}

// Four cost VALUES per call (the energy's cost sums), specialised and branch-free like the derivatives.
template <int CID, int LID>
static __device__ __forceinline__ D4 cost_value4_as(const pls_cost& c, double y, const D4& f, const double* exp_table) {
  pls_cost cc = c;
  cc.cost_id = CID;
  cc.link_id = LID;
  const FlatMath m{exp_table};
  D4 r;
  r.a = cost_value(cc, y, f.a, m);
  r.b = cost_value(cc, y, f.b, m);
  r.c = cost_value(cc, y, f.c, m);
  r.d = cost_value(cc, y, f.d, m);
  return r;
}
static __device__ __noinline__ D4 cost_value4(const pls_cost* c, const double* exp_table, double y, D4 f) {
  const pls_cost cc = *c;
#define PLS_CASE(CID, LID) \
  case (CID * 4 + LID): return cost_value4_as<CID, LID>(cc, y, f, exp_table);
#define PLS_CASES(CID) PLS_CASE(CID, 0) PLS_CASE(CID, 1) PLS_CASE(CID, 2) PLS_CASE(CID, 3)
  switch (cc.cost_id * 4 + cc.link_id) {
    PLS_CASES(0) PLS_CASES(1) PLS_CASES(2) PLS_CASES(3) PLS_CASES(4)
  }
#undef PLS_CASES
#undef PLS_CASE
  return f;
}

// Four cost VALUES and four DERIVATIVES per call, for the fused epilogue of pls_forward_step_f64 (the training loop's energy
// comes from the same F tile as the gradient): one specialised body evaluates both, so whatever the two share -- the link
// (sigmoid: one exp and one division instead of two of each), the residuals, the Student-t / mixture denominators -- is
// computed once, and the eight dependent FP64 chains interleave instead of running as two calls of four.
struct D8 {
  D4 d, c;
};
template <int CID, int LID, int CF>
static __device__ __forceinline__ D8 cost_both4_as(const pls_cost& c, double y, const D4& f, const double* exp_table) {
  pls_cost cc = c;
  cc.cost_id = CID;
  cc.link_id = LID;
  cc.closed_form = CF;
  const FlatMath m{exp_table};
  D8 r;
  r.d.a = cost_derivative(cc, y, f.a, m);
  r.d.b = cost_derivative(cc, y, f.b, m);
  r.d.c = cost_derivative(cc, y, f.c, m);
  r.d.d = cost_derivative(cc, y, f.d, m);
  r.c.a = cost_value(cc, y, f.a, m);
  r.c.b = cost_value(cc, y, f.b, m);
  r.c.c = cost_value(cc, y, f.c, m);
  r.c.d = cost_value(cc, y, f.d, m);
  return r;
}
static __device__ __noinline__ D8 cost_both4(const pls_cost* c, const double* exp_table, double y, D4 f) {
  const int fast_case = c->reserved;  // as in cost_derivative4: the two closed forms run at scale skip the struct copy and the jump table
  if (fast_case == kCasePoissonSquareCF) {
    pls_cost cb;  // (no parameter is read)
    cb.link_jitter = 0.0;
    return cost_both4_as<PLS_COST_POISSON, PLS_LINK_SQUARE, 1>(cb, y, f, exp_table);
  }
  if (fast_case == kCaseBernoulliSigmoidCF) {
    pls_cost cb;
    cb.link_jitter = c->link_jitter;
    return cost_both4_as<PLS_COST_BERNOULLI, PLS_LINK_SIGMOID, 1>(cb, y, f, exp_table);
  }
  const pls_cost cc = *c;
#define PLS_CASE(CID, LID)                                                   \
  case (CID * 8 + LID * 2 + 0): return cost_both4_as<CID, LID, 0>(cc, y, f, exp_table); \
  case (CID * 8 + LID * 2 + 1): return cost_both4_as<CID, LID, 1>(cc, y, f, exp_table);
#define PLS_CASES(CID) PLS_CASE(CID, 0) PLS_CASE(CID, 1) PLS_CASE(CID, 2) PLS_CASE(CID, 3)
  switch (cc.cost_id * 8 + cc.link_id * 2 + (cc.closed_form != 0)) {
    PLS_CASES(0) PLS_CASES(1) PLS_CASES(2) PLS_CASES(3) PLS_CASES(4)
  }
#undef PLS_CASES
#undef PLS_CASE
  return D8{f, f};
}

template <int RT>
struct Tile {
  static constexpr int NT = 32 / RT;          // n8 column tiles per warp
  static constexpr int BR = tile_rows(RT);    // rows per CTA
  static constexpr int BJ = tile_cols(RT);    // columns per CTA
  static constexpr int NPR = NT / 2;          // 16-column blocks = column pairs (one LDS.128 each)
  static constexpr int STAGE_BYTES = NPR * BLOCK_BYTES;
  static constexpr int SF = BJ + 2;           // row stride (doubles) of the forward epilogue's staging tile
};

constexpr int SMEM_HEADER = 4096;  // barriers + counters + exp table [0, 1024) | y of the rows of two tiles [1024, 3072)

// shared memory: header | STAGES x stage (1024-aligned) | STAGES x BK x sp points
template <int RT>
__host__ __device__ inline size_t gen_gemm_smem_bytes(int sp) {
  // the forward epilogue stages the output tile through ONE pipeline stage (32 rows at a time): no extra staging memory
  const size_t pipeline = (size_t)STAGES * Tile<RT>::STAGE_BYTES + sizeof(double) * (size_t)(STAGES * BK * sp);
  const size_t cost_scratch = sizeof(double) * 2 * NTHREADS;  // [PHASES][BJ]
  return 1024 /* alignment slack */ + SMEM_HEADER + pipeline + cost_scratch;
}
// with room for the per-warp cost sums [NWARPS][BJ] of the register cost-sum epilogue (shares the scratch region)
template <int RT>
__host__ __device__ inline size_t gen_gemm_smem_bytes_wbuf(int sp) {
  const size_t pipeline = (size_t)STAGES * Tile<RT>::STAGE_BYTES + sizeof(double) * (size_t)(STAGES * BK * sp);
  return 1024 + SMEM_HEADER + pipeline + sizeof(double) * (size_t)(NWARPS * Tile<RT>::BJ);
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
constexpr int KSRC_LINEAR = 0, KSRC_RBF = 1, KSRC_CACHED = 2;

// ---- thread-block cluster helpers (CL = 2: the two CTAs of a cluster share the generation of the Gram values) -------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, unsigned rank) {  // shared::cta address -> shared::cluster address in `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// two doubles into the peer's shared memory; completion is signalled as 16 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_v2(uint32_t remote_addr, double a, double b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "d"(a), "d"(b),
               "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");  // (default semantics, as CUTLASS's ClusterBarrier::arrive)
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope: data written by the peer CTA
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CWAIT_DONE;\n"
      "bra CWAIT_LOOP;\n"
      "CWAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// L2 eviction policies.  The persistent forward CTAs that share a Gram row slab (the 16 column tiles of a row tile) drift
// apart by several tile times while 8.6 GB of Dc stream through L2 per launch: the cached Gram values are loaded evict_last and
// Dc is stored evict_first so that the slab survives until its last reader (ncu: DRAM reads of a forward launch 9.9 -> see
// profiles/ncu_gen_gemm_cached_r01.txt).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// read-only loads of cached Gram values (each value is used once per tile: keep it out of L1, keep it in L2)
__device__ __forceinline__ double2 ldg_stream_v2(const double* ptr, uint64_t pol) {
  double2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ double ldg_stream(const double* ptr, uint64_t pol) {
  double v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol));
  return v;
}

__device__ __forceinline__ unsigned atom_add_acq_rel_shared(unsigned* addr, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}
// Stage-release counter.  Relaxed on purpose: a warp reaches this only after the DMMAs that consumed its LDS data of the
// stage have issued (in-order issue: the loads have returned), so when the last warp sees the count the stage is no longer
// being read and the tensor-map copy may overwrite it; acq_rel would put a MEMBAR in front of every release.
__device__ __forceinline__ unsigned atom_add_shared(unsigned* addr, unsigned v) {
  unsigned old;
  asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}

// KSRC: where the Gram values come from -- generated from the augmented points (KSRC_RBF: exponent DMMAs + exp, KSRC_LINEAR:
// exponent DMMAs only) or read from a Gram matrix the caller keeps in HBM (KSRC_CACHED: p.gram, rows = training points,
// columns = inducing points; no exponent work at all on the FP64 pipe, one streaming load per value instead).
// EPI: the forward epilogue (PLS_EPI_*), a template parameter so that every kernel carries only its own epilogue's code and
// register budget; the backward role is instantiated with EPI = -1.
// NS: 256-column accumulator sets per CTA tile (1, or 2 with the second set parked in tensor memory; RT = 1 and the
// register epilogues only).
// CL: CTAs per cluster (1, or 2 with NS = 2: the two CTAs of a cluster work on adjacent 512-column tiles of the same rows and
// reduction range; the Gram values of chunk c are formed by the CTA of rank c mod 2 only and mailed to its peer through distributed
// shared memory, so every generated value feeds 1024 columns).
template <int NKD, bool BACKWARD, int KSRC, int RT, int EPI, int NS, int CL = 1>
__global__ void __launch_bounds__(NTHREADS, 1)
    gen_gemm_kernel(const GenGemmParams p, const __grid_constant__ CUtensorMap tm3, const __grid_constant__ CUtensorMap tm2) {
  using T = Tile<RT>;
  constexpr int NT = T::NT, BR = T::BR, BJ = T::BJ, NPR = T::NPR, STAGE_BYTES = T::STAGE_BYTES;
  constexpr int BJT = BJ * NS;  // columns of the CTA tile
  static_assert(NS == 1 || (NS == 2 && RT == 1 && KSRC != KSRC_CACHED), "accumulator parking: 64 x 512 tiles of generated Gram values");
  static_assert(CL == 1 || (CL == 2 && NS == 2), "Gram sharing across a cluster builds on the 64 x 512 tile");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // the swizzle needs 1024-byte stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                      // [STAGES]
  unsigned* released = reinterpret_cast<unsigned*>(smem_raw + 64);             // [STAGES] warps done with the stage
  double* sExp = reinterpret_cast<double*>(smem_raw + 128);                    // 2^(j/64)
  pls_cost* sCost = reinterpret_cast<pls_cost*>(smem_raw + 704);              // the cost, for the called functor
  double* sYall = reinterpret_cast<double*>(smem_raw + 1024);                  // [2][128] targets of the tile's rows (forward)
  unsigned char* sB = smem_raw + SMEM_HEADER;                                  // [STAGES][NPR][BK rows][128 bytes], swizzled
  double* sP = reinterpret_cast<double*>(sB + STAGES * STAGE_BYTES);           // [STAGES][BK][sp]
  double* sC = sP + STAGES * BK * p.sp;                                        // [PHASES][BJ] cost-sum scratch (staged epilogue) /
                                                                               // [NWARPS][BJ] per-warp cost sums (register epilogue)
  unsigned* tile_done = released + 4;                                          // [2] warps that published their sums, by tile parity
  volatile unsigned* combined = released + 6;                                  // tiles whose sums have been combined
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem_raw + 96);          // NS = 2: base address of the tensor-memory allocation
  uint64_t* kfull = reinterpret_cast<uint64_t*>(smem_raw + 3072);              // CL = 2, [NWARPS]: the peer's Gram values have arrived in sK
  uint64_t* kfree = kfull + NWARPS;                                            // CL = 2, [NWARPS]: the peer has read the values I mailed
  double* sK = sC + 2 * NTHREADS;                                              // CL = 2, [NWARPS][32 lanes][8]: mailbox for the peer's Gram values

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int g = lane >> 2;  // DMMA group id
  const int t = lane & 3;   // DMMA thread-in-group
  const int sp = p.sp;

  // ---- tiles of this CTA ------------------------------------------------------------------------------------------------
  // backward role: ONE tile = (inducing-row tile, column tile, split of the training rows), row tiles fastest so that the CTAs
  //   streaming the same Dc slab run together and share it in L2;
  // forward role: PERSISTENT -- the grid is one CTA per SM and CTA b takes tiles b, b + grid, b + 2 grid, ... (column tiles
  //   fastest: the CTAs in flight share training rows and the whole W).  The copy pipeline runs across tile boundaries (the
  //   first stages of the next tile land during the epilogue of the current one) and nothing is re-initialised per tile.
  const int64_t n_row_tiles = (p.n_rows + BR - 1) / BR;
  const int64_t n_col_tiles = (p.j + BJT * CL - 1) / (BJT * CL);  // tiles of a CLUSTER: CL adjacent CTA tiles
  const int64_t total_tiles = n_row_tiles * n_col_tiles * (BACKWARD ? p.splits : 1);
  const unsigned crank = (CL == 2) ? cluster_ctarank() : 0u;      // this CTA's column position inside its cluster's tile
  const unsigned cid = blockIdx.x / CL, ncl = gridDim.x / CL;     // cluster index / clusters in the grid
  const int my_tiles = BACKWARD ? 1 : (int)((total_tiles - (int64_t)cid + ncl - 1) / ncl);

  // reduction range: the forward role reduces over all inducing points in every tile; the backward role over its split
  int64_t begin = 0, end = p.red_total;
  int64_t bw_rt = 0, bw_ct = 0;
  int split = 0;
  if (BACKWARD) {
    int64_t bid = cid;
    bw_rt = bid % n_row_tiles;
    bid /= n_row_tiles;
    bw_ct = bid % n_col_tiles;
    split = (int)(bid / n_col_tiles);
    const int64_t total_chunks = (p.red_total + BK - 1) / BK;
    const int64_t per = (total_chunks + p.splits - 1) / p.splits;
    begin = (int64_t)split * per * BK;
    end = begin + per * BK;
    if (end > p.red_total) end = p.red_total;
    if (begin > end) begin = end;
  }
  const int red_len = (int)(end - begin);
  const int nchunks = (red_len + BK - 1) / BK;         // 32-point chunks per tile
  const int nblocks = nchunks * NS;                    // streamed blocks (pipeline stages) per tile: one per chunk and column half
  const int total_gc = my_tiles * nblocks;             // blocks this CTA streams, over all its tiles (the launcher checks the range)
  // (32-bit unsigned arithmetic: this runs on ONE lane at every stage refill while the rest of its warp waits -- the 64-bit
  // division it used to contain cost the slowest warp of the CTA several hundred cycles per block; the launcher checks the range)
  const unsigned nct32 = (unsigned)n_col_tiles;
  auto tile_coords = [&](int ti, int64_t& rt, int64_t& ct) {
    if (BACKWARD) {
      rt = bw_rt;
      ct = bw_ct;
    } else {
      const unsigned tile = cid + (unsigned)ti * ncl;
      const unsigned q = tile / nct32;
      ct = (int64_t)(tile - q * nct32);
      rt = (int64_t)q;
    }
  };

  // ---- one-time setup -------------------------------------------------------------------------------------------
  // point rows past the end of the reduction are never copied: keep them finite
  for (int i = tid; i < STAGES * BK * sp; i += NTHREADS) sP[i] = 0.0;
  if (tid < 64) sExp[tid] = kExp2Table[tid];
  if (tid == 0) {
    *sCost = p.cost;
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      released[s] = 0;
    }
    released[4] = released[5] = released[6] = 0;
    if (CL == 2)
      for (int w = 0; w < NWARPS; ++w) {
        mbar_init(&kfull[w], 1);
        mbar_init(&kfree[w], 1);
      }
    fence_mbar_init();
  }
  if (NS == 2 && warp == 0) tmem_alloc_512(tmem_base_s);  // all 512 columns: 2 warps per lane quarter x 2 slots x 128 columns
  fence_proxy_async();  // generic-proxy zero fill ordered before the async-proxy bulk copies
  if (NS == 2) tmem_fence_before_sync();
  __syncthreads();
  int nrecv = 0, nsent = 0;  // CL = 2: chunks whose Gram values this warp has received from / mailed to its peer
  if (CL == 2) {
    if (lane == 0) mbar_expect_tx(&kfull[warp], 32u * 64u);  // armed for the first delivery: 8 doubles from each of the peer warp's lanes
    cluster_sync_all();  // both CTAs' barriers exist and are armed before anything is mailed
  }
  uint32_t tslot = 0;  // this warp's two parking slots: tslot, tslot + 128 (its own 32 lanes, 256 of the 512 columns)
  if (NS == 2) {
    tmem_fence_after_sync();
    tslot = *tmem_base_s + ((32u * (unsigned)(warp & 3)) << 16) + 256u * (unsigned)(warp >> 2);
  }

  // One thread fills a stage: one tensor-map copy for the streamed tile (all BJ/16 column blocks; rows or columns outside
  // the matrix arrive as zeros), one bulk copy for the reduction-point rows.  The 3-D map cannot describe a partial last
  // column block, so a tile that contains one uses the 2-D map block by block.  gc = chunk index over all tiles of the CTA.
  // NS = 2: block b of a tile is chunk b / 2, and the halves alternate A B | B A | A B ... so that a chunk boundary never
  // needs a swap of the accumulator sets.  The Gram values of chunk c + 1 are formed during the SECOND block of chunk c, so
  // that block carries the reduction points of chunk c + 1 (read from the stage in use, before it is released: no warp ever
  // waits on a block ahead of the one it multiplies); the tile's first block carries the points of chunk 0 for the prologue.
  // issue_block: block b of a tile in column position ct_i into `stage`.
  auto issue_block = [&](int ct_i, int b, int stage) {
    const int c = (NS == 2) ? (b >> 1) : b;
    const int half = (NS == 2) ? (((c & 1) != 0) != ((b & 1) != 0) ? 1 : 0) : 0;
    const int cp = (NS == 2) ? ((b == 0) ? 0 : ((b & 1) ? c + 1 : nchunks)) : c;  // chunk whose points travel with this block
    const bool with_points = KSRC != KSRC_CACHED && cp < nchunks;
    const int64_t j0 = ((int64_t)ct_i * CL + crank) * BJT + half * BJ;
    const bool use3d = p.tma3d && (j0 + BJ <= p.full_blocks * 16 || p.full_blocks * 16 == p.ldb);
    const int64_t k0 = begin + (int64_t)c * BK;
    const int64_t kp0 = begin + (int64_t)cp * BK;
    const int kc = (int)((end - kp0 < BK) ? (end - kp0) : BK);  // points copied
    uint64_t* bar = &full[stage];
    mbar_expect_tx(bar, (uint32_t)(STAGE_BYTES + (with_points ? kc * sp * 8 : 0)));
    unsigned char* dst = sB + stage * STAGE_BYTES;
    if (use3d) {
      tma_load_3d(dst, &tm3, 0, (int)k0, (int)(j0 >> 4), bar);
    } else {
#pragma unroll 1
      for (int blk = 0; blk < NPR; ++blk) tma_load_2d(dst + blk * BLOCK_BYTES, &tm2, (int)j0 + 16 * blk, (int)k0, bar);
    }
    if (with_points) bulk_g2s(sP + stage * BK * sp, p.red_aug + kp0 * sp, (uint32_t)(kc * sp * 8), bar);  // (sp = 0 when cached)
  };
  // general form: block gc of this CTA's stream (two divisions: used for the first STAGES blocks and for tiles shorter than the pipeline)
  auto issue = [&](int gc) {
    const int ti = (int)((unsigned)gc / (unsigned)nblocks);
    int64_t rt, ct;
    tile_coords(ti, rt, ct);
    issue_block((int)ct, gc - ti * nblocks, (int)((unsigned)gc % (unsigned)STAGES));
  };
  // refill of the stage a warp has just released with the block STAGES further on: same stage; the block belongs to the tile the
  // warp is in (column position ct_cur) or to this CTA's next tile, gridDim.x tiles further on -- no division on this path, which
  // one lane runs while the rest of its warp waits
  const int gstep = BACKWARD ? 0 : (int)(ncl % nct32);
  auto issue_from = [&](int gc_cur, int ct_cur, int b_cur, int stage) {
    int b = b_cur + STAGES, ct_i = ct_cur;
    if (b >= nblocks) {
      b -= nblocks;
      ct_i += gstep;
      if (ct_i >= (int)nct32) ct_i -= (int)nct32;
      if (b >= nblocks) {
        issue(gc_cur + STAGES);
        return;
      }
    }
    issue_block(ct_i, b, stage);
  };
  if (tid == 0) {
    for (int gc = 0; gc < STAGES && gc < total_gc; ++gc) issue(gc);
  }

  // row-side exponent fragments (A operand of the S DMMA): row g of each of this warp's RT row tiles.  Coordinates only;
  // the c entry (column d of the augmented row) seeds the accumulator together with the point's c.
  double a2[RT][NKD];
  double crow[RT];
  auto load_rows = [&](int64_t row0, double (&a)[RT][NKD], double (&cr)[RT]) {
    if (KSRC == KSRC_CACHED) return;
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      const int64_t r = row0 + (warp * RT + h) * 8 + g;
      const bool rv = r < p.n_rows;
#pragma unroll
      for (int kd = 0; kd < NKD; ++kd) {
        const int dd = t + 4 * kd;
        a[h][kd] = (rv && dd < p.d) ? p.rows_aug[r * sp + dd] : 0.0;
      }
      cr[h] = rv ? p.rows_aug[r * sp + p.d] : 0.0;
    }
  };

  double acc[RT][NT][2];

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
    const double v = (KSRC == KSRC_RBF) ? gram_exp_fast(sv, sExp) : sv;
    return valid ? v : 0.0;
  };
  // one k4 step: brow = this lane's 16-byte chunk of its reduction row in column block 0
  auto mma_step = [&](const unsigned char* brow, const double (&ka)[RT]) {
#pragma unroll
    for (int pr = 0; pr < NPR; ++pr) {
      const double2 bv = *reinterpret_cast<const double2*>(brow + pr * BLOCK_BYTES);
#pragma unroll
      for (int h = 0; h < RT; ++h) {
        dmma(acc[h][2 * pr][0], acc[h][2 * pr][1], ka[h], bv.x);
        dmma(acc[h][2 * pr + 1][0], acc[h][2 * pr + 1][1], ka[h], bv.y);
      }
    }
  };
  // swizzled position of columns (2g, 2g+1) of rows 2t and 2t+1 within an 8-row group of a column block
  const int off0 = (2 * t) * 128 + ((g ^ (2 * t)) << 4);
  const int off1 = (2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4);

  // Gram values are generated one BLOCK of LA groups ahead of the DMMAs that consume them: the 2 RT LA exponent chains of a
  // block are independent, so they issue back to back at the pipe's rate instead of stalling the (in-order) warp on each
  // dependent FP64 op.  LA = 4 (a whole stage) for RT = 1; 2 for RT = 2 (register budget).
  constexpr int GROUPS = BK / 8;
  constexpr int LA = (RT == 1) ? GROUPS : GROUPS / 2;
  constexpr int NBLK = GROUPS / LA;
  // Gram values of groups [g0, g0 + LA) of the point tile Pt; `first` = this thread's first reduction point of group g0,
  // counted from `begin` (points past the end of the reduction give 0)
  const double* kbase[RT];  // cached Gram: this thread's row (forward) / column (backward) of the tile, set per tile
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  auto gram_block = [&](const double* Pt, int g0, int first, double (&k)[LA][2][RT]) {
    if (KSRC == KSRC_CACHED) {
      // points first + 8 q and first + 8 q + 1 (from `begin`) of each of this thread's rows: forward = two adjacent columns of a
      // Gram row (one 16-byte load), backward = the same column of two adjacent Gram rows.  The loads are issued a block ahead
      // of the DMMAs that consume them, like the generated values.
#pragma unroll
      for (int q = 0; q < LA; ++q) {
        const int64_t pt = begin + first + 8 * q;
#pragma unroll
        for (int h = 0; h < RT; ++h) {
          double2 v;
          if (BACKWARD) {
            v.x = ldg_stream(kbase[h] + pt * p.ldk, pol_keep);
            v.y = ldg_stream(kbase[h] + (pt + 1) * p.ldk, pol_keep);
          } else {
            v = ldg_stream_v2(kbase[h] + pt, pol_keep);
          }
          // no masking: a point past the end of the reduction meets a zero row of the streamed matrix (the tensor map
          // zero-fills rows outside it), and the cache's padding is finite by contract.  A select here would be the first use
          // of the load and stall the warp for the whole memory latency before the block's DMMAs (measured: 9 % of the samples).
          k[q][0][h] = v.x;
          k[q][1][h] = v.y;
        }
      }
      return;
    }
#pragma unroll
    for (int q = 0; q < LA; ++q) {
      double s[RT][2];
      exponent_tile(Pt, g0 + q, s);
#pragma unroll
      for (int h = 0; h < RT; ++h) {
        k[q][0][h] = gram_value(s[h][0], first + 8 * q < red_len);
        k[q][1][h] = gram_value(s[h][1], first + 8 * q + 1 < red_len);
      }
    }
  };

  // Forward epilogues.  PREDICTION and COST_DERIVATIVE are written straight from the accumulator registers with 256-bit
  // stores: no shared-memory staging and NO block-wide synchronisation, so the warps of the persistent CTA drift apart and
  // one warp's epilogue overlaps the others' tensor work.  The Gaussian / identity closed form (one multiply-subtract per
  // element) is inline, every other cost functor is called four values at a time.  The epilogues that produce cost sums
  // need the tile's rows together and are staged through shared memory.
  const bool deriv_direct = !BACKWARD && EPI == PLS_EPI_COST_DERIVATIVE;
  const bool gauss_cost = p.cost.cost_id == PLS_COST_GAUSSIAN && p.cost.link_id == PLS_LINK_IDENTITY && p.cost.closed_form != 0;
  const bool gauss_direct = deriv_direct && gauss_cost;
  // The cost-sum epilogues (COST, and the fused derivative + cost of pls_forward_step_f64) take the register path as well when
  // the launcher found room for the per-warp sums and a tile has at least STAGES + 1 stages (so that no warp can finish
  // tile i + 1 before the slowest warp has finished tile i -- the copy pipeline bounds the drift to STAGES stages): every
  // warp reduces its 8 rows per column with a shuffle reduce-scatter, publishes 256 sums, and the LAST warp of the tile adds
  // the 8 partials in warp order (fixed order: deterministic) and writes the row tile's sums.  No block barrier.
  const bool sums_direct = !BACKWARD && p.wbuf_ok && nblocks > STAGES &&
                           (EPI == PLS_EPI_COST || EPI == PLS_EPI_COST_DERIVATIVE_AND_COST);
  const bool store_direct = deriv_direct || (sums_direct && EPI == PLS_EPI_COST_DERIVATIVE_AND_COST);
  const bool gauss_store = store_direct && gauss_cost;
  const bool direct = !BACKWARD && (EPI == PLS_EPI_PREDICTION || deriv_direct || sums_direct);
  const double inv_noise = gauss_cost ? (1.0 / p.cost.observation_noise) : 1.0;  // as cost_derivative(): gaussian.py:75-88
  const double half_inv_noise = gauss_cost ? (1.0 / (2.0 * p.cost.observation_noise)) : 1.0;  // as cost_value(): gaussian.py:54-73
  const bool wide_store = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31u) == 0);

  int stage = 0;
  uint32_t phase = 0;  // of the chunk being consumed (stage = gc % STAGES, phase = (gc / STAGES) & 1, kept incrementally)
  int gc = 0;
  if (my_tiles > 0 && nchunks > 0) {
    int64_t rt0, ct0;
    tile_coords(0, rt0, ct0);
    load_rows(rt0 * BR, a2, crow);
  }

#pragma unroll 1
  for (int ti = 0; ti < my_tiles; ++ti) {
    int64_t rt, ct;
    tile_coords(ti, rt, ct);
    const int64_t row0 = rt * BR;
    double* sY = sYall + (ti & 1) * 128;
    if (!BACKWARD && !direct && tid < BR) sY[tid] = (row0 + tid < p.n_rows) ? p.y[row0 + tid] : 0.0;
    double yreg[RT];  // direct Gaussian epilogue: the targets of this thread's rows (requested now, used after the main loop)
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      const int64_t r = row0 + (warp * RT + h) * 8 + g;
      yreg[h] = (!BACKWARD && direct && EPI != PLS_EPI_PREDICTION && r < p.n_rows) ? p.y[r] : 0.0;
    }
#pragma unroll
    for (int h = 0; h < RT; ++h)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[h][nt][0] = 0.0;
        acc[h][nt][1] = 0.0;
      }

    if (KSRC == KSRC_CACHED) {
#pragma unroll
      for (int h = 0; h < RT; ++h) {
        const int64_t r = row0 + (warp * RT + h) * 8 + g;  // rows past n_rows are readable padding (see pls_b200.h)
        kbase[h] = BACKWARD ? p.gram + r : p.gram + r * p.ldk;
      }
    }
    // Gram values of the block being multiplied and of the next one, in two register sets that swap roles block by block.
    // No set is ever copied into the other: the only readers of a freshly produced (cached mode: freshly LOADED) set are the
    // DMMAs of the following block, a whole block of tensor work later, so the load latency hides behind it.  (With a
    // kc = kn copy ptxas hoists each move to right after the last DMMA that reads its target, ~30 DMMAs after the load.)
    double kA[LA][2][RT], kB[LA][2][RT];
    // CL = 2: the Gram values of chunk c are formed by the CTA of rank c mod 2 and mailed to the same warp and lane of the peer CTA
    // (the two CTAs hold the same rows and reduction range, so the fragment layouts coincide): 4 x 16 bytes per lane through
    // st.async, completion counted as transaction bytes on the peer's kfull barrier; the peer reads its mailbox at the start of the
    // chunk, re-arms the barrier and, once the values have been consumed, releases the mailbox with an arrive on my kfree barrier.
    auto mail_k = [&](const double (&k)[LA][2][RT]) {
      if (nsent > 0) mbar_wait_cluster(&kfree[warp], (uint32_t)((nsent - 1) & 1));  // the peer has read what I mailed last
      const uint32_t rbox = map_to_cta(smem_u32(sK + (warp * 32 + lane) * 8), crank ^ 1u);
      const uint32_t rbar = map_to_cta(smem_u32(&kfull[warp]), crank ^ 1u);
#pragma unroll
      for (int q = 0; q < LA; ++q) st_async_v2(rbox + 16u * q, k[q][0][0], k[q][1][0], rbar);
      ++nsent;
    };
    auto receive_k = [&](double (&k)[LA][2][RT]) {
      mbar_wait_cluster(&kfull[warp], (uint32_t)(nrecv & 1));
      const double2* box = reinterpret_cast<const double2*>(sK + (warp * 32 + lane) * 8);
#pragma unroll
      for (int q = 0; q < LA; ++q) {
        const double2 v = box[q];
        k[q][0][0] = v.x;
        k[q][1][0] = v.y;
      }
      ++nrecv;
    };
    if (nchunks > 0) {
      mbar_wait(&full[stage], phase);
      if (CL == 1 || crank == 0u) {
        gram_block(sP + stage * BK * sp, 0, 2 * t, kA);
        if constexpr (CL == 2) mail_k(kA);
      }
    }
    // One chunk = one 32-point stage = 4 groups, fully unrolled (264 DMMAs).  The next stage was issued two chunk times ago;
    // its Gram values (same tile only: they depend on the tile's rows) are formed before this stage's last block of DMMAs.
    // k0 holds the chunk's first block on entry; the chunk's successor block ends up in k1 (NBLK odd) or k0 (NBLK even).
    auto run_chunk = [&](int c, double (&k0)[LA][2][RT], double (&k1)[LA][2][RT]) {
      const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
      const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
      const bool more = c + 1 < nchunks;
      const unsigned char* bchunk = sB + stage * STAGE_BYTES;
      const double* Pt = sP + stage * BK * sp;
      const int base = c * BK + 2 * t;  // this thread's first reduction point of the chunk, counted from `begin`
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        double (&kc)[LA][2][RT] = (blk & 1) ? k1 : k0;
        double (&kn)[LA][2][RT] = (blk & 1) ? k0 : k1;  // the next block: of this chunk, or the first of the tile's next chunk
        if (blk + 1 < NBLK) {
          gram_block(Pt, (blk + 1) * LA, base + 8 * (blk + 1) * LA, kn);
        } else if (more) {
          mbar_wait(&full[nstage], nphase);
          gram_block(sP + nstage * BK * sp, 0, base + BK, kn);
        }
#pragma unroll
        for (int q = 0; q < LA; ++q) {
          const unsigned char* bgrp = bchunk + (blk * LA + q) * 1024;
          mma_step(bgrp + off0, kc[q][0]);
          mma_step(bgrp + off1, kc[q][1]);
        }
      }
      // chunk gc fully consumed by this warp.  The last warp to release a stage refills it with chunk gc + STAGES (which may
      // belong to a later tile of this CTA): nobody waits to issue a copy.  Forward role, last chunk of a tile: the refill is
      // deferred until the epilogue has used the stage as its staging buffer.
      __syncwarp();
      if (lane == 0) {
        if (atom_add_shared(&released[stage], 1u) == NWARPS - 1) {
          released[stage] = 0;
          if (gc + STAGES < total_gc && (BACKWARD || direct || more)) issue_from(gc, (int)ct, c, stage);
        }
      }
      stage = nstage;
      phase = nphase;
      ++gc;
    };
    // NS = 2: one chunk = TWO streamed blocks (the two 256-column halves) sharing the chunk's Gram values k0.  The first block
    // continues the half whose accumulators are in registers; the sets then swap through tensor memory and the second block
    // runs the other half, so consecutive chunks visit the halves as A B | B A | A B ...  The NEXT chunk's Gram values (k1) are
    // formed during the second block, from the points that arrived with the next chunk's first block.
    auto dmma_block = [&](const double (&k)[LA][2][RT]) {
      const unsigned char* bchunk = sB + stage * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < LA; ++q) {
        const unsigned char* bgrp = bchunk + q * 1024;
        mma_step(bgrp + off0, k[q][0]);
        mma_step(bgrp + off1, k[q][1]);
      }
    };
    auto release_block = [&](int b_cur, bool free_mailbox) {  // as in run_chunk: the last warp to release a stage refills it (no deferral: register epilogues only)
      __syncwarp();
      if (CL == 2 && free_mailbox && lane == 0) {
        // every lane's mailbox values have been consumed by the block's DMMAs: re-arm my barrier for the next delivery, then tell
        // the peer that its next one may come
        mbar_expect_tx(&kfull[warp], 32u * 64u);
        mbar_arrive_remote(map_to_cta(smem_u32(&kfree[warp]), crank ^ 1u));
      }
      if (lane == 0) {
        if (atom_add_shared(&released[stage], 1u) == NWARPS - 1) {
          released[stage] = 0;
          if (gc + STAGES < total_gc) issue_from(gc, (int)ct, b_cur, stage);
        }
      }
      stage = (stage + 1 == STAGES) ? 0 : stage + 1;
      if (stage == 0) phase ^= 1u;
      ++gc;
    };
    auto run_pair = [&](int c, double (&k0)[LA][2][RT], double (&k1)[LA][2][RT]) {
      if (c > 0) mbar_wait(&full[stage], phase);  // (chunk 0: the tile prologue waited for this block, which carries its points)
      const bool mine = CL == 1 || (unsigned)(c & 1) == crank;  // this chunk's Gram values were formed here (else: mailed by the peer)
      if constexpr (CL == 2) {
        if (!mine) receive_k(k0);
      }
      dmma_block(k0);
      release_block(2 * c, !mine);
      // park the finished half's accumulators, fetch the other half's (zero before its first block)
      if constexpr (NS == 2) {
        tmem_store64(tslot + (unsigned)(c & 1) * 128u, acc[0]);
        tmem_wait_st();
        if (c == 0) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            acc[0][nt][0] = 0.0;
            acc[0][nt][1] = 0.0;
          }
        } else {
          tmem_load64(tslot + (unsigned)((c & 1) ^ 1) * 128u, acc[0]);
        }
      }
      mbar_wait(&full[stage], phase);
      if (c + 1 < nchunks && (CL == 1 || !mine)) {  // (CL = 2: the next chunk is mine exactly when this one was not)
        gram_block(sP + stage * BK * sp, 0, (c + 1) * BK + 2 * t, k1);  // the next chunk's points came with this block
        if constexpr (CL == 2) mail_k(k1);
      }
      dmma_block(k0);
      release_block(2 * c + 1, false);
    };
#pragma unroll 1
    for (int c = 0; c < nchunks;) {
      if constexpr (NS == 2) {
        run_pair(c, kA, kB);
        ++c;
        if (c >= nchunks) break;
        run_pair(c, kB, kA);  // (a single copy of the pair + 16 register moves was tried: forward 32.1 -> 29.8 TFLOP/s)
        ++c;
      } else {
        run_chunk(c, kA, kB);
        ++c;
        if (NBLK & 1) {  // the successor block sits in kB: run the next chunk with the sets swapped
          if (c >= nchunks) break;
          run_chunk(c, kB, kA);
          ++c;
        }
      }
    }
    const int estage = (stage == 0) ? STAGES - 1 : stage - 1;  // the stage of the tile's last chunk

    // ---- epilogue ---------------------------------------------------------------------------------------------------
    // Column map: thread (g,t) holds, for column pair pr, the 4 consecutive columns j0 + 16 pr + 4 t + {0,1,2,3} as
    // acc[h][2pr][0], acc[h][2pr+1][0], acc[h][2pr][1], acc[h][2pr+1][1]; rows row0 + 8 (RT warp + h) + g.
    // NS = 2: the epilogue runs once per 256-column half -- first the half whose accumulators are in registers after the last
    // chunk, then the parked one.  vt numbers the halves of this CTA's tiles (the cost-sum hand-over below counts in them).
#pragma unroll 1
    for (int eh = 0; eh < NS; ++eh) {
    int half = 0;
    if constexpr (NS == 2) {
      const int hreg = ((nchunks - 1) & 1) ? 0 : 1;
      half = (eh == 0) ? hreg : 1 - hreg;
      if (eh == 1 && nchunks > 0) tmem_load64(tslot + (unsigned)((nchunks - 1) & 1) * 128u, acc[0]);
    }
    const int64_t j0 = (ct * CL + crank) * BJT + half * BJ;
    const int vt = ti * NS + eh;
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
      continue;
    }

    if (direct) {
      const int64_t cols_here = (p.j - j0 < BJ) ? (p.j - j0) : BJ;

      if (sums_direct) {  // the per-warp buffer is single: the previous tile's sums must have been combined (never spins)
        while (*combined < (unsigned)vt) {
        }
      }
#pragma unroll
      for (int pr = 0; pr < NPR; ++pr) {
        const int col = 16 * pr + 4 * t;
        double cs[4] = {0.0, 0.0, 0.0, 0.0};  // cost of this thread's rows, columns col .. col + 3
#pragma unroll
        for (int h = 0; h < RT; ++h) {
          const int64_t r = row0 + (warp * RT + h) * 8 + g;
          const bool rv = r < p.n_rows;
          double v[4] = {acc[h][2 * pr][0], acc[h][2 * pr + 1][0], acc[h][2 * pr][1], acc[h][2 * pr + 1][1]};
          if (EPI == PLS_EPI_COST_DERIVATIVE_AND_COST && sums_direct && !gauss_cost && p.both) {  // both from one specialised call
            const D8 b = cost_both4(sCost, sExp, yreg[h], D4{v[0], v[1], v[2], v[3]});
            cs[0] += rv ? b.c.a : 0.0;
            cs[1] += rv ? b.c.b : 0.0;
            cs[2] += rv ? b.c.c : 0.0;
            cs[3] += rv ? b.c.d : 0.0;
            v[0] = b.d.a;
            v[1] = b.d.b;
            v[2] = b.d.c;
            v[3] = b.d.d;
          } else if (sums_direct) {
            if (gauss_cost) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const double d = v[e] - yreg[h];
                cs[e] += rv ? half_inv_noise * (d * d) : 0.0;
              }
            } else {
              const D4 c4 = cost_value4(sCost, sExp, yreg[h], D4{v[0], v[1], v[2], v[3]});
              cs[0] += rv ? c4.a : 0.0;
              cs[1] += rv ? c4.b : 0.0;
              cs[2] += rv ? c4.c : 0.0;
              cs[3] += rv ? c4.d : 0.0;
            }
          }
          if (EPI == PLS_EPI_COST) continue;  // sums only
          if (EPI == PLS_EPI_COST_DERIVATIVE_AND_COST && sums_direct && !gauss_cost && p.both) {
            // (derivatives already in v)
          } else if (gauss_direct || gauss_store) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = inv_noise * (v[e] - yreg[h]);
          } else if (store_direct) {
            const D4 d = cost_derivative4(sCost, sExp, yreg[h], D4{v[0], v[1], v[2], v[3]});
            v[0] = d.a;
            v[1] = d.b;
            v[2] = d.c;
            v[3] = d.d;
          }
          if (!rv) continue;
          double* orow = p.out + r * p.ldo + j0;
          if (col + 3 < cols_here) {
            if (wide_store) {
              asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(orow + col), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]),
                           "l"(pol_stream)
                           : "memory");
            } else {
              double2* dst = reinterpret_cast<double2*>(orow + col);
              dst[0] = make_double2(v[0], v[1]);
              dst[1] = make_double2(v[2], v[3]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < cols_here) orow[col + e] = v[e];
          }
        }
        if (sums_direct) {
          // reduce-scatter over the 8 rows (g) of the warp: g bit 2 splits the 4 columns in halves, bit 1 picks one, bit 0
          // completes the sum; lane (g, t) ends with column col + 2 g2 + g1 (both members of a g0 pair hold it)
          const bool g2 = (g & 4) != 0, g1 = (g & 2) != 0;
          double k0 = g2 ? cs[2] : cs[0], k1 = g2 ? cs[3] : cs[1];
          k0 += __shfl_xor_sync(0xffffffffu, g2 ? cs[0] : cs[2], 16);
          k1 += __shfl_xor_sync(0xffffffffu, g2 ? cs[1] : cs[3], 16);
          double kk = g1 ? k1 : k0;
          kk += __shfl_xor_sync(0xffffffffu, g1 ? k0 : k1, 8);
          kk += __shfl_xor_sync(0xffffffffu, kk, 4);
          if ((g & 1) == 0) sC[warp * BJ + col + (g2 ? 2 : 0) + (g1 ? 1 : 0)] = kk;
        }
      }
      if (sums_direct) {
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = (atom_add_acq_rel_shared(&tile_done[vt & 1], 1u) == NWARPS - 1) ? 1u : 0u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {  // every warp has published this tile's sums: add the partials in warp order
          double* orow = ((EPI == PLS_EPI_COST) ? p.out + rt * p.ldo : p.out2 + rt * p.ldo2) + j0;
          for (int c = lane; c < (int)cols_here; c += 32) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) v += sC[w * BJ + c];
            orow[c] = v;
          }
          __syncwarp();
          if (lane == 0) {
            tile_done[vt & 1] = 0;
            __threadfence_block();
            *combined = (unsigned)(vt + 1);
          }
        }
      }
      if (eh == NS - 1 && ti + 1 < my_tiles) {
        int64_t rtn, ctn;
        tile_coords(ti + 1, rtn, ctn);
        load_rows(rtn * BR, a2, crow);
      }
      continue;
    }

    // Forward role, staged epilogue.  The next tile's row fragments are pulled into L2 now; they are loaded once the
    // accumulators are dead.
    int64_t next_row0 = 0;
    if (ti + 1 < my_tiles) {
      int64_t rtn, ctn;
      tile_coords(ti + 1, rtn, ctn);
      next_row0 = rtn * BR;
      const int64_t r = next_row0 + warp * RT * 8 + g;
      if (r < p.n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.rows_aug + r * sp + 4 * t));
    }
    // The F tile goes through the stage the tile's last chunk occupied (the other two stages already receive the next
    // tile's first chunks), 32 rows at a time, 16-byte chunks XOR-swizzled with the row parity (conflict-free both ways);
    // a coalesced pass then applies the cost functor -- inlined once, rolled loop -- and writes whole 128-byte lines.
    const int64_t rows_here = (p.n_rows - row0 < BR) ? (p.n_rows - row0) : BR;
    const int64_t cols_here = (p.j - j0 < BJ) ? (p.j - j0) : BJ;
    const bool dcost = (EPI == PLS_EPI_COST_DERIVATIVE || EPI == PLS_EPI_COST_DERIVATIVE_AND_COST);
    const bool with_cost = (EPI == PLS_EPI_COST || EPI == PLS_EPI_COST_DERIVATIVE_AND_COST);
    const bool store = (EPI != PLS_EPI_COST);
    constexpr int PAIRS = BJ / 2;             // double2 per row
    constexpr int PHASES = NTHREADS / PAIRS;  // a thread keeps its column pair and visits rows rphase, rphase + PHASES, ...
    constexpr int PASS_ROWS = 32;
    constexpr int NPASS = BR / PASS_ROWS;
    unsigned char* sF = sB + estage * STAGE_BYTES;  // [32 rows][BJ doubles], chunk index ^= row & 1
    const int col = 2 * (tid % PAIRS);
    const int rphase = tid / PAIRS;
    double csum0 = 0.0, csum1 = 0.0;  // cost of this thread's two columns over its rows, increasing row order
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
      __syncthreads();  // pass 0: every warp has left the main loop (the stage is free); later: the previous pass was read
#pragma unroll
      for (int h = 0; h < RT; ++h) {
        const int lr = (warp * RT + h) * 8 + g - pass * PASS_ROWS;  // row within the pass
        if (lr >= 0 && lr < PASS_ROWS) {
          unsigned char* frow = sF + (size_t)lr * (BJ * 8);
#pragma unroll
          for (int pr = 0; pr < NPR; ++pr) {
            const int ch = 8 * pr + 2 * t;  // 16-byte chunk of columns 16 pr + 4 t (+1); the next chunk holds (+2, +3)
            *reinterpret_cast<double2*>(frow + ((ch ^ (lr & 1)) << 4)) = make_double2(acc[h][2 * pr][0], acc[h][2 * pr + 1][0]);
            *reinterpret_cast<double2*>(frow + (((ch + 1) ^ (lr & 1)) << 4)) = make_double2(acc[h][2 * pr][1], acc[h][2 * pr + 1][1]);
          }
        }
      }
      __syncthreads();
      if (col < cols_here) {
#pragma unroll 2
        for (int lr = rphase; lr < PASS_ROWS; lr += PHASES) {
          const int r = pass * PASS_ROWS + lr;
          if (r >= rows_here) break;
          double2 v = *reinterpret_cast<const double2*>(sF + (size_t)lr * (BJ * 8) + ((((col >> 1)) ^ (lr & 1)) << 4));
          const double yv = (dcost || with_cost) ? sY[r] : 0.0;
          if (with_cost) {
            csum0 += cost_value(p.cost, yv, v.x);
            csum1 += cost_value(p.cost, yv, v.y);
          }
          if (store) {
            if (dcost) {
              v.x = cost_derivative(p.cost, yv, v.x);
              v.y = cost_derivative(p.cost, yv, v.y);
            }
            double* dst = p.out + (row0 + r) * p.ldo + j0 + col;
            if (col + 1 < cols_here) *reinterpret_cast<double2*>(dst) = v;
            else dst[0] = v.x;
          }
        }
      }
    }
    if (with_cost) {  // per-column sums over the tile's rows -> (tiles x J): PLS_EPI_COST writes out, the fused epilogue out2
      sC[rphase * BJ + col] = csum0;
      sC[rphase * BJ + col + 1] = csum1;
    }
    __syncthreads();  // the staging stage has been read; sC is complete
    if (with_cost && tid < cols_here) {
      double v = 0.0;
#pragma unroll
      for (int ph = 0; ph < PHASES; ++ph) v += sC[ph * BJ + tid];
      if (EPI == PLS_EPI_COST) p.out[rt * p.ldo + j0 + tid] = v;
      else p.out2[rt * p.ldo2 + j0 + tid] = v;
    }
    if (tid == 0 && gc - 1 + STAGES < total_gc) {  // (staged epilogue only: the direct one never defers)
      fence_proxy_async();  // the staging writes (generic proxy) are ordered before the copy (async proxy) into the same stage
      issue(gc - 1 + STAGES);  // the refill deferred at the tile's last chunk
    }
    if (ti + 1 < my_tiles) load_rows(next_row0, a2, crow);
    }  // halves
  }
  if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still mail into its shared memory or arrive on its barriers
  if (NS == 2) {  // every warp has drained its parking slots: release the tensor memory
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(*tmem_base_s);
  }
}

template <int NKD, bool BACKWARD, int KSRC, int RT, int EPI, int NS = 1, int CL = 1>
cudaError_t launch_one(const pls_ctx* ctx, GenGemmParams p, cudaStream_t stream) {
  using T = Tile<RT>;
  if (KSRC == KSRC_CACHED) p.sp = 0;  // no point rows are staged
  // tiles of a cluster (CL adjacent CTA tiles); the grid counts clusters until the end of this block
  int64_t grid = ((p.n_rows + T::BR - 1) / T::BR) * ((p.j + T::BJ * NS * CL - 1) / (T::BJ * NS * CL));
  if (BACKWARD) grid *= p.splits;
  else {
    const int64_t tiles = grid;
    if (grid > ctx->sm_count / CL) grid = ctx->sm_count / CL;  // persistent forward: one CTA per SM walks the tiles
    // the kernel counts the chunks a CTA streams over all its tiles in 32 bits
    if (grid > 0 && ((tiles + grid - 1) / grid) * ((p.red_total + BK - 1) / BK) * NS > 2147483647LL) return cudaErrorInvalidConfiguration;
    if (tiles + grid > 2147483647LL) return cudaErrorInvalidConfiguration;  // the kernel numbers its tiles in 32 bits
  }
  if (grid <= 0 || p.red_total <= 0) return cudaSuccess;
  grid *= CL;
  if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
  size_t smem = gen_gemm_smem_bytes<RT>(p.sp) + (CL == 2 ? sizeof(double) * NTHREADS * 8 : 0);  // + the Gram mailbox
  if ((int64_t)smem > ctx->max_smem_optin) return cudaErrorInvalidConfiguration;
  p.wbuf_ok = 0;
  if (CL == 2 && !BACKWARD && (EPI == PLS_EPI_COST || EPI == PLS_EPI_COST_DERIVATIVE_AND_COST)) return cudaErrorInvalidConfiguration;  // (no room for the per-warp sums)
  p.both = ctx->fused_functor;
  p.cost.reserved = p.cost.cost_id * 8 + p.cost.link_id * 2 + (p.cost.closed_form != 0);  // case index for cost_derivative4's fast paths
  if (CL == 1 && !BACKWARD && (int64_t)gen_gemm_smem_bytes_wbuf<RT>(p.sp) <= ctx->max_smem_optin) {
    if (gen_gemm_smem_bytes_wbuf<RT>(p.sp) > smem) smem = gen_gemm_smem_bytes_wbuf<RT>(p.sp);
    p.wbuf_ok = 1;
  }
  CUtensorMap tm3, tm2;
  cudaError_t e = make_stream_maps(ctx, p.b, p.red_total, p.ldb, T::NPR, &tm3, &tm2, &p.tma3d);
  if (e != cudaSuccess) return e;
  p.full_blocks = p.ldb / 16;
  e = cudaFuncSetAttribute(gen_gemm_kernel<NKD, BACKWARD, KSRC, RT, EPI, NS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if constexpr (CL == 1) {
    gen_gemm_kernel<NKD, BACKWARD, KSRC, RT, EPI, NS, CL><<<(unsigned)grid, NTHREADS, smem, stream>>>(p, tm3, tm2);
    return cudaGetLastError();
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gen_gemm_kernel<NKD, BACKWARD, KSRC, RT, EPI, NS, CL>, p, tm3, tm2);
  }
}

// ROLE: -1 = backward, otherwise the forward epilogue PLS_EPI_*
template <int NKD, int ROLE>
cudaError_t launch_role(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  constexpr bool BW = ROLE < 0;
  if (p.gram) {  // the cached-Gram kernels do not depend on the exponent depth: they live in the NKD = 1 units only
    if constexpr (NKD == 1) {
      return (p.rt == 1) ? launch_one<1, BW, KSRC_CACHED, 1, ROLE>(ctx, p, stream) : launch_one<1, BW, KSRC_CACHED, 2, ROLE>(ctx, p, stream);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  const bool rbf = (p.kernel_id == PLS_KERNEL_RBF);
  // 64 x 512 tiles with the second accumulator set parked in tensor memory (NS = 2): RBF, an even number of 256-column tiles,
  // and an epilogue that runs from registers (the cost-sum epilogues need room for the per-warp sums and > STAGES blocks per tile)
  if (rbf && p.rt == 1 && choose_tile_ns(ctx, p.j, false, p.n_rows, p.red_total, BW) == 2) {
    const bool sums = ROLE == PLS_EPI_COST || ROLE == PLS_EPI_COST_DERIVATIVE_AND_COST;
    const bool sums_from_registers = (int64_t)gen_gemm_smem_bytes_wbuf<1>(p.sp) <= ctx->max_smem_optin && 2 * ((p.red_total + BK - 1) / BK) > STAGES;
    // ... and, for the roles without cost sums, pairs of CTAs that share the generation of the Gram values (CL = 2): needs whole
    // 1024-column cluster tiles and room for the mailbox
    if constexpr (ROLE < 0 || ROLE == PLS_EPI_PREDICTION || ROLE == PLS_EPI_COST_DERIVATIVE) {
      if (choose_cluster(ctx, p.j, p.n_rows, p.red_total, BW) == 2 &&
          (int64_t)(gen_gemm_smem_bytes<1>(p.sp) + sizeof(double) * NTHREADS * 8) <= ctx->max_smem_optin)
        return launch_one<NKD, BW, KSRC_RBF, 1, ROLE, 2, 2>(ctx, p, stream);
    }
    if (!sums || sums_from_registers) return launch_one<NKD, BW, KSRC_RBF, 1, ROLE, 2>(ctx, p, stream);
  }
  if (p.rt == 1) return rbf ? launch_one<NKD, BW, KSRC_RBF, 1, ROLE>(ctx, p, stream) : launch_one<NKD, BW, KSRC_LINEAR, 1, ROLE>(ctx, p, stream);
  return rbf ? launch_one<NKD, BW, KSRC_RBF, 2, ROLE>(ctx, p, stream) : launch_one<NKD, BW, KSRC_LINEAR, 2, ROLE>(ctx, p, stream);
}

}  // namespace

}  // namespace pls
