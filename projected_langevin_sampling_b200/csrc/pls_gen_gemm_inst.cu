// One instantiation unit of the generated-operand GEMM per exponent depth NKD = ceil((D + 2) / 4); compiled once per
// value with -DPLS_NKD=<k> so the build can run them in parallel.
#include "pls_gen_gemm.cuh"

#ifndef PLS_NKD
#error "compile with -DPLS_NKD=<1..7>"
#endif

namespace pls {

#define PLS_CAT2(a, b) a##b
#define PLS_CAT(a, b) PLS_CAT2(a, b)

cudaError_t PLS_CAT(launch_gen_gemm_nkd, PLS_NKD)(bool backward, const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return backward ? launch_kind<PLS_NKD, true>(ctx, p, stream) : launch_kind<PLS_NKD, false>(ctx, p, stream);
}

}  // namespace pls
