// Dispatch of the generated-operand GEMM (kernel in pls_gen_gemm.cuh, instantiated per NKD in pls_gen_gemm_inst.cu).
#include "pls_common.cuh"
#include "pls_internal.h"

namespace pls {

#define PLS_DECL(k) \
  cudaError_t launch_gen_gemm_nkd##k(bool backward, const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream);
PLS_DECL(1) PLS_DECL(2) PLS_DECL(3) PLS_DECL(4) PLS_DECL(5) PLS_DECL(6) PLS_DECL(7)
#undef PLS_DECL

static cudaError_t dispatch(bool backward, const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  if (p.rt != 1 && p.rt != 2) return cudaErrorInvalidValue;
  switch (point_ksteps(p.d)) {
    case 1: return launch_gen_gemm_nkd1(backward, ctx, p, stream);
    case 2: return launch_gen_gemm_nkd2(backward, ctx, p, stream);
    case 3: return launch_gen_gemm_nkd3(backward, ctx, p, stream);
    case 4: return launch_gen_gemm_nkd4(backward, ctx, p, stream);
    case 5: return launch_gen_gemm_nkd5(backward, ctx, p, stream);
    case 6: return launch_gen_gemm_nkd6(backward, ctx, p, stream);
    case 7: return launch_gen_gemm_nkd7(backward, ctx, p, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_gen_gemm_forward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return dispatch(false, ctx, p, stream);
}

cudaError_t launch_gen_gemm_backward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return dispatch(true, ctx, p, stream);
}

}  // namespace pls
