// One instantiation unit of the generated-operand GEMM per (exponent depth NKD = ceil(D / 4), role): compiled with
// -DPLS_NKD=<1..7> -DPLS_ROLE=<0..4> so the build can run them in parallel.  Roles 0..3 are the forward epilogues PLS_EPI_*,
// role 4 is the backward kernel.
#include "pls_gen_gemm.cuh"

#if !defined(PLS_NKD) || !defined(PLS_ROLE)
#error "compile with -DPLS_NKD=<1..7> -DPLS_ROLE=<0..4>"
#endif

namespace pls {

#define PLS_CAT4(a, b, c, d) a##b##c##d
#define PLS_NAME(a, b, c, d) PLS_CAT4(a, b, c, d)

cudaError_t PLS_NAME(launch_gen_gemm_nkd, PLS_NKD, _role, PLS_ROLE)(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return launch_role<PLS_NKD, (PLS_ROLE == 4) ? -1 : PLS_ROLE>(ctx, p, stream);
}

}  // namespace pls
