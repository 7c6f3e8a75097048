// Dispatch of the generated-operand GEMM (kernel in pls_gen_gemm.cuh, instantiated per NKD in pls_gen_gemm_inst.cu).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "pls_common.cuh"
#include "pls_internal.h"

namespace pls {

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 encode_tiled() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  }();
  return fn;
}

cudaError_t make_stream_maps(const pls_ctx*, const double* b, int64_t rows, int64_t ldb, int blocks, CUtensorMap* tm3,
                             CUtensorMap* tm2, int* tma3d_ok) {
  auto encode = encode_tiled();
  if (!encode) return cudaErrorNotSupported;
  if (rows <= 0 || ldb < 2 || (ldb & 1) || rows > 0x7fffffffLL || ldb > 0x7fffffffLL) return cudaErrorInvalidValue;
  const cuuint32_t ones[3] = {1, 1, 1};
  {  // [rows][ldb], box 32 rows x 16 columns
    const cuuint64_t dims[2] = {(cuuint64_t)ldb, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ldb * 8};
    const cuuint32_t box[2] = {16, (cuuint32_t)BK};
    if (encode(tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(b), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  *tma3d_ok = 0;
  const int64_t full_blocks = ldb / 16;
  if (full_blocks > 0) {  // [ldb/16 column blocks][rows][16]: dimension 2 strides by 128 bytes INSIDE a row
    const cuuint64_t dims[3] = {16, (cuuint64_t)rows, (cuuint64_t)full_blocks};
    const cuuint64_t strides[2] = {(cuuint64_t)ldb * 8, 128};
    const cuuint32_t box[3] = {16, (cuuint32_t)BK, (cuuint32_t)blocks};
    if (encode(tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(b), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      *tma3d_ok = 1;
  }
  if (!*tma3d_ok) *tm3 = *tm2;
  return cudaSuccess;
}

#define PLS_DECL_ROLE(k, r) \
  cudaError_t launch_gen_gemm_nkd##k##_role##r(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream);
#define PLS_DECL(k) PLS_DECL_ROLE(k, 0) PLS_DECL_ROLE(k, 1) PLS_DECL_ROLE(k, 2) PLS_DECL_ROLE(k, 3) PLS_DECL_ROLE(k, 4)
PLS_DECL(1) PLS_DECL(2) PLS_DECL(3) PLS_DECL(4) PLS_DECL(5) PLS_DECL(6) PLS_DECL(7)
#undef PLS_DECL
#undef PLS_DECL_ROLE

using Launcher = cudaError_t (*)(const pls_ctx*, const GenGemmParams&, cudaStream_t);
#define PLS_ROW(k) {launch_gen_gemm_nkd##k##_role0, launch_gen_gemm_nkd##k##_role1, launch_gen_gemm_nkd##k##_role2, \
                    launch_gen_gemm_nkd##k##_role3, launch_gen_gemm_nkd##k##_role4}
static const Launcher kLaunchers[MAX_NKD][5] = {PLS_ROW(1), PLS_ROW(2), PLS_ROW(3), PLS_ROW(4), PLS_ROW(5), PLS_ROW(6), PLS_ROW(7)};
#undef PLS_ROW

// role: 0..3 = forward epilogue PLS_EPI_*, 4 = backward
static cudaError_t dispatch(int role, const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  if ((p.rt != 1 && p.rt != 2) || role < 0 || role > 4) return cudaErrorInvalidValue;
  const int nkd = p.gram ? 1 : point_ksteps(p.d);
  if (nkd < 1 || nkd > MAX_NKD) return cudaErrorInvalidValue;
  return kLaunchers[nkd - 1][role](ctx, p, stream);
}

cudaError_t launch_gen_gemm_forward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return dispatch(p.epilogue, ctx, p, stream);
}

cudaError_t launch_gen_gemm_backward(const pls_ctx* ctx, const GenGemmParams& p, cudaStream_t stream) {
  return dispatch(4, ctx, p, stream);
}

}  // namespace pls
