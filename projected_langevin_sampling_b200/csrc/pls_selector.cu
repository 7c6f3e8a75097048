// ConditionalVariance inducing-point selector = greedy pivoted Cholesky of k(X, X)
// (reference: src/inducing_point_selectors/conditional_variance.py:27-120), HBM-bound.
//
// Per iteration i the reference does, on N-vectors:   column = round(k(x, x_j), 20); column[j] += jitter;
//   e = (column - c_j . C[:i]) / sqrt(d[j]);  C[i] = e;  d = clip(d - e^2, 0);  next pivot = last entry of argsort(d)
//   not chosen yet;  stop early when sum(d) < threshold.
// Here one kernel per iteration streams the i previous rows of C once (coalesced over n: the algorithmic traffic,
// 8*i*N bytes), fuses the column Gram, the rank-1 update, the clip and a per-block (max, index, sum) reduction; a
// one-block kernel then picks the pivot.  No host round trip inside the loop: pivot, sqrt(d_j) and the stop flag stay
// in device memory, launches are queued back to back.
//
// Ties: the first pivot is np.argmax (first maximum); later pivots are "last of argsort" -- implemented as the HIGHEST
// index among exactly equal maxima, which is what a stable argsort gives (numpy's default sort is unstable, so the
// reference itself is implementation-defined on exact ties; see DESIGN.md "selector ties").
#include "pls_aux.h"
#include "pls_common.cuh"

namespace pls {

namespace {

constexpr int CV_THREADS = 256;
constexpr int CV_CJ_CHUNK = 2048;
constexpr int CV_HDR = 8;  // scratch header doubles: [0]=sqrt(d_j) [1]=pivot [2]=stop [3]=n_selected [4]=sum(d)

struct Best {
  double val;
  long long idx;
};

__device__ __forceinline__ Best better(Best a, Best b, bool tie_low) {
  if (b.val > a.val) return b;
  if (b.val == a.val && b.idx >= 0 && (a.idx < 0 || (tie_low ? b.idx < a.idx : b.idx > a.idx))) return b;
  return a;
}

__device__ __forceinline__ void block_reduce_store(Best best, double sum, bool tie_low, double* part) {
  __shared__ double s_val[CV_THREADS / 32];
  __shared__ long long s_idx[CV_THREADS / 32];
  __shared__ double s_sum[CV_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other;
    other.val = __shfl_xor_sync(0xffffffffu, best.val, o);
    other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
    best = better(best, other, tie_low);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) {
    s_val[warp] = best.val;
    s_idx[warp] = best.idx;
    s_sum[warp] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best b = {s_val[0], s_idx[0]};
    double s = s_sum[0];
    for (int w = 1; w < CV_THREADS / 32; ++w) {
      Best o = {s_val[w], s_idx[w]};
      b = better(b, o, tie_low);
      s += s_sum[w];
    }
    part[0] = b.val;
    part[1] = __longlong_as_double(b.idx);
    part[2] = s;
  }
}

__device__ __forceinline__ double aug_dot(const double* a, const double* b, int d) {
  double s = 0.0;
  for (int k = 0; k < d; ++k) s = fma(a[k], b[k], s);
  s = fma(a[d], b[d + 1], s);
  s = fma(a[d + 1], b[d], s);
  return s;
}

__global__ void __launch_bounds__(CV_THREADS) cv_init_kernel(int kernel_id, const double* __restrict__ xa, int64_t n, int d,
                                                             int sp, double kdiag, double jitter, double* __restrict__ di,
                                                             unsigned char* __restrict__ taken, double* __restrict__ parts) {
  const int64_t i = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  Best best = {-1.0, -1};
  double sum = 0.0;
  if (i < n) {
    double v;
    if (kernel_id == PLS_KERNEL_RBF) {
      v = kdiag;  // gpytorch's diag of a stationary kernel on identical inputs is exactly the outputscale
    } else {
      const double* r = xa + i * sp;
      v = 0.0;
      for (int k = 0; k < d; ++k) v = fma(r[k], r[k], v);
    }
    v += jitter;
    di[i] = v;
    taken[i] = 0;
    best.val = v;
    best.idx = i;
    sum = fmax(v, 0.0);
  }
  block_reduce_store(best, sum, /*tie_low=*/true, parts + 3 * (int64_t)blockIdx.x);
}

// one block: reduce the per-block partials, publish the pivot
__global__ void __launch_bounds__(1024) cv_finalize_kernel(double* __restrict__ scratch, const double* __restrict__ parts,
                                                           int64_t nparts, int slot, int first, double threshold,
                                                           int has_threshold, unsigned char* __restrict__ taken,
                                                           int64_t* __restrict__ indices) {
  long long* hdr_i = reinterpret_cast<long long*>(scratch);
  if (!first && hdr_i[2] != 0) return;  // already stopped
  __shared__ double s_val[1024];
  __shared__ long long s_idx[1024];
  __shared__ double s_sum[1024];
  const bool tie_low = first != 0;
  Best best = {-1.0, -1};
  double sum = 0.0;
  for (int64_t b = threadIdx.x; b < nparts; b += 1024) {
    Best o = {parts[3 * b], __double_as_longlong(parts[3 * b + 1])};
    best = better(best, o, tie_low);
    sum += parts[3 * b + 2];
  }
  s_val[threadIdx.x] = best.val;
  s_idx[threadIdx.x] = best.idx;
  s_sum[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      Best a = {s_val[threadIdx.x], s_idx[threadIdx.x]};
      Best b = {s_val[threadIdx.x + o], s_idx[threadIdx.x + o]};
      a = better(a, b, tie_low);
      s_val[threadIdx.x] = a.val;
      s_idx[threadIdx.x] = a.idx;
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long piv = s_idx[0];
    if (piv >= 0) {
      scratch[0] = sqrt(s_val[0]);
      hdr_i[1] = piv;
      indices[slot] = piv;
      taken[piv] = 1;
      hdr_i[3] = slot + 1;
    } else {
      hdr_i[2] = 1;  // nothing left to choose
    }
    scratch[4] = s_sum[0];
    // conditional_variance.py:111-116: after choosing the next pivot, stop if tr(Kff - Qff) < threshold
    if (!first && has_threshold && s_sum[0] < threshold) hdr_i[2] = 1;
  }
}

__global__ void __launch_bounds__(CV_THREADS) cv_update_kernel(int kernel_id, const double* __restrict__ xa, int64_t n, int d,
                                                               int sp, int iter, double jitter, double* __restrict__ ci,
                                                               double* __restrict__ di, const unsigned char* __restrict__ taken,
                                                               const double* __restrict__ scratch, double* __restrict__ parts) {
  const long long* hdr_i = reinterpret_cast<const long long*>(scratch);
  if (hdr_i[2] != 0) return;
  __shared__ double s_cj[CV_CJ_CHUNK];
  __shared__ double s_piv[32];
  const int64_t piv = hdr_i[1];
  const double dj = scratch[0];
  const int64_t i = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  if (threadIdx.x < sp) s_piv[threadIdx.x] = xa[piv * sp + threadIdx.x];

  // dot = c_j . C[:iter][n], rows of C streamed once, coalesced over n
  double dot = 0.0;
  for (int l0 = 0; l0 < iter; l0 += CV_CJ_CHUNK) {
    const int lc = (iter - l0 < CV_CJ_CHUNK) ? (iter - l0) : CV_CJ_CHUNK;
    __syncthreads();
    for (int l = threadIdx.x; l < lc; l += CV_THREADS) s_cj[l] = ci[(int64_t)(l0 + l) * n + piv];
    __syncthreads();
    if (i < n) {
      const double* col = ci + (int64_t)l0 * n + i;
      int l = 0;
      for (; l + 4 <= lc; l += 4) {
        const double c0 = col[(int64_t)(l + 0) * n], c1 = col[(int64_t)(l + 1) * n];
        const double c2 = col[(int64_t)(l + 2) * n], c3 = col[(int64_t)(l + 3) * n];
        dot = fma(s_cj[l + 0], c0, dot);
        dot = fma(s_cj[l + 1], c1, dot);
        dot = fma(s_cj[l + 2], c2, dot);
        dot = fma(s_cj[l + 3], c3, dot);
      }
      for (; l < lc; ++l) dot = fma(s_cj[l], col[(int64_t)l * n], dot);
    }
  }
  __syncthreads();

  Best best = {-1.0, -1};
  double sum = 0.0;
  if (i < n) {
    double col = aug_dot(xa + i * sp, s_piv, d);
    if (kernel_id == PLS_KERNEL_RBF) col = gram_exp(col);
    col = __ddiv_rn(rint(__dmul_rn(col, 1e20)), 1e20);  // np.round(column, 20): multiply, rint, divide (:95)
    if (i == piv) col += jitter;                         // :96
    const double e = (col - dot) / dj;                   // :97
    ci[(int64_t)iter * n + i] = e;
    double dn = di[i] - e * e;  // :100-103
    dn = fmax(dn, 0.0);
    di[i] = dn;
    sum = dn;
    if (!taken[i]) {
      best.val = dn;
      best.idx = i;
    }
  }
  block_reduce_store(best, sum, /*tie_low=*/false, parts + 3 * (int64_t)blockIdx.x);
}

}  // namespace

int64_t cv_scratch_doubles(int64_t n) {
  const int64_t nb = (n + CV_THREADS - 1) / CV_THREADS;
  return CV_HDR + 3 * nb + (n + 7) / 8 + 1;
}

cudaError_t run_cv_select(const pls_ctx* ctx, int kernel_id, const double* xp_aug, int64_t n, int d, double kdiag, int m,
                          double jitter, double threshold, int has_threshold, double* ci, double* di, double* scratch,
                          int64_t* indices_out, int* n_selected_out, cudaStream_t stream) {
  (void)ctx;
  const int sp = point_stride(d);
  const int64_t nb = (n + CV_THREADS - 1) / CV_THREADS;
  if (nb > 2147483647LL) return cudaErrorInvalidConfiguration;
  double* parts = scratch + CV_HDR;
  unsigned char* taken = reinterpret_cast<unsigned char*>(parts + 3 * nb);
  cudaError_t e;
  if ((e = cudaMemsetAsync(scratch, 0, sizeof(double) * CV_HDR, stream)) != cudaSuccess) return e;
  cv_init_kernel<<<(unsigned)nb, CV_THREADS, 0, stream>>>(kernel_id, xp_aug, n, d, sp, kdiag, jitter, di, taken, parts);
  cv_finalize_kernel<<<1, 1024, 0, stream>>>(scratch, parts, nb, 0, 1, threshold, has_threshold, taken, indices_out);
  for (int i = 0; i < m - 1; ++i) {
    cv_update_kernel<<<(unsigned)nb, CV_THREADS, 0, stream>>>(kernel_id, xp_aug, n, d, sp, i, jitter, ci, di, taken, scratch, parts);
    cv_finalize_kernel<<<1, 1024, 0, stream>>>(scratch, parts, nb, i + 1, 0, threshold, has_threshold, taken, indices_out);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  long long nsel = 0;
  if ((e = cudaMemcpyAsync(&nsel, reinterpret_cast<long long*>(scratch) + 3, sizeof(long long), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  *n_selected_out = (int)nsel;
  return cudaSuccess;
}

}  // namespace pls
