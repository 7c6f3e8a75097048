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


// ---------------------------------------------------------------------------------------------------------------
// Row-sharded selector: every rank holds rows [n_offset, n_offset + n_local) of the permuted points.  Per pivot each rank
// publishes ONE fixed-size candidate record; the host all-gathers the records (the only exchange) and every rank picks
// the same pivot from them.  Scratch layout: header | pivot record (x_aug row, then c[0:m-1, pivot]) | parts | taken.
//   candidate record: [0] value, [1] global index (int64 bits, -1 = none), [2] local sum(d), [3] unused,
//                     [4, 4+SP) the candidate's augmented point, [4+SP, 4+SP+m-1) its column of C (rows filled so far)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cv_candidate_kernel(const double* __restrict__ parts, int64_t nparts, int tie_low,
                                                            const double* __restrict__ xa, int64_t n_local, int64_t n_offset,
                                                            int sp, const double* __restrict__ ci, int filled,
                                                            const double* __restrict__ scratch, double* __restrict__ cand) {
  const long long* hdr_i = reinterpret_cast<const long long*>(scratch);
  __shared__ double s_val[1024];
  __shared__ long long s_idx[1024];
  __shared__ double s_sum[1024];
  const bool stopped = hdr_i[2] != 0;
  Best best = {-1.0, -1};
  double sum = 0.0;
  if (!stopped) {
    for (int64_t b = threadIdx.x; b < nparts; b += 1024) {
      Best o = {parts[3 * b], __double_as_longlong(parts[3 * b + 1])};
      best = better(best, o, tie_low != 0);
      sum += parts[3 * b + 2];
    }
  }
  s_val[threadIdx.x] = best.val;
  s_idx[threadIdx.x] = best.idx;
  s_sum[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      Best a = {s_val[threadIdx.x], s_idx[threadIdx.x]};
      Best b = {s_val[threadIdx.x + o], s_idx[threadIdx.x + o]};
      a = better(a, b, tie_low != 0);
      s_val[threadIdx.x] = a.val;
      s_idx[threadIdx.x] = a.idx;
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    }
    __syncthreads();
  }
  const long long loc = s_idx[0];
  if (threadIdx.x == 0) {
    cand[0] = s_val[0];
    cand[1] = __longlong_as_double(loc >= 0 ? loc + n_offset : -1LL);
    cand[2] = s_sum[0];
    cand[3] = 0.0;
  }
  if (loc >= 0) {
    for (int k = threadIdx.x; k < sp; k += 1024) cand[4 + k] = xa[loc * sp + k];
    for (int l = threadIdx.x; l < filled; l += 1024) cand[4 + sp + l] = ci[(int64_t)l * n_local + loc];
  }
}

// one block: choose the pivot for `slot` among the ranks' candidates, publish it in the scratch header + pivot record
__global__ void __launch_bounds__(256) cv_pick_kernel(const double* __restrict__ cands, int world, int64_t rec, int slot,
                                                      int first, int sp, int filled, double threshold, int has_threshold,
                                                      int64_t n_local, int64_t n_offset, double* __restrict__ scratch,
                                                      unsigned char* __restrict__ taken, int64_t* __restrict__ indices) {
  long long* hdr_i = reinterpret_cast<long long*>(scratch);
  if (!first && hdr_i[2] != 0) return;  // already stopped
  __shared__ int s_win;
  if (threadIdx.x == 0) {
    Best best = {-1.0, -1};
    int win = -1;
    double sum = 0.0;
    for (int r = 0; r < world; ++r) {  // fixed rank order: the same decision on every rank
      const double* c = cands + (int64_t)r * rec;
      Best o = {c[0], __double_as_longlong(c[1])};
      const Best nb = better(best, o, first != 0);
      if (nb.idx != best.idx) win = r;
      best = nb;
      sum += c[2];
    }
    s_win = win;
    if (best.idx >= 0) {
      scratch[0] = sqrt(best.val);
      hdr_i[1] = best.idx;
      indices[slot] = best.idx;
      if (best.idx >= n_offset && best.idx < n_offset + n_local) taken[best.idx - n_offset] = 1;
      hdr_i[3] = slot + 1;
    } else {
      hdr_i[2] = 1;  // nothing left to choose
    }
    scratch[4] = sum;
    if (!first && has_threshold && sum < threshold) hdr_i[2] = 1;  // conditional_variance.py:111-116
  }
  __syncthreads();
  const int win = s_win;
  if (win < 0) return;
  const double* c = cands + (int64_t)win * rec + 4;
  double* pivot = scratch + CV_HDR;
  for (int k = threadIdx.x; k < sp + filled; k += 256) pivot[k] = c[k];
}

__global__ void __launch_bounds__(CV_THREADS) cv_update_sharded_kernel(int kernel_id, const double* __restrict__ xa, int64_t n,
                                                                       int64_t n_offset, int d, int sp, int iter, double jitter,
                                                                       double* __restrict__ ci, double* __restrict__ di,
                                                                       const unsigned char* __restrict__ taken,
                                                                       const double* __restrict__ scratch,
                                                                       double* __restrict__ parts) {
  const long long* hdr_i = reinterpret_cast<const long long*>(scratch);
  if (hdr_i[2] != 0) return;
  __shared__ double s_cj[CV_CJ_CHUNK];
  __shared__ double s_piv[32];
  const int64_t piv = hdr_i[1];  // global index
  const double dj = scratch[0];
  const double* pivot = scratch + CV_HDR;  // [x_aug row | c[0:iter, pivot]]
  const int64_t i = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  if (threadIdx.x < sp) s_piv[threadIdx.x] = pivot[threadIdx.x];

  double dot = 0.0;  // same order of accumulation as cv_update_kernel: results are identical to the unsharded selector
  for (int l0 = 0; l0 < iter; l0 += CV_CJ_CHUNK) {
    const int lc = (iter - l0 < CV_CJ_CHUNK) ? (iter - l0) : CV_CJ_CHUNK;
    __syncthreads();
    for (int l = threadIdx.x; l < lc; l += CV_THREADS) s_cj[l] = pivot[sp + l0 + l];
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
    col = __ddiv_rn(rint(__dmul_rn(col, 1e20)), 1e20);
    if (i + n_offset == piv) col += jitter;
    const double e = (col - dot) / dj;
    ci[(int64_t)iter * n + i] = e;
    double dn = di[i] - e * e;
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

// ---- row-sharded selector: host side ----------------------------------------------------------------------------------
int64_t cv_shard_scratch_doubles(int64_t n_local, int d, int m) {
  const int64_t nb = (n_local + CV_THREADS - 1) / CV_THREADS;
  return CV_HDR + point_stride(d) + m + 3 * (nb > 0 ? nb : 1) + (n_local + 7) / 8 + 1;
}
int64_t cv_candidate_doubles(int d, int m) { return 4 + point_stride(d) + m; }

namespace {
struct ShardLayout {
  int sp;
  int64_t nb;
  double* parts;
  unsigned char* taken;
};
ShardLayout shard_layout(double* scratch, int64_t n_local, int d, int m) {
  ShardLayout l;
  l.sp = point_stride(d);
  l.nb = (n_local + CV_THREADS - 1) / CV_THREADS;
  l.parts = scratch + CV_HDR + l.sp + m;
  l.taken = reinterpret_cast<unsigned char*>(l.parts + 3 * (l.nb > 0 ? l.nb : 1));
  return l;
}
}  // namespace

cudaError_t cv_shard_begin(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, double kdiag, int m,
                           double jitter, double* di, double* scratch, double* cand, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  if (l.nb > 2147483647LL) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if ((e = cudaMemsetAsync(scratch, 0, sizeof(double) * (CV_HDR + l.sp + m), stream)) != cudaSuccess) return e;
  if (l.nb > 0)
    cv_init_kernel<<<(unsigned)l.nb, CV_THREADS, 0, stream>>>(kernel_id, xa, n_local, d, l.sp, kdiag, jitter, di, l.taken, l.parts);
  cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, l.nb, /*tie_low=*/1, xa, n_local, n_offset, l.sp, nullptr, 0, scratch, cand);
  return cudaGetLastError();
}

cudaError_t cv_shard_pick(const double* cands, int world, int slot, int d, int m, double threshold, int has_threshold,
                          int64_t n_local, int64_t n_offset, double* scratch, int64_t* indices, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  const int filled = slot > 0 ? slot : 0;  // rows of C that exist when pivot `slot` is chosen
  cv_pick_kernel<<<1, 256, 0, stream>>>(cands, world, cv_candidate_doubles(d, m), slot, slot == 0, l.sp, filled, threshold,
                                        has_threshold, n_local, n_offset, scratch, l.taken, indices);
  return cudaGetLastError();
}

cudaError_t cv_shard_update(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, int iter, int m,
                            double jitter, double* ci, double* di, double* scratch, double* cand, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  if (l.nb > 0)
    cv_update_sharded_kernel<<<(unsigned)l.nb, CV_THREADS, 0, stream>>>(kernel_id, xa, n_local, n_offset, d, l.sp, iter, jitter, ci,
                                                                        di, l.taken, scratch, l.parts);
  cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, l.nb, /*tie_low=*/0, xa, n_local, n_offset, l.sp, ci, iter + 1, scratch, cand);
  return cudaGetLastError();
}

cudaError_t cv_shard_finish(const double* scratch, int* n_selected_out, cudaStream_t stream) {
  long long nsel = 0;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(&nsel, reinterpret_cast<const long long*>(scratch) + 3, sizeof(long long), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  *n_selected_out = (int)nsel;
  return cudaSuccess;
}

}  // namespace pls
