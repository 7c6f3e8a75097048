// ConditionalVariance inducing-point selector = greedy pivoted Cholesky of k(X, X)
// (reference: src/inducing_point_selectors/conditional_variance.py:27-120), HBM-bound.
//
// Per iteration i the reference does, on N-vectors:   column = round(k(x, x_j), 20); column[j] += jitter;
//   e = (column - c_j . C[:i]) / sqrt(d[j]);  C[i] = e;  d = clip(d - e^2, 0);  next pivot = last entry of argsort(d)
//   not chosen yet;  stop early when sum(d) < threshold.
// Here one kernel per iteration streams the i previous rows of C once (coalesced over n: the algorithmic traffic,
// 8*i*N bytes), fuses the column Gram, the rank-1 update, the clip and a per-block (max, index, multiplicity, runner-up,
// sum) reduction; a one-block kernel then picks the pivot and publishes its record (augmented point + column of C) for the
// next update.  Pivot, sqrt(d_j) and the stop flag stay in device memory; launches are queued back to back.
//
// Ties.  The first pivot is np.argmax (first maximum).  Later pivots are "the last entry of np.argsort(d) not chosen yet"
// (:105-109); numpy's default argsort is unstable, so among EXACTLY equal maxima the reference's choice is whatever numpy's
// sort routine does with that array on that host.  The kernels count how often the maximum is attained.  When it is
// attained once the pivot is unambiguous and chosen on the device.  When it is attained several times the selector either
//   * pauses (PLS_CV_TIES_HOST): the host copies d, asks the caller's pls_cv_tie_fn -- the Python layer passes the
//     reference's own two lines, reversed(np.argsort(d)) -- and the chosen pivot is pushed back with a forced pick; or
//   * applies the built-in rule (PLS_CV_TIES_HIGHEST_INDEX = what a stable argsort gives), without any host round trip.
// Launches are queued in batches (growing while no tie shows up, back to one after a tie), the header is read between them.
//
// The row-sharded selector (pls_cv_shard_*) runs the SAME update kernel (identical accumulation order by construction) on
// each rank's rows; the single-GPU selector is the one-rank case with the candidate / pick pair fused into one kernel.
#include <vector>

#include "pls_aux.h"
#include "pls_common.cuh"

namespace pls {

namespace {

constexpr int CV_THREADS = 256;
constexpr int CV_CJ_CHUNK = 2048;
constexpr int CV_PART = 5;  // doubles per block partial: max, index, sum(d), multiplicity of the max, runner-up value
// scratch header (doubles / int64 bit patterns), documented in include/pls_b200.h:
//   [0] sqrt(d_j)  [1] pivot (global index, i64)  [2] stop (i64)  [3] n_selected (i64)  [4] sum(d)
//   [5] tie pending (i64)  [6] slot of the pending tie (i64)  [7] multiplicity of that tie (i64)
//   [8] min over pivots of (max - runner-up) / max  [9] number of tied picks so far (i64)
constexpr int CV_HDR = PLS_CV_HEADER_DOUBLES;
static_assert(CV_HDR >= 10, "header too small");

struct Best {
  double val;
  long long idx;  // < 0: none
  long long cnt;  // how many candidates attain val
  double val2;    // largest candidate value strictly below val (-1 if none)
};

__device__ __forceinline__ Best no_best() { return Best{-1.0, -1, 0, -1.0}; }

__device__ __forceinline__ Best better(const Best& a, const Best& b, bool tie_low) {
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  Best r;
  if (b.val > a.val) {
    r = b;
    r.val2 = fmax(b.val2, a.val);
  } else if (a.val > b.val) {
    r = a;
    r.val2 = fmax(a.val2, b.val);
  } else {
    r = (tie_low ? b.idx < a.idx : b.idx > a.idx) ? b : a;
    r.cnt = a.cnt + b.cnt;
    r.val2 = fmax(a.val2, b.val2);
  }
  return r;
}

__device__ __forceinline__ Best shfl_xor_best(const Best& b, int o) {
  Best r;
  r.val = __shfl_xor_sync(0xffffffffu, b.val, o);
  r.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
  r.cnt = __shfl_xor_sync(0xffffffffu, b.cnt, o);
  r.val2 = __shfl_xor_sync(0xffffffffu, b.val2, o);
  return r;
}

__device__ __forceinline__ void store_part(double* part, const Best& b, double sum) {
  part[0] = b.val;
  part[1] = __longlong_as_double(b.idx);
  part[2] = sum;
  part[3] = __longlong_as_double(b.cnt);
  part[4] = b.val2;
}
__device__ __forceinline__ Best load_part(const double* part) {
  return Best{part[0], __double_as_longlong(part[1]), __double_as_longlong(part[3]), part[4]};
}

// per-block (max, index, sum, ...) of an update / init launch -> parts[blockIdx.x]
__device__ __forceinline__ void block_reduce_store(Best best, double sum, bool tie_low, double* part) {
  __shared__ Best s_best[CV_THREADS / 32];
  __shared__ double s_sum[CV_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best = better(best, shfl_xor_best(best, o), tie_low);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) {
    s_best[warp] = best;
    s_sum[warp] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best b = s_best[0];
    double s = s_sum[0];
    for (int w = 1; w < CV_THREADS / 32; ++w) {
      b = better(b, s_best[w], tie_low);
      s += s_sum[w];
    }
    store_part(part, b, s);
  }
}

// One 1024-thread block reduces the per-block partials of a launch; the result is valid in EVERY thread.
__device__ __forceinline__ void reduce_parts(const double* __restrict__ parts, int64_t nparts, bool tie_low, Best* best_out,
                                             double* sum_out) {
  __shared__ Best s_best[1024];
  __shared__ double s_sum[1024];
  Best best = no_best();
  double sum = 0.0;
  for (int64_t b = threadIdx.x; b < nparts; b += 1024) {
    best = better(best, load_part(parts + CV_PART * b), tie_low);
    sum += parts[CV_PART * b + 2];
  }
  s_best[threadIdx.x] = best;
  s_sum[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_best[threadIdx.x] = better(s_best[threadIdx.x], s_best[threadIdx.x + o], tie_low);
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    }
    __syncthreads();
  }
  *best_out = s_best[0];
  *sum_out = s_sum[0];
}

__device__ __forceinline__ double aug_dot(const double* a, const double* b, int d) {
  double s = 0.0;
  for (int k = 0; k < d; ++k) s = fma(a[k], b[k], s);
  s = fma(a[d], b[d + 1], s);
  s = fma(a[d + 1], b[d], s);
  return s;
}

__global__ void __launch_bounds__(CV_THREADS) cv_init_kernel(int kernel_id, const double* __restrict__ xa, int64_t n,
                                                             int64_t n_offset, int d, int sp, double kdiag, double jitter,
                                                             double* __restrict__ di, unsigned char* __restrict__ taken,
                                                             double* __restrict__ parts) {
  const int64_t i = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  Best best = no_best();
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
    best = Best{v, i + n_offset, 1, -1.0};
    sum = fmax(v, 0.0);
  }
  block_reduce_store(best, sum, /*tie_low=*/true, parts + CV_PART * (int64_t)blockIdx.x);
}

// The rank-1 update of iteration `iter` with the pivot published in the scratch header + pivot record
// ([x_aug row | c[0:iter, pivot]]); rows [n_offset, n_offset + n) of the permuted set.  ONE kernel for the single-GPU and the
// row-sharded selector: the accumulation order of c_j . C[:iter] is the same whatever the sharding.
__global__ void __launch_bounds__(CV_THREADS) cv_update_kernel(int kernel_id, const double* __restrict__ xa, int64_t n,
                                                               int64_t n_offset, int d, int sp, int iter, double jitter,
                                                               double* __restrict__ ci, double* __restrict__ di,
                                                               const unsigned char* __restrict__ taken,
                                                               const double* __restrict__ scratch, double* __restrict__ parts) {
  const long long* hdr_i = reinterpret_cast<const long long*>(scratch);
  if (hdr_i[2] != 0 || hdr_i[5] != 0) return;  // stopped, or waiting for the host to resolve a tie
  __shared__ double s_cj[CV_CJ_CHUNK];
  __shared__ double s_piv[32];
  const int64_t piv = hdr_i[1];  // global index
  const double dj = scratch[0];
  const double* pivot = scratch + CV_HDR;
  const int64_t i = (int64_t)blockIdx.x * CV_THREADS + threadIdx.x;
  if (threadIdx.x < sp) s_piv[threadIdx.x] = pivot[threadIdx.x];

  // dot = c_j . C[:iter][n], rows of C streamed once, coalesced over n
  double dot = 0.0;
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

  Best best = no_best();
  double sum = 0.0;
  if (i < n) {
    double col = aug_dot(xa + i * sp, s_piv, d);
    if (kernel_id == PLS_KERNEL_RBF) col = gram_exp(col);
    col = __ddiv_rn(rint(__dmul_rn(col, 1e20)), 1e20);  // np.round(column, 20): multiply, rint, divide (:95)
    if (i + n_offset == piv) col += jitter;              // :96
    const double e = (col - dot) / dj;                   // :97
    ci[(int64_t)iter * n + i] = e;
    double dn = di[i] - e * e;  // :100-103
    dn = fmax(dn, 0.0);
    di[i] = dn;
    sum = dn;
    if (!taken[i]) best = Best{dn, i + n_offset, 1, -1.0};
  }
  block_reduce_store(best, sum, /*tie_low=*/false, parts + CV_PART * (int64_t)blockIdx.x);
}

// thread 0 of the deciding block: publish the pivot for `slot`, or leave the decision to the host
//   tie_mode: PLS_CV_TIES_HIGHEST_INDEX / PLS_CV_TIES_HOST; forced: the candidate IS the host's decision
__device__ __forceinline__ bool decide(double* scratch, const Best& best, double sum, int slot, int first, int tie_mode, int forced,
                                       double threshold, int has_threshold, int64_t* indices) {
  long long* hdr_i = reinterpret_cast<long long*>(scratch);
  bool published = false;
  if (first) scratch[8] = 1.0 / 0.0;
  if (best.idx < 0) {
    hdr_i[2] = 1;  // nothing left to choose
  } else if (!first && !forced && tie_mode == PLS_CV_TIES_HOST && best.cnt > 1) {
    hdr_i[5] = 1;
    hdr_i[6] = slot;
    hdr_i[7] = best.cnt;
    hdr_i[9] += 1;
    scratch[8] = 0.0;
  } else {
    scratch[0] = sqrt(best.val);
    hdr_i[1] = best.idx;
    indices[slot] = best.idx;
    hdr_i[3] = slot + 1;
    hdr_i[5] = 0;
    published = true;
    if (!first && !forced) {
      if (best.cnt > 1) {
        hdr_i[9] += 1;
        scratch[8] = 0.0;
      } else if (best.val > 0.0 && best.val2 >= 0.0) {
        scratch[8] = fmin(scratch[8], (best.val - best.val2) / best.val);
      }
    }
  }
  if (!forced) {
    scratch[4] = sum;
    // conditional_variance.py:111-116: after choosing the next pivot, stop if tr(Kff - Qff) < threshold
    if (!first && has_threshold && sum < threshold) hdr_i[2] = 1;
  }
  return published;
}

// single GPU, one block: reduce the partials, decide, and copy the pivot's record (augmented point, column of C)
__global__ void __launch_bounds__(1024) cv_finalize_kernel(double* __restrict__ scratch, const double* __restrict__ parts,
                                                           int64_t nparts, int slot, int first, int tie_mode, double threshold,
                                                           int has_threshold, const double* __restrict__ xa, int64_t n, int sp,
                                                           const double* __restrict__ ci, unsigned char* __restrict__ taken,
                                                           int64_t* __restrict__ indices) {
  long long* hdr_i = reinterpret_cast<long long*>(scratch);
  if (!first && (hdr_i[2] != 0 || hdr_i[5] != 0)) return;
  __shared__ int s_pub;
  Best best;
  double sum;
  reduce_parts(parts, nparts, first != 0, &best, &sum);
  if (threadIdx.x == 0) {
    const bool pub = decide(scratch, best, sum, slot, first, tie_mode, 0, threshold, has_threshold, indices);
    if (pub) taken[best.idx] = 1;
    s_pub = pub;
  }
  __syncthreads();
  if (!s_pub) return;
  double* pivot = scratch + CV_HDR;
  const int filled = slot;  // rows of C that exist when pivot `slot` is chosen
  for (int k = threadIdx.x; k < sp; k += 1024) pivot[k] = xa[best.idx * sp + k];
  for (int l = threadIdx.x; l < filled; l += 1024) pivot[sp + l] = ci[(int64_t)l * n + best.idx];
}

// ---------------------------------------------------------------------------------------------------------------
// Row-sharded selector: every rank holds rows [n_offset, n_offset + n_local) of the permuted points.  Per pivot each rank
// publishes ONE fixed-size candidate record; the host all-gathers the records (the only exchange) and every rank picks
// the same pivot from them.  Scratch layout: header | pivot record (x_aug row, then c[0:m-1, pivot]) | parts | taken.
//   candidate record: [0] value, [1] global index (int64 bits, -1 = none), [2] local sum(d), [3] multiplicity of the value
//                     among the rank's candidates (int64 bits), [4] the rank's runner-up value,
//                     [8, 8+SP) the candidate's augmented point, [8+SP, 8+SP+m-1) its column of C (rows filled so far)
// forced >= 0: the record describes the point with that global index (if this rank holds it) instead of the arg-max:
// how a pivot chosen by the host on a tie is pushed back.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CV_REC = 8;

__global__ void __launch_bounds__(1024) cv_candidate_kernel(const double* __restrict__ parts, int64_t nparts, int tie_low,
                                                            long long forced, const double* __restrict__ xa, int64_t n_local,
                                                            int64_t n_offset, int sp, const double* __restrict__ ci,
                                                            const double* __restrict__ di, int filled,
                                                            const double* __restrict__ scratch, double* __restrict__ cand) {
  const long long* hdr_i = reinterpret_cast<const long long*>(scratch);
  const bool stopped = hdr_i[2] != 0 && forced < 0;
  Best best = no_best();
  double sum = 0.0;
  if (!stopped) reduce_parts(parts, nparts, tie_low != 0, &best, &sum);
  if (forced >= 0) {
    const bool mine = forced >= n_offset && forced < n_offset + n_local;
    best = mine ? Best{di[forced - n_offset], forced, 1, -1.0} : no_best();
  }
  const long long loc = best.idx >= 0 ? best.idx - n_offset : -1;
  if (threadIdx.x == 0) {
    cand[0] = best.val;
    cand[1] = __longlong_as_double(best.idx);
    cand[2] = sum;
    cand[3] = __longlong_as_double(best.cnt);
    cand[4] = best.val2;
  }
  if (loc >= 0) {
    for (int k = threadIdx.x; k < sp; k += 1024) cand[CV_REC + k] = xa[loc * sp + k];
    for (int l = threadIdx.x; l < filled; l += 1024) cand[CV_REC + sp + l] = ci[(int64_t)l * n_local + loc];
  }
}

// one block: choose the pivot for `slot` among the ranks' candidates, publish it in the scratch header + pivot record
__global__ void __launch_bounds__(256) cv_pick_kernel(const double* __restrict__ cands, int world, int64_t rec, int slot,
                                                      int first, int tie_mode, int forced, int sp, int filled, double threshold,
                                                      int has_threshold, int64_t n_local, int64_t n_offset,
                                                      double* __restrict__ scratch, unsigned char* __restrict__ taken,
                                                      int64_t* __restrict__ indices) {
  long long* hdr_i = reinterpret_cast<long long*>(scratch);
  if (!first && !forced && (hdr_i[2] != 0 || hdr_i[5] != 0)) return;  // stopped / waiting for the host
  __shared__ int s_win;
  if (threadIdx.x == 0) {
    Best best = no_best();
    int win = -1;
    double sum = 0.0;
    for (int r = 0; r < world; ++r) {  // fixed rank order: the same decision on every rank
      const double* c = cands + (int64_t)r * rec;
      const Best o = Best{c[0], __double_as_longlong(c[1]), __double_as_longlong(c[3]), c[4]};
      const Best nb = better(best, o, first != 0);
      if (nb.idx != best.idx) win = r;
      best = nb;
      sum += c[2];
    }
    const bool pub = decide(scratch, best, sum, slot, first, tie_mode, forced, threshold, has_threshold, indices);
    if (pub && best.idx >= n_offset && best.idx < n_offset + n_local) taken[best.idx - n_offset] = 1;
    s_win = pub ? win : -1;
  }
  __syncthreads();
  const int win = s_win;
  if (win < 0) return;
  const double* c = cands + (int64_t)win * rec + CV_REC;
  double* pivot = scratch + CV_HDR;
  for (int k = threadIdx.x; k < sp + filled; k += 256) pivot[k] = c[k];
}

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
  l.taken = reinterpret_cast<unsigned char*>(l.parts + CV_PART * (l.nb > 0 ? l.nb : 1));
  return l;
}

}  // namespace

int64_t cv_shard_scratch_doubles(int64_t n_local, int d, int m) {
  const int64_t nb = (n_local + CV_THREADS - 1) / CV_THREADS;
  return CV_HDR + point_stride(d) + m + CV_PART * (nb > 0 ? nb : 1) + (n_local + 7) / 8 + 1;
}
int64_t cv_candidate_doubles(int d, int m) { return CV_REC + point_stride(d) + m; }
// single GPU: the same layout plus one candidate record (used when a host-resolved pivot is pushed back)
int64_t cv_scratch_doubles(int64_t n, int d, int m) { return cv_shard_scratch_doubles(n, d, m) + cv_candidate_doubles(d, m); }

cudaError_t run_cv_select(const pls_ctx* ctx, int kernel_id, const double* xp_aug, int64_t n, int d, double kdiag, int m,
                          double jitter, double threshold, int has_threshold, int tie_mode, pls_cv_tie_fn tie_fn, void* tie_user,
                          double* ci, double* di, double* scratch, int64_t* indices_out, int* n_selected_out, cudaStream_t stream) {
  (void)ctx;
  const ShardLayout l = shard_layout(scratch, n, d, m);
  const int sp = l.sp;
  const int64_t nb = l.nb;
  if (nb > 2147483647LL) return cudaErrorInvalidConfiguration;
  double* cand = scratch + cv_shard_scratch_doubles(n, d, m);
  const int64_t rec = cv_candidate_doubles(d, m);
  if (tie_mode == PLS_CV_TIES_HOST && !tie_fn) return cudaErrorInvalidValue;
  cudaError_t e;
  if ((e = cudaMemsetAsync(scratch, 0, sizeof(double) * (CV_HDR + sp + m), stream)) != cudaSuccess) return e;
  cv_init_kernel<<<(unsigned)nb, CV_THREADS, 0, stream>>>(kernel_id, xp_aug, n, 0, d, sp, kdiag, jitter, di, l.taken, l.parts);
  cv_finalize_kernel<<<1, 1024, 0, stream>>>(scratch, l.parts, nb, 0, 1, tie_mode, threshold, has_threshold, xp_aug, n, sp, ci,
                                             l.taken, indices_out);
  long long hdr[CV_HDR];
  std::vector<double> di_host;
  std::vector<int64_t> chosen;
  int i = 0, batch = (tie_mode == PLS_CV_TIES_HOST) ? 8 : m;
  while (i < m - 1) {
    const int end = (m - 1 - i < batch) ? (m - 1) : (i + batch);
    for (; i < end; ++i) {
      cv_update_kernel<<<(unsigned)nb, CV_THREADS, 0, stream>>>(kernel_id, xp_aug, n, 0, d, sp, i, jitter, ci, di, l.taken, scratch,
                                                                l.parts);
      cv_finalize_kernel<<<1, 1024, 0, stream>>>(scratch, l.parts, nb, i + 1, 0, tie_mode, threshold, has_threshold, xp_aug, n, sp,
                                                 ci, l.taken, indices_out);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (tie_mode != PLS_CV_TIES_HOST) break;
    if ((e = cudaMemcpyAsync(hdr, scratch, sizeof(hdr), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    if (hdr[5] != 0) {  // the maximum is attained hdr[7] times: the caller's rule decides (conditional_variance.py:105-109)
      const int slot = (int)hdr[6];
      di_host.resize((size_t)n);
      chosen.resize((size_t)slot);
      if ((e = cudaMemcpyAsync(di_host.data(), di, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
      if ((e = cudaMemcpyAsync(chosen.data(), indices_out, sizeof(int64_t) * (size_t)slot, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
      if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
      const int64_t piv = tie_fn(tie_user, di_host.data(), n, chosen.data(), slot);
      if (piv < 0 || piv >= n) return cudaErrorInvalidValue;
      for (int64_t c : chosen)
        if (c == piv) return cudaErrorInvalidValue;
      cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, nb, 0, piv, xp_aug, n, 0, sp, ci, di, slot, scratch, cand);
      cv_pick_kernel<<<1, 256, 0, stream>>>(cand, 1, rec, slot, 0, tie_mode, 1, sp, slot, threshold, has_threshold, n, 0, scratch,
                                            l.taken, indices_out);
      if (hdr[2] != 0) break;  // the early stop was decided together with this pivot
      i = slot;
      batch = 1;
    } else {
      if (hdr[2] != 0) break;
      batch = batch < 64 ? 2 * batch : 64;
    }
  }
  long long nsel = 0;
  if ((e = cudaMemcpyAsync(&nsel, reinterpret_cast<long long*>(scratch) + 3, sizeof(long long), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  *n_selected_out = (int)nsel;
  return cudaGetLastError();
}

// ---- row-sharded selector: host side ----------------------------------------------------------------------------------
cudaError_t cv_shard_begin(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, double kdiag, int m,
                           double jitter, double* di, double* scratch, double* cand, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  if (l.nb > 2147483647LL) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if ((e = cudaMemsetAsync(scratch, 0, sizeof(double) * (CV_HDR + l.sp + m), stream)) != cudaSuccess) return e;
  if (l.nb > 0)
    cv_init_kernel<<<(unsigned)l.nb, CV_THREADS, 0, stream>>>(kernel_id, xa, n_local, n_offset, d, l.sp, kdiag, jitter, di, l.taken,
                                                              l.parts);
  cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, l.nb, /*tie_low=*/1, -1, xa, n_local, n_offset, l.sp, nullptr, di, 0, scratch,
                                              cand);
  return cudaGetLastError();
}

cudaError_t cv_shard_pick(const double* cands, int world, int slot, int d, int m, double threshold, int has_threshold, int tie_mode,
                          int forced, int64_t n_local, int64_t n_offset, double* scratch, int64_t* indices, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  const int filled = slot > 0 ? slot : 0;  // rows of C that exist when pivot `slot` is chosen
  cv_pick_kernel<<<1, 256, 0, stream>>>(cands, world, cv_candidate_doubles(d, m), slot, slot == 0, tie_mode, forced, l.sp, filled,
                                        threshold, has_threshold, n_local, n_offset, scratch, l.taken, indices);
  return cudaGetLastError();
}

cudaError_t cv_shard_update(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, int iter, int m,
                            double jitter, double* ci, double* di, double* scratch, double* cand, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  if (l.nb > 0)
    cv_update_kernel<<<(unsigned)l.nb, CV_THREADS, 0, stream>>>(kernel_id, xa, n_local, n_offset, d, l.sp, iter, jitter, ci, di,
                                                                l.taken, scratch, l.parts);
  cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, l.nb, /*tie_low=*/0, -1, xa, n_local, n_offset, l.sp, ci, di, iter + 1, scratch,
                                              cand);
  return cudaGetLastError();
}

cudaError_t cv_shard_force(const double* xa, int64_t n_local, int64_t n_offset, int d, int m, int slot, int64_t pivot,
                           const double* ci, const double* di, double* scratch, double* cand, cudaStream_t stream) {
  const ShardLayout l = shard_layout(scratch, n_local, d, m);
  cv_candidate_kernel<<<1, 1024, 0, stream>>>(l.parts, l.nb, 0, pivot, xa, n_local, n_offset, l.sp, ci, di, slot, scratch, cand);
  return cudaGetLastError();
}

cudaError_t cv_shard_status(const double* scratch, int64_t* status4, cudaStream_t stream) {
  long long hdr[CV_HDR];
  cudaError_t e;
  if ((e = cudaMemcpyAsync(hdr, scratch, sizeof(hdr), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  status4[0] = hdr[3];
  status4[1] = hdr[2];
  status4[2] = hdr[5];
  status4[3] = hdr[6];
  return cudaSuccess;
}

}  // namespace pls
