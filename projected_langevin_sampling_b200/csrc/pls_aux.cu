// Everything around the hot kernel: point-set preparation, dense Gram (setup/predict), the small dense DGEMM with the
// Langevin-update epilogue, split reduction, elementwise cost kernels and the Philox normal stream.
// None of these is on the roofline of the step (together < 0.5 % of its FP64 work at the headline shape); they are
// written for correctness, determinism and coalesced access.
#include "pls_aux.h"
#include "pls_cost.cuh"

namespace pls {

// ---------------------------------------------------------------------------------------------------------------
// augmented points
// ---------------------------------------------------------------------------------------------------------------
__global__ void prepare_points_kernel(int kernel_id, const double* __restrict__ x, int64_t n, int d, int64_t ldx,
                                      DimVec inv_ls, DimVec centre, double c_extra, int sp, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* xi = x + i * ldx;
  double* o = out + i * sp;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    double v = xi[k];
    if (kernel_id == PLS_KERNEL_RBF) {
      v = (v - centre.v[k]) * inv_ls.v[k];
      sq = fma(v, v, sq);
    }
    o[k] = v;
  }
  if (kernel_id == PLS_KERNEL_RBF) {
    o[d] = -0.5 * sq + c_extra;
    o[d + 1] = 1.0;
  } else {
    o[d] = 0.0;
    o[d + 1] = 0.0;
  }
  for (int k = d + 2; k < sp; ++k) o[k] = 0.0;
}

cudaError_t launch_prepare_points(int kernel_id, const double* x, int64_t n, int d, int64_t ldx, const DimVec& inv_ls,
                                  const DimVec& centre, double c_extra, int sp, double* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  prepare_points_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(kernel_id, x, n, d, ldx, inv_ls,
                                                                                        centre, c_extra, sp, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// dense Gram from augmented sets
// ---------------------------------------------------------------------------------------------------------------
__global__ void gram_kernel(int kernel_id, const double* __restrict__ ra, int64_t nr, const double* __restrict__ ca,
                            int64_t nc, int d, int sp, double* __restrict__ out, int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= nr || c >= nc) return;
  const double* a = ra + r * sp;
  const double* b = ca + c * sp;
  double s = 0.0;
  for (int k = 0; k < d; ++k) s = fma(a[k], b[k], s);
  s = fma(a[d], b[d + 1], s);
  s = fma(a[d + 1], b[d], s);
  out[r * ldo + c] = (kernel_id == PLS_KERNEL_RBF) ? gram_exp(s) : s;
}

cudaError_t launch_gram(int kernel_id, const double* ra, int64_t nr, const double* ca, int64_t nc, int d, int sp,
                        double* out, int64_t ldo, cudaStream_t stream) {
  if (nr <= 0 || nc <= 0) return cudaSuccess;
  dim3 block(32, 8);
  dim3 grid((unsigned)((nc + 31) / 32), (unsigned)((nr + 7) / 8));
  if (grid.y > 65535) return cudaErrorInvalidConfiguration;
  gram_kernel<<<grid, block, 0, stream>>>(kernel_id, ra, nr, ca, nc, d, sp, out, ldo);
  return cudaGetLastError();
}

// The same Gram at stream speed, for the buffers the pls_*_cached_f64 kernels read (cache or per-chunk staging): one CTA =
// 32 rows x 128 columns; each thread keeps ITS column point in registers, the rows are staged in shared memory and broadcast,
// four rows are in flight per thread, the exp is the contraction kernels' own table-driven routine (<= 2 ulp).
constexpr int GF_ROWS = 32, GF_COLS = 128;
template <int NK>  // (d + 2) rounded up to a multiple of 4
__global__ void __launch_bounds__(GF_COLS) gram_fill_kernel(int kernel_id, const double* __restrict__ ra, int64_t nr,
                                                            const double* __restrict__ ca, int64_t nc, int d, int sp,
                                                            double* __restrict__ out, int64_t ldo) {
  __shared__ double srow[GF_ROWS * NK];
  __shared__ double tbl[64];
  const int tid = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * GF_COLS + tid;
  const int64_t r0 = (int64_t)blockIdx.y * GF_ROWS;
  const int nrows = (int)((nr - r0 < GF_ROWS) ? (nr - r0) : GF_ROWS);
  if (tid < 64) tbl[tid] = kExp2Table[tid];
  // rows: coordinates, then the two affine entries, zero padded to NK; rows past the end are zeros
  for (int i = tid; i < GF_ROWS * NK; i += GF_COLS) {
    const int r = i / NK, k = i - r * NK;
    srow[i] = (r < nrows && k < d + 2) ? ra[(r0 + r) * sp + k] : 0.0;
  }
  // this thread's column point, with its two affine entries swapped so that s = sum_k row[k] * col[k]  (gram_kernel above)
  double bv[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int src = (k < d) ? k : ((k == d) ? d + 1 : d);
    bv[k] = (c < nc && k < d + 2) ? ca[c * sp + src] : 0.0;
  }
  __syncthreads();
  const bool rbf = kernel_id == PLS_KERNEL_RBF;
#pragma unroll 1
  for (int r = 0; r < nrows; r += 4) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NK; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = fma(srow[(r + i) * NK + k], bv[k], s[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double v = rbf ? gram_exp_fast(s[i], tbl) : s[i];
      if (c < nc && r + i < nrows) out[(r0 + r + i) * ldo + c] = v;
    }
  }
}

cudaError_t launch_gram_fill(int kernel_id, const double* ra, int64_t nr, const double* ca, int64_t nc, int d, int sp, double* out,
                             int64_t ldo, cudaStream_t stream) {
  if (nr <= 0 || nc <= 0) return cudaSuccess;
  dim3 grid((unsigned)((nc + GF_COLS - 1) / GF_COLS), (unsigned)((nr + GF_ROWS - 1) / GF_ROWS));
  if ((nr + GF_ROWS - 1) / GF_ROWS > 65535) return cudaErrorInvalidConfiguration;
  switch ((d + 2 + 3) / 4) {
#define PLS_GF(q) \
  case q: gram_fill_kernel<4 * q><<<grid, GF_COLS, 0, stream>>>(kernel_id, ra, nr, ca, nc, d, sp, out, ldo); break;
    PLS_GF(1) PLS_GF(2) PLS_GF(3) PLS_GF(4) PLS_GF(5) PLS_GF(6) PLS_GF(7)
#undef PLS_GF
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

__global__ void gram_exp_kernel(const double* __restrict__ x, int64_t n, int fast, double* __restrict__ out) {
  __shared__ double tbl[64];
  if (threadIdx.x < 64) tbl[threadIdx.x] = kExp2Table[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fast ? gram_exp_fast(x[i], tbl) : gram_exp(x[i]);
}

// out = (base ? base : 0) + a x + b y + c z on (rows x j) matrices: the InducingPointBasis update (inducing_point.py:140-149)
__global__ void lincomb3_kernel(int64_t rows, int64_t j, double a, const double* __restrict__ x, int64_t ldx, double b,
                                const double* __restrict__ y, int64_t ldy, double c, const double* __restrict__ z, int64_t ldz,
                                const double* base, int64_t ldb, double* out, int64_t ldo) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (col >= j || r >= rows) return;
  double v = a * x[r * ldx + col];
  v = v + b * y[r * ldy + col];
  v = v + c * z[r * ldz + col];
  if (base) v = base[r * ldb + col] + v;
  out[r * ldo + col] = v;
}

cudaError_t launch_lincomb3(int64_t rows, int64_t j, double a, const double* x, int64_t ldx, double b, const double* y, int64_t ldy,
                            double c, const double* z, int64_t ldz, const double* base, int64_t ldb, double* out, int64_t ldo,
                            cudaStream_t stream) {
  if (rows <= 0 || j <= 0) return cudaSuccess;
  if (rows > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((j + 255) / 256), (unsigned)rows);
  lincomb3_kernel<<<grid, 256, 0, stream>>>(rows, j, a, x, ldx, b, y, ldy, c, z, ldz, base, ldb, out, ldo);
  return cudaGetLastError();
}

// the branch-free arithmetic of the register epilogue (FlatMath), element by element: op 0 = a / b, 1 = log a, 2 = exp a
__global__ void flat_math_kernel(int op, const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
  __shared__ double tbl[64];
  if (threadIdx.x < 64) tbl[threadIdx.x] = kExp2Table[threadIdx.x];
  __syncthreads();
  const FlatMath m{tbl};
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (op == 0) ? m.div(a[i], b[i]) : ((op == 1) ? m.logn(a[i]) : m.expo(a[i]));
}

cudaError_t launch_flat_math(int op, const double* a, const double* b, int64_t n, double* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  flat_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(op, a, b, n, out);
  return cudaGetLastError();
}

cudaError_t launch_gram_exp(const double* x, int64_t n, int fast, double* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  gram_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n, fast, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

// standard normals for the column pair (2 * pair, 2 * pair + 1) of `row` at step `step`: one Philox block and one
// Box-Muller transform give both, so the stream does not depend on how the particle axis is sharded or tiled.
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t step, int64_t row, int64_t pair, double& even, double& odd) {
  uint32_t c[4] = {(uint32_t)row, (uint32_t)pair, (uint32_t)step, (uint32_t)(step >> 32) ^ (uint32_t)((uint64_t)row >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint64_t b1 = ((uint64_t)c[0] << 32) | c[1];
  const uint64_t b2 = ((uint64_t)c[2] << 32) | c[3];
  const double u1 = ((double)(b1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0, 1)
  const double u2 = ((double)(b2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  even = rad * cs;
  odd = rad * sn;
}

// standard normal for element (row, global column)
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t step, int64_t row, int64_t gcol) {
  double even, odd;
  philox_normal_pair(seed, step, row, gcol >> 1, even, odd);
  return (gcol & 1) ? odd : even;
}

__global__ void philox_fill_kernel(uint64_t seed, uint64_t step, int64_t rows, int64_t j, int64_t joff, double* out,
                                   int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (c >= j || r >= rows) return;
  out[r * ldo + c] = philox_normal(seed, step, r, joff + c);
}

cudaError_t launch_philox_fill(uint64_t seed, uint64_t step, int64_t rows, int64_t j, int64_t joff, double* out,
                               int64_t ldo, cudaStream_t stream) {
  if (rows <= 0 || j <= 0) return cudaSuccess;
  if (rows > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((j + 255) / 256), (unsigned)rows);
  philox_fill_kernel<<<grid, 256, 0, stream>>>(seed, step, rows, j, joff, out, ldo);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// small dense DGEMM (DMMA), 64 x 64 x 16 tiles, 4 warps of 32 x 32.  C = op(A) B with a store or Langevin-update epilogue.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16, GSA = GK + 4, GSB = GT + 4;

template <bool TRANS_A, bool UPDATE>
__global__ void __launch_bounds__(128) small_gemm_kernel(const SmallGemmParams p) {
  __shared__ double As[GT * GSA];
  __shared__ double Bs[GK * GSB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int64_t row0 = (int64_t)blockIdx.y * GT, col0 = (int64_t)blockIdx.x * GT;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  for (int64_t k0 = 0; k0 < p.k; k0 += GK) {
    // A tile -> As[r][kk]
    for (int i = tid; i < GT * GK; i += 128) {
      int r, kk;
      if (TRANS_A) {
        r = i % GT;
        kk = i / GT;
      } else {
        kk = i % GK;
        r = i / GK;
      }
      const int64_t gr = row0 + r, gk = k0 + kk;
      double v = 0.0;
      if (gr < p.rows && gk < p.k) v = TRANS_A ? p.a[gk * p.lda + gr] : p.a[gr * p.lda + gk];
      As[r * GSA + kk] = v;
    }
    for (int i = tid; i < GK * GT; i += 128) {
      const int cc = i % GT, kk = i / GT;
      const int64_t gc = col0 + cc, gk = k0 + kk;
      Bs[kk * GSB + cc] = (gc < p.j && gk < p.k) ? p.b[gk * p.ldb + gc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < GK / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = As[(wm * 32 + mt * 8 + g) * GSA + ks * 4 + t];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = Bs[(ks * 4 + t) * GSB + wn * 32 + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
    __syncthreads();
  }

  const double sq2eta = sqrt(2.0 * p.eta);
  const uint64_t step = (UPDATE && p.step_counter) ? p.step + *p.step_counter : p.step;
  // a small grid is launched gridDim.z times over: every copy forms the (cheap) product, copy z finishes row group mt == z,
  // which spreads the noise generation of the update epilogue over more SMs
  const int zsel = (gridDim.z > 1) ? (int)blockIdx.z : -1;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int64_t r = row0 + wm * 32 + mt * 8 + g;
    if (r >= p.rows || (zsel >= 0 && zsel != mt)) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int64_t cpair = col0 + wn * 32 + nt * 8 + 2 * t;  // this thread's two adjacent columns
      double xi2[2] = {0.0, 0.0};
      if (UPDATE && p.noise_mode == PLS_NOISE_PHILOX && cpair < p.j) {
        const int64_t gc = p.j_global_offset + cpair;
        if ((gc & 1) == 0) {  // the two columns are one Philox pair
          philox_normal_pair(p.seed, step, r, gc >> 1, xi2[0], xi2[1]);
        } else {
          xi2[0] = philox_normal(p.seed, step, r, gc);
          xi2[1] = philox_normal(p.seed, step, r, gc + 1);
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t c = cpair + e;
        if (c >= p.j) continue;
        double v = acc[mt][nt][e];
        if (UPDATE) {
          // delta = -eta * (V~^T k(Z,X) Dc) - eta * (1/lambda) P + sqrt(2 eta) xi      orthonormal.py:151-158
          const double pv = p.particles[r * p.ldp + c];
          double xi = xi2[e];
          if (p.noise_mode == PLS_NOISE_GIVEN) xi = p.xi[r * p.ldxi + c];
          double delta = -p.eta * v - p.eta * (p.inv_lambda[r] * pv);
          delta = delta + sq2eta * xi;
          v = p.in_place ? (pv + delta) : delta;
        }
        p.c[r * p.ldc + c] = v;
      }
    }
  }
}

__global__ void advance_counter_kernel(uint64_t* counter, uint64_t increment) { *counter += increment; }

cudaError_t launch_advance_counter(uint64_t* counter, uint64_t increment, cudaStream_t stream) {
  advance_counter_kernel<<<1, 1, 0, stream>>>(counter, increment);
  return cudaGetLastError();
}

cudaError_t launch_small_gemm(const SmallGemmParams& p, bool trans_a, bool update, cudaStream_t stream) {
  if (p.rows <= 0 || p.j <= 0) return cudaSuccess;
  dim3 grid((unsigned)((p.j + GT - 1) / GT), (unsigned)((p.rows + GT - 1) / GT));
  if (grid.y > 65535) return cudaErrorInvalidConfiguration;
  if (update && p.noise_mode == PLS_NOISE_PHILOX && (uint64_t)grid.x * grid.y <= 148) grid.z = 4;  // see the epilogue
  if (trans_a) {
    if (update) small_gemm_kernel<true, true><<<grid, 128, 0, stream>>>(p);
    else small_gemm_kernel<true, false><<<grid, 128, 0, stream>>>(p);
  } else {
    if (update) small_gemm_kernel<false, true><<<grid, 128, 0, stream>>>(p);
    else small_gemm_kernel<false, false><<<grid, 128, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// reductions / elementwise
// ---------------------------------------------------------------------------------------------------------------
__global__ void reduce_splits_kernel(const double* __restrict__ gp, int splits, int64_t rows, int64_t j, int64_t ldg,
                                     double* __restrict__ out, int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (c >= j) return;
  double s = 0.0;
  for (int k = 0; k < splits; ++k) s += gp[((int64_t)k * rows + r) * ldg + c];
  out[r * ldo + c] = s;
}

cudaError_t launch_reduce_splits(const double* gp, int splits, int64_t rows, int64_t j, int64_t ldg, double* out,
                                 int64_t ldo, cudaStream_t stream) {
  if (rows <= 0 || j <= 0) return cudaSuccess;
  if (rows > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((j + 255) / 256), (unsigned)rows);
  reduce_splits_kernel<<<grid, 256, 0, stream>>>(gp, splits, rows, j, ldg, out, ldo);
  return cudaGetLastError();
}

__global__ void cost_derivative_kernel(const pls_cost cost, const double* __restrict__ y, const double* __restrict__ f,
                                       int64_t ldf, int64_t n, int64_t j, double* __restrict__ out, int64_t ldo) {
  const int64_t total = n * j;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / j, c = i - r * j;
    out[r * ldo + c] = cost_derivative(cost, y[r], f[r * ldf + c]);
  }
}

cudaError_t launch_cost_derivative(const pls_cost& cost, const double* y, const double* f, int64_t ldf, int64_t n,
                                   int64_t j, double* out, int64_t ldo, int sm_count, cudaStream_t stream) {
  if (n <= 0 || j <= 0) return cudaSuccess;
  const int64_t total = n * j;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  cost_derivative_kernel<<<(unsigned)blocks, 256, 0, stream>>>(cost, y, f, ldf, n, j, out, ldo);
  return cudaGetLastError();
}

// partial[tile][c] = sum over the tile's (<=128) rows of c(y, F): one thread per column, rows in increasing order
__global__ void cost_value_kernel(const pls_cost cost, const double* __restrict__ y, const double* __restrict__ f,
                                  int64_t ldf, int64_t n, int64_t j, double* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tile = blockIdx.y;
  if (c >= j) return;
  const int64_t r0 = tile * 128;
  const int64_t r1 = (r0 + 128 < n) ? r0 + 128 : n;
  double s = 0.0;
  for (int64_t r = r0; r < r1; ++r) s += cost_value(cost, y[r], f[r * ldf + c]);
  partial[tile * j + c] = s;
}

cudaError_t launch_cost_value(const pls_cost& cost, const double* y, const double* f, int64_t ldf, int64_t n, int64_t j,
                              double* partial, cudaStream_t stream) {
  if (n <= 0 || j <= 0) return cudaSuccess;
  const int64_t tiles = (n + 127) / 128;
  if (tiles > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((j + 127) / 128), (unsigned)tiles);
  cost_value_kernel<<<grid, 128, 0, stream>>>(cost, y, f, ldf, n, j, partial);
  return cudaGetLastError();
}

__global__ void energy_terms_kernel(const double* __restrict__ partial, int64_t tiles, int64_t ldpart,
                                    const double* __restrict__ p, int64_t ldp, int64_t m_k,
                                    const double* __restrict__ inv_lambda, int64_t j, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j) return;
  double s = 0.0;
  for (int64_t t = 0; t < tiles; ++t) s += partial[t * ldpart + c];
  if (p != nullptr) {
    double q = 0.0;
    for (int64_t r = 0; r < m_k; ++r) {
      const double pv = p[r * ldp + c];
      q += pv * (inv_lambda[r] * pv);  // P * (diag(1/lambda) @ P), summed over rows    orthonormal.py:120-124
    }
    s = s + 0.5 * q;
  }
  out[c] = s;
}

cudaError_t launch_energy_terms(const double* partial, int64_t tiles, int64_t ldpart, const double* p, int64_t ldp,
                                int64_t m_k, const double* inv_lambda, int64_t j, double* out, cudaStream_t stream) {
  if (j <= 0) return cudaSuccess;
  energy_terms_kernel<<<(unsigned)((j + 127) / 128), 128, 0, stream>>>(partial, tiles, ldpart, p, ldp, m_k, inv_lambda, j, out);
  return cudaGetLastError();
}

}  // namespace pls
