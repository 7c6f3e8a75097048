// C-ABI entry points of libpls_b200.so (declared in include/pls_b200.h): argument validation, error reporting and
// dispatch to the kernels.  Nothing here allocates device memory or synchronises (except pls_cv_select_f64).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pls_aux.h"
#include "pls_common.cuh"

namespace {

thread_local std::string g_create_error;

int fail(pls_ctx* ctx, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf;
  else g_create_error = buf;
  return 1;
}

int check_cuda(pls_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(ctx, "%s: %s", what, cudaGetErrorString(e));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_cost(pls_ctx* ctx, const pls_cost* c) {
  if (!c) return fail(ctx, "cost is NULL");
  if (c->cost_id < PLS_COST_GAUSSIAN || c->cost_id > PLS_COST_STUDENT_T) return fail(ctx, "unknown cost_id %d", c->cost_id);
  if (c->link_id < PLS_LINK_IDENTITY || c->link_id > PLS_LINK_SQUARE) return fail(ctx, "unknown link_id %d", c->link_id);
  return 0;
}

// CUDA-event pair around one launch of the hot kernel while pls_profile_begin is in effect
struct ProfileScope {
  pls_ctx* ctx;
  cudaStream_t stream;
  pls_ctx::ProfileRecord rec{};
  bool on = false;
  ProfileScope(pls_ctx* c, int role, double flops, cudaStream_t s) : ctx(c), stream(s) {
    if (!c || !c->profiling) return;
    rec.role = role;
    rec.flops = flops;
    if (cudaEventCreate(&rec.e0) != cudaSuccess) return;
    if (cudaEventCreate(&rec.e1) != cudaSuccess) {
      cudaEventDestroy(rec.e0);
      return;
    }
    cudaEventRecord(rec.e0, s);
    on = true;
  }
  ~ProfileScope() {
    if (!on) return;
    cudaEventRecord(rec.e1, stream);
    ctx->profile.push_back(rec);
  }
};

int check_kernel(pls_ctx* ctx, int kernel_id, int d) {
  if (kernel_id != PLS_KERNEL_RBF && kernel_id != PLS_KERNEL_LINEAR) return fail(ctx, "unknown kernel_id %d", kernel_id);
  if (d < 1 || d > pls::MAX_D) return fail(ctx, "input dimension d=%d outside [1, %d]", d, pls::MAX_D);
  return 0;
}

}  // namespace

extern "C" {

int pls_abi_version(void) { return PLS_ABI_VERSION; }

int pls_ctx_create(int device, pls_ctx** out) {
  if (!out) return fail(nullptr, "pls_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, "pls_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, "pls_ctx_create: device %d out of range [0, %d)", device, count);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, "pls_ctx_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, "pls_ctx_create: device %d is sm_%d%d; libpls_b200 is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  pls_ctx* c = new pls_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (const char* env = getenv("PLS_B200_TILE_RT")) c->tile_rt = atoi(env);
  if (const char* env = getenv("PLS_B200_TILE_NS")) c->tile_ns = atoi(env);
  if (const char* env = getenv("PLS_B200_FUSED_FUNCTOR")) c->fused_functor = atoi(env);
  if (const char* env = getenv("PLS_B200_CLUSTER")) c->cluster = atoi(env);
  *out = c;
  return 0;
}

void pls_ctx_destroy(pls_ctx* ctx) {
  if (ctx)
    for (auto& r : ctx->profile) {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
  delete ctx;
}

const char* pls_last_error(const pls_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int pls_sm_count(const pls_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

static void drop_profile(pls_ctx* ctx) {
  for (auto& r : ctx->profile) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->profile.clear();
}

int pls_profile_begin(pls_ctx* ctx) {
  if (!ctx) return 1;
  drop_profile(ctx);
  ctx->profiling = true;
  return 0;
}

int pls_profile_end(pls_ctx* ctx, double* out6) {
  if (!ctx) return 1;
  ctx->profiling = false;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  cudaError_t err = cudaSuccess;
  for (auto& r : ctx->profile) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (e != cudaSuccess) err = e;
    double* a = acc + (r.role == 0 ? 0 : 3);
    a[0] += ms;
    a[1] += 1.0;
    a[2] += r.flops;
  }
  drop_profile(ctx);
  if (out6)
    for (int i = 0; i < 6; ++i) out6[i] = acc[i];
  return check_cuda(ctx, err, "pls_profile_end");
}

int pls_point_stride(int d) {
  if (d < 1 || d > pls::MAX_D) return -1;
  return pls::point_stride(d);
}

int pls_backward_splits(const pls_ctx* ctx, int64_t n_rows, int64_t m, int64_t j) {
  // Enough CTAs for >= 8 waves of one CTA per SM when the reduction is long enough (splits no shorter than 8 pipeline
  // stages), and among the admissible counts the one that fills whole waves best (ties: fewer splits, less Gp traffic).
  const int sms = (ctx && ctx->sm_count > 0) ? ctx->sm_count : 148;
  const int rt = pls::choose_tile_rt(ctx, j);
  const int ns = pls::choose_tile_ns(ctx, j, false, m, n_rows, true);
  const int64_t br = pls::tile_rows(rt), bj = pls::tile_cols(rt) * ns * (ns == 2 ? pls::choose_cluster(ctx, j, m, n_rows, true) : 1);  // (the generated-Gram default)
  const int64_t tiles = ((m + br - 1) / br) * ((j + bj - 1) / bj);
  if (tiles <= 0 || n_rows <= 0) return 1;
  const int64_t chunks = (n_rows + pls::BK - 1) / pls::BK;
  int64_t max_splits = chunks / 8;
  if (max_splits < 1) max_splits = 1;
  if (max_splits > 4096) max_splits = 4096;
  int64_t want = (8LL * sms + tiles - 1) / tiles;
  if (want < 1) want = 1;
  int64_t hi = want + 64, lo = want;
  if (hi > max_splits) hi = max_splits;
  if (lo > hi) lo = (hi + 1) / 2;  // short reduction: trade split length for a full wave
  if (lo < 1) lo = 1;
  int64_t best = lo;
  double best_eff = 0.0;
  for (int64_t s = lo; s <= hi; ++s) {
    const int64_t ctas = tiles * s;
    const int64_t waves = (ctas + sms - 1) / sms;
    const double eff = (double)ctas / (double)(waves * sms);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = s;
    }
  }
  return (int)best;
}

int pls_forward_tile_rows(const pls_ctx* ctx, int64_t j) { return pls::tile_rows(pls::choose_tile_rt(ctx, j)); }

void pls_set_tile_shape(pls_ctx* ctx, int rt) {
  if (ctx) ctx->tile_rt = (rt == 1 || rt == 2) ? rt : 0;
}

void pls_set_tile_sets(pls_ctx* ctx, int ns) {
  if (ctx) ctx->tile_ns = (ns == 1 || ns == 2) ? ns : 0;
}

void pls_set_tile_cluster(pls_ctx* ctx, int mode) {
  if (ctx) ctx->cluster = (mode == 1 || mode == 2) ? mode : 0;
}

int pls_prepare_points_f64(pls_ctx* ctx, int kernel_id, const double* x, int64_t n, int d, int64_t ldx,
                           const double* inv_lengthscale, const double* centre, double c_extra, double* out, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (n < 0 || ldx < d || !out || (!x && n > 0)) return fail(ctx, "pls_prepare_points_f64: bad arguments");
  pls::DimVec ls{}, ce{};
  for (int k = 0; k < d; ++k) {
    ls.v[k] = (kernel_id == PLS_KERNEL_RBF && inv_lengthscale) ? inv_lengthscale[k] : 1.0;
    ce.v[k] = (kernel_id == PLS_KERNEL_RBF && centre) ? centre[k] : 0.0;
  }
  return check_cuda(ctx,
                    pls::launch_prepare_points(kernel_id, x, n, d, ldx, ls, ce, c_extra, pls::point_stride(d), out,
                                               (cudaStream_t)stream),
                    "pls_prepare_points_f64");
}

int pls_gram_f64(pls_ctx* ctx, int kernel_id, const double* rows_aug, int64_t n_rows, const double* cols_aug,
                 int64_t n_cols, int d, double* out, int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (n_rows < 0 || n_cols < 0 || ldo < n_cols || !out) return fail(ctx, "pls_gram_f64: bad arguments");
  return check_cuda(ctx,
                    pls::launch_gram(kernel_id, rows_aug, n_rows, cols_aug, n_cols, d, pls::point_stride(d), out, ldo,
                                     (cudaStream_t)stream),
                    "pls_gram_f64");
}

int pls_gram_fill_f64(pls_ctx* ctx, int kernel_id, const double* rows_aug, int64_t n_rows, const double* cols_aug,
                      int64_t n_cols, int d, double* out, int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (n_rows < 0 || n_cols < 0 || ldo < n_cols || !out) return fail(ctx, "pls_gram_fill_f64: bad arguments");
  if (n_rows > 65535LL * 32) return fail(ctx, "pls_gram_fill_f64: at most %lld rows per call", 65535LL * 32);
  return check_cuda(ctx,
                    pls::launch_gram_fill(kernel_id, rows_aug, n_rows, cols_aug, n_cols, d, pls::point_stride(d), out, ldo,
                                          (cudaStream_t)stream),
                    "pls_gram_fill_f64");
}

int pls_gemm_f64(pls_ctx* ctx, int trans_a, const double* a, int64_t lda, const double* b, int64_t ldb, double* c,
                 int64_t ldc, int64_t rows, int64_t j, int64_t k, void* stream) {
  if (!ctx) return 1;
  if (rows < 0 || j < 0 || k < 0 || !a || !b || !c || ldb < j || ldc < j || lda < (trans_a ? rows : k))
    return fail(ctx, "pls_gemm_f64: bad arguments");
  pls::SmallGemmParams p{};
  p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.c = c; p.ldc = ldc; p.rows = rows; p.j = j; p.k = k;
  return check_cuda(ctx, pls::launch_small_gemm(p, trans_a != 0, false, (cudaStream_t)stream), "pls_gemm_f64");
}

// Checks of a cached Gram argument (pls_*_cached_f64): k is (rows x ldk), ldk = pls_gram_cache_ld(m)
static int check_gram(pls_ctx* ctx, const char* who, const double* k, int64_t ldk, int64_t m) {
  if (!k || !aligned16(k)) return fail(ctx, "%s: the cached Gram must be a 16-byte aligned device pointer", who);
  if (ldk != pls_gram_cache_ld(m)) return fail(ctx, "%s: ldk must be pls_gram_cache_ld(m) = %lld", who, (long long)pls_gram_cache_ld(m));
  return 0;
}

static int forward_common(pls_ctx* ctx, const char* who, int kernel_id, const double* xa, int64_t n, const double* za, int64_t m,
                          int d, const double* gram, int64_t ldk, const double* w, int64_t ldw, int64_t j, int epilogue,
                          const pls_cost* cost, const double* y, double* out, int64_t ldo, double* out2, int64_t ldo2, void* stream) {
  if (!ctx) return 1;
  if (!gram && check_kernel(ctx, kernel_id, d)) return 1;
  if (epilogue < PLS_EPI_PREDICTION || epilogue > PLS_EPI_COST_DERIVATIVE_AND_COST) return fail(ctx, "%s: unknown epilogue %d", who, epilogue);
  if (n < 0 || m < 0 || j < 0 || !w || !out || (!gram && (!xa || !za))) return fail(ctx, "%s: bad arguments", who);
  if (gram && check_gram(ctx, who, gram, ldk, m)) return 1;
  if (ldw < j || (ldw & 1) || !aligned16(w)) return fail(ctx, "%s: w must be 16-byte aligned with an even ldw >= j", who);
  if (ldo < j) return fail(ctx, "%s: ldo < j", who);
  if (epilogue != PLS_EPI_COST && ((ldo & 1) || !aligned16(out))) return fail(ctx, "%s: out must be 16-byte aligned with an even ldo", who);
  if (epilogue == PLS_EPI_COST_DERIVATIVE_AND_COST && (!out2 || ldo2 < j)) return fail(ctx, "%s: bad cost-sum output", who);
  pls::GenGemmParams p{};
  if (epilogue != PLS_EPI_PREDICTION) {
    if (check_cost(ctx, cost)) return 1;
    if (!y) return fail(ctx, "%s: y is NULL", who);
    p.cost = *cost;
  }
  p.rows_aug = xa; p.n_rows = n; p.red_aug = za; p.red_total = m; p.b = w; p.ldb = ldw; p.j = j;
  p.gram = gram; p.ldk = ldk;
  p.sp = pls::point_stride(d); p.d = d; p.kernel_id = kernel_id; p.epilogue = epilogue; p.splits = 1; p.accumulate = 0;
  p.rt = pls::choose_tile_rt(ctx, j);
  p.out = out; p.ldo = ldo; p.out2 = out2; p.ldo2 = ldo2; p.y = y;
  ProfileScope scope(ctx, 0, 2.0 * (double)n * (double)m * (double)j, (cudaStream_t)stream);
  return check_cuda(ctx, pls::launch_gen_gemm_forward(ctx, p, (cudaStream_t)stream), who);
}

int64_t pls_gram_cache_ld(int64_t m) { return m <= 0 ? 128 : (m + 127) / 128 * 128; }
int64_t pls_gram_cache_rows(int64_t n) { return n <= 0 ? 128 : (n + 127) / 128 * 128; }

int pls_forward_f64(pls_ctx* ctx, int kernel_id, const double* xa, int64_t n, const double* za, int64_t m, int d,
                    const double* w, int64_t ldw, int64_t j, int epilogue, const pls_cost* cost, const double* y,
                    double* out, int64_t ldo, void* stream) {
  if (ctx && epilogue == PLS_EPI_COST_DERIVATIVE_AND_COST) return fail(ctx, "pls_forward_f64: unknown epilogue %d", epilogue);
  return forward_common(ctx, "pls_forward_f64", kernel_id, xa, n, za, m, d, nullptr, 0, w, ldw, j, epilogue, cost, y, out, ldo,
                        nullptr, 0, stream);
}

int pls_forward_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t n, int64_t m, const double* w, int64_t ldw,
                           int64_t j, int epilogue, const pls_cost* cost, const double* y, double* out, int64_t ldo, void* stream) {
  if (ctx && epilogue == PLS_EPI_COST_DERIVATIVE_AND_COST) return fail(ctx, "pls_forward_cached_f64: unknown epilogue %d", epilogue);
  if (ctx && !k) return fail(ctx, "pls_forward_cached_f64: k is NULL");
  return forward_common(ctx, "pls_forward_cached_f64", PLS_KERNEL_RBF, nullptr, n, nullptr, m, 1, k, ldk, w, ldw, j, epilogue, cost, y,
                        out, ldo, nullptr, 0, stream);
}

int pls_forward_step_f64(pls_ctx* ctx, int kernel_id, const double* xa, int64_t n, const double* za, int64_t m, int d,
                         const double* w, int64_t ldw, int64_t j, const pls_cost* cost, const double* y, double* dc,
                         int64_t lddc, double* cost_partial, int64_t ldcp, void* stream) {
  if (ctx && (!dc || !cost_partial || !y)) return fail(ctx, "pls_forward_step_f64: bad arguments");
  return forward_common(ctx, "pls_forward_step_f64", kernel_id, xa, n, za, m, d, nullptr, 0, w, ldw, j, PLS_EPI_COST_DERIVATIVE_AND_COST,
                        cost, y, dc, lddc, cost_partial, ldcp, stream);
}

int pls_forward_step_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t n, int64_t m, const double* w, int64_t ldw,
                                int64_t j, const pls_cost* cost, const double* y, double* dc, int64_t lddc, double* cost_partial,
                                int64_t ldcp, void* stream) {
  if (ctx && (!k || !dc || !cost_partial || !y)) return fail(ctx, "pls_forward_step_cached_f64: bad arguments");
  return forward_common(ctx, "pls_forward_step_cached_f64", PLS_KERNEL_RBF, nullptr, n, nullptr, m, 1, k, ldk, w, ldw, j,
                        PLS_EPI_COST_DERIVATIVE_AND_COST, cost, y, dc, lddc, cost_partial, ldcp, stream);
}

static int backward_common(pls_ctx* ctx, const char* who, int kernel_id, const double* za, int64_t m, const double* xa, int64_t n,
                           int d, const double* gram, int64_t ldk, const double* dc, int64_t lddc, int64_t j, double* gp, int64_t ldg,
                           int splits, int accumulate, void* stream) {
  if (!ctx) return 1;
  if (!gram && check_kernel(ctx, kernel_id, d)) return 1;
  if (n < 0 || m < 0 || j < 0 || !dc || !gp || splits < 1 || (!gram && (!xa || !za))) return fail(ctx, "%s: bad arguments", who);
  if (gram && check_gram(ctx, who, gram, ldk, m)) return 1;
  if (lddc < j || (lddc & 1) || !aligned16(dc)) return fail(ctx, "%s: dc must be 16-byte aligned with an even lddc >= j", who);
  if (ldg < j || (ldg & 1) || !aligned16(gp)) return fail(ctx, "%s: gp must be 16-byte aligned with an even ldg >= j", who);
  pls::GenGemmParams p{};
  p.rows_aug = za; p.n_rows = m; p.red_aug = xa; p.red_total = n; p.b = dc; p.ldb = lddc; p.j = j;
  p.gram = gram; p.ldk = ldk;
  p.sp = pls::point_stride(d); p.d = d; p.kernel_id = kernel_id; p.epilogue = -1; p.splits = splits;
  p.accumulate = accumulate; p.out = gp; p.ldo = ldg; p.y = nullptr; p.rt = pls::choose_tile_rt(ctx, j);
  if (n == 0 && !accumulate && m > 0)  // an empty row shard contributes a zero gradient (no kernel is launched)
    return check_cuda(ctx, cudaMemsetAsync(gp, 0, sizeof(double) * (size_t)splits * (size_t)m * (size_t)ldg, (cudaStream_t)stream), who);
  ProfileScope scope(ctx, 1, 2.0 * (double)n * (double)m * (double)j, (cudaStream_t)stream);
  return check_cuda(ctx, pls::launch_gen_gemm_backward(ctx, p, (cudaStream_t)stream), who);
}

int pls_backward_f64(pls_ctx* ctx, int kernel_id, const double* za, int64_t m, const double* xa, int64_t n, int d,
                     const double* dc, int64_t lddc, int64_t j, double* gp, int64_t ldg, int splits, int accumulate,
                     void* stream) {
  return backward_common(ctx, "pls_backward_f64", kernel_id, za, m, xa, n, d, nullptr, 0, dc, lddc, j, gp, ldg, splits, accumulate, stream);
}

int pls_backward_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t m, int64_t n, const double* dc, int64_t lddc,
                            int64_t j, double* gp, int64_t ldg, int splits, int accumulate, void* stream) {
  if (ctx && !k) return fail(ctx, "pls_backward_cached_f64: k is NULL");
  return backward_common(ctx, "pls_backward_cached_f64", PLS_KERNEL_RBF, nullptr, m, nullptr, n, 1, k, ldk, dc, lddc, j, gp, ldg, splits,
                         accumulate, stream);
}

int pls_reduce_splits_f64(pls_ctx* ctx, const double* gp, int splits, int64_t rows, int64_t j, int64_t ldg, double* out,
                          int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (!gp || !out || splits < 1 || rows < 0 || j < 0 || ldg < j || ldo < j) return fail(ctx, "pls_reduce_splits_f64: bad arguments");
  return check_cuda(ctx, pls::launch_reduce_splits(gp, splits, rows, j, ldg, out, ldo, (cudaStream_t)stream), "pls_reduce_splits_f64");
}

int pls_project_update_f64(pls_ctx* ctx, const double* vt, int64_t ldv, int64_t m, int64_t m_k, const double* gm,
                           int64_t ldg, const double* p_, int64_t ldp, int64_t j, const double* inv_lambda, double eta,
                           int noise_mode, const double* xi, int64_t ldxi, uint64_t seed, uint64_t step,
                           int64_t j_global_offset, int in_place, double* out, int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (!vt || !gm || !p_ || !inv_lambda || !out || m < 0 || m_k < 0 || j < 0 || ldv < m_k || ldg < j || ldp < j || ldo < j)
    return fail(ctx, "pls_project_update_f64: bad arguments");
  if (noise_mode < PLS_NOISE_NONE || noise_mode > PLS_NOISE_PHILOX) return fail(ctx, "pls_project_update_f64: unknown noise_mode %d", noise_mode);
  if (noise_mode == PLS_NOISE_GIVEN && (!xi || ldxi < j)) return fail(ctx, "pls_project_update_f64: noise buffer missing");
  if (in_place && (out != p_ || ldo != ldp)) return fail(ctx, "pls_project_update_f64: in_place requires out == p");
  pls::SmallGemmParams p{};
  p.a = vt; p.lda = ldv; p.b = gm; p.ldb = ldg; p.c = out; p.ldc = ldo; p.rows = m_k; p.j = j; p.k = m;
  p.particles = p_; p.ldp = ldp; p.inv_lambda = inv_lambda; p.xi = xi; p.ldxi = ldxi; p.eta = eta;
  p.noise_mode = noise_mode; p.in_place = in_place; p.seed = seed; p.step = step; p.j_global_offset = j_global_offset;
  p.step_counter = ctx->step_counter;
  return check_cuda(ctx, pls::launch_small_gemm(p, true, true, (cudaStream_t)stream), "pls_project_update_f64");
}

void pls_set_step_counter(pls_ctx* ctx, const uint64_t* counter_dev) {
  if (ctx) ctx->step_counter = counter_dev;
}

int pls_advance_step_counter(pls_ctx* ctx, uint64_t* counter_dev, uint64_t increment, void* stream) {
  if (!ctx) return 1;
  if (!counter_dev) return fail(ctx, "pls_advance_step_counter: NULL counter");
  return check_cuda(ctx, pls::launch_advance_counter(counter_dev, increment, (cudaStream_t)stream), "pls_advance_step_counter");
}

int pls_cost_derivative_f64(pls_ctx* ctx, const pls_cost* cost, const double* y, const double* f, int64_t ldf, int64_t n,
                            int64_t j, double* out, int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (check_cost(ctx, cost)) return 1;
  if (!y || !f || !out || n < 0 || j < 0 || ldf < j || ldo < j) return fail(ctx, "pls_cost_derivative_f64: bad arguments");
  return check_cuda(ctx, pls::launch_cost_derivative(*cost, y, f, ldf, n, j, out, ldo, ctx->sm_count, (cudaStream_t)stream),
                    "pls_cost_derivative_f64");
}

int pls_cost_value_f64(pls_ctx* ctx, const pls_cost* cost, const double* y, const double* f, int64_t ldf, int64_t n,
                       int64_t j, double* partial, double* out, void* stream) {
  if (!ctx) return 1;
  if (check_cost(ctx, cost)) return 1;
  if (!y || !f || !out || !partial || n < 0 || j < 0 || ldf < j) return fail(ctx, "pls_cost_value_f64: bad arguments");
  if (check_cuda(ctx, pls::launch_cost_value(*cost, y, f, ldf, n, j, partial, (cudaStream_t)stream), "pls_cost_value_f64")) return 1;
  return check_cuda(ctx, pls::launch_energy_terms(partial, (n + 127) / 128, j, nullptr, 0, 0, nullptr, j, out, (cudaStream_t)stream),
                    "pls_cost_value_f64");
}

int pls_energy_terms_f64(pls_ctx* ctx, const double* partial, int64_t tiles, int64_t ldpart, const double* p, int64_t ldp,
                         int64_t m_k, const double* inv_lambda, int64_t j, double* out, void* stream) {
  if (!ctx) return 1;
  if (!partial || !out || tiles < 0 || ldpart < j || j < 0) return fail(ctx, "pls_energy_terms_f64: bad arguments");
  if (p && (!inv_lambda || ldp < j || m_k < 0)) return fail(ctx, "pls_energy_terms_f64: bad particle arguments");
  return check_cuda(ctx, pls::launch_energy_terms(partial, tiles, ldpart, p, ldp, m_k, inv_lambda, j, out, (cudaStream_t)stream),
                    "pls_energy_terms_f64");
}

int pls_philox_normal_f64(pls_ctx* ctx, uint64_t seed, uint64_t step, int64_t rows, int64_t j, int64_t j_global_offset,
                          double* out, int64_t ldo, void* stream) {
  if (!ctx) return 1;
  if (!out || rows < 0 || j < 0 || ldo < j) return fail(ctx, "pls_philox_normal_f64: bad arguments");
  return check_cuda(ctx, pls::launch_philox_fill(seed, step, rows, j, j_global_offset, out, ldo, (cudaStream_t)stream),
                    "pls_philox_normal_f64");
}

int pls_gram_exp_f64(pls_ctx* ctx, const double* x, int64_t n, int fast, double* out, void* stream) {
  if (!ctx) return 1;
  if (!x || !out || n < 0) return fail(ctx, "pls_gram_exp_f64: bad arguments");
  return check_cuda(ctx, pls::launch_gram_exp(x, n, fast, out, (cudaStream_t)stream), "pls_gram_exp_f64");
}

int pls_lincomb3_f64(pls_ctx* ctx, int64_t rows, int64_t j, double a, const double* x, int64_t ldx, double b, const double* y,
                     int64_t ldy, double c, const double* z, int64_t ldz, const double* base, int64_t ldb, double* out, int64_t ldo,
                     void* stream) {
  if (!ctx) return 1;
  if (rows < 0 || j < 0 || !x || !y || !z || !out || ldx < j || ldy < j || ldz < j || ldo < j || (base && ldb < j))
    return fail(ctx, "pls_lincomb3_f64: bad arguments");
  return check_cuda(ctx, pls::launch_lincomb3(rows, j, a, x, ldx, b, y, ldy, c, z, ldz, base, ldb, out, ldo, (cudaStream_t)stream),
                    "pls_lincomb3_f64");
}

int pls_flat_math_f64(pls_ctx* ctx, int op, const double* a, const double* b, int64_t n, double* out, void* stream) {
  if (!ctx) return 1;
  if (op < 0 || op > 2 || !a || !out || n < 0 || (op == 0 && !b)) return fail(ctx, "pls_flat_math_f64: bad arguments");
  return check_cuda(ctx, pls::launch_flat_math(op, a, b ? b : a, n, out, (cudaStream_t)stream), "pls_flat_math_f64");
}

int64_t pls_cv_scratch_doubles(int64_t n, int d, int m) {
  return (n < 0 || d < 1 || d > pls::MAX_D || m < 2) ? 0 : pls::cv_scratch_doubles(n, d, m);
}

int pls_cv_select_f64(pls_ctx* ctx, int kernel_id, const double* xp_aug, int64_t n, int d, double kdiag, int m,
                      double jitter, double threshold, int has_threshold, int tie_mode, pls_cv_tie_fn tie_fn, void* tie_user,
                      double* ci, double* di, double* scratch, int64_t* indices_out, int* n_selected_out, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (!xp_aug || !ci || !di || !scratch || !indices_out || !n_selected_out) return fail(ctx, "pls_cv_select_f64: NULL argument");
  if (m < 2) return fail(ctx, "pls_cv_select_f64: Must have at least 2 inducing points");
  if (n < m) return fail(ctx, "pls_cv_select_f64: m=%d exceeds the number of points n=%lld", m, (long long)n);
  if (tie_mode != PLS_CV_TIES_HIGHEST_INDEX && tie_mode != PLS_CV_TIES_HOST) return fail(ctx, "pls_cv_select_f64: unknown tie_mode %d", tie_mode);
  if (tie_mode == PLS_CV_TIES_HOST && !tie_fn) return fail(ctx, "pls_cv_select_f64: PLS_CV_TIES_HOST needs a tie_fn");
  const cudaError_t e = pls::run_cv_select(ctx, kernel_id, xp_aug, n, d, kdiag, m, jitter, threshold, has_threshold, tie_mode, tie_fn,
                                           tie_user, ci, di, scratch, indices_out, n_selected_out, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue && tie_mode == PLS_CV_TIES_HOST)
    return fail(ctx, "pls_cv_select_f64: the tie callback returned an index that is out of range or already chosen");
  return check_cuda(ctx, e, "pls_cv_select_f64");
}

int64_t pls_cv_shard_scratch_doubles(int64_t n_local, int d, int m) {
  return (n_local < 0 || d < 1 || d > pls::MAX_D || m < 2) ? 0 : pls::cv_shard_scratch_doubles(n_local, d, m);
}
int64_t pls_cv_candidate_doubles(int d, int m) { return (d < 1 || d > pls::MAX_D || m < 2) ? 0 : pls::cv_candidate_doubles(d, m); }

int pls_cv_shard_begin_f64(pls_ctx* ctx, int kernel_id, const double* xa_local, int64_t n_local, int64_t n_offset, int d,
                           double kdiag, int m, double jitter, double* di, double* scratch, double* candidate, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (m < 2) return fail(ctx, "pls_cv_shard_begin_f64: Must have at least 2 inducing points");
  if (n_local < 0 || n_offset < 0 || !scratch || !candidate || (n_local > 0 && (!xa_local || !di)))
    return fail(ctx, "pls_cv_shard_begin_f64: bad arguments");
  return check_cuda(ctx, pls::cv_shard_begin(kernel_id, xa_local, n_local, n_offset, d, kdiag, m, jitter, di, scratch, candidate,
                                             (cudaStream_t)stream), "pls_cv_shard_begin_f64");
}

int pls_cv_shard_pick_f64(pls_ctx* ctx, const double* candidates, int world, int slot, int d, int m, double threshold,
                          int has_threshold, int tie_mode, int forced, int64_t n_local, int64_t n_offset, double* scratch,
                          int64_t* indices_out, void* stream) {
  if (!ctx) return 1;
  if (!candidates || world < 1 || slot < 0 || slot >= m || d < 1 || d > pls::MAX_D || !scratch || !indices_out ||
      (tie_mode != PLS_CV_TIES_HIGHEST_INDEX && tie_mode != PLS_CV_TIES_HOST))
    return fail(ctx, "pls_cv_shard_pick_f64: bad arguments");
  return check_cuda(ctx, pls::cv_shard_pick(candidates, world, slot, d, m, threshold, has_threshold, tie_mode, forced != 0, n_local,
                                            n_offset, scratch, indices_out, (cudaStream_t)stream), "pls_cv_shard_pick_f64");
}

int pls_cv_shard_update_f64(pls_ctx* ctx, int kernel_id, const double* xa_local, int64_t n_local, int64_t n_offset, int d, int iter,
                            int m, double jitter, double* ci, double* di, double* scratch, double* candidate, void* stream) {
  if (!ctx) return 1;
  if (check_kernel(ctx, kernel_id, d)) return 1;
  if (iter < 0 || iter >= m - 1 || n_local < 0 || !scratch || !candidate || (n_local > 0 && (!xa_local || !ci || !di)))
    return fail(ctx, "pls_cv_shard_update_f64: bad arguments");
  return check_cuda(ctx, pls::cv_shard_update(kernel_id, xa_local, n_local, n_offset, d, iter, m, jitter, ci, di, scratch, candidate,
                                              (cudaStream_t)stream), "pls_cv_shard_update_f64");
}

int pls_cv_shard_force_f64(pls_ctx* ctx, const double* xa_local, int64_t n_local, int64_t n_offset, int d, int m, int slot,
                           int64_t pivot, const double* ci, const double* di, double* scratch, double* candidate, void* stream) {
  if (!ctx) return 1;
  if (slot < 1 || slot >= m || pivot < 0 || n_local < 0 || d < 1 || d > pls::MAX_D || !scratch || !candidate ||
      (n_local > 0 && (!xa_local || !ci || !di)))
    return fail(ctx, "pls_cv_shard_force_f64: bad arguments");
  return check_cuda(ctx, pls::cv_shard_force(xa_local, n_local, n_offset, d, m, slot, pivot, ci, di, scratch, candidate,
                                             (cudaStream_t)stream), "pls_cv_shard_force_f64");
}

int pls_cv_shard_status(pls_ctx* ctx, const double* scratch, int64_t* status4, void* stream) {
  if (!ctx) return 1;
  if (!scratch || !status4) return fail(ctx, "pls_cv_shard_status: NULL argument");
  return check_cuda(ctx, pls::cv_shard_status(scratch, status4, (cudaStream_t)stream), "pls_cv_shard_status");
}

int pls_cv_shard_finish(pls_ctx* ctx, const double* scratch, int* n_selected_out, void* stream) {
  if (!ctx) return 1;
  if (!scratch || !n_selected_out) return fail(ctx, "pls_cv_shard_finish: NULL argument");
  int64_t st[4];
  if (check_cuda(ctx, pls::cv_shard_status(scratch, st, (cudaStream_t)stream), "pls_cv_shard_finish")) return 1;
  *n_selected_out = (int)st[0];
  return 0;
}

}  // extern "C"
