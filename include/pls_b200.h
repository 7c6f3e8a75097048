/*
 * pls_b200.h -- C ABI of the B200-native Projected Langevin Sampling (PLS) hot path.
 *
 * One shared library, libpls_b200.so (built from projected_langevin_sampling_b200/csrc/ by
 * __graft_entry__.build()).  Plain pointers and sizes only: no torch types, no C++ in the signatures.
 *
 * The reference (jswu18/projected-langevin-sampling) is pure Python and has no FFI layer; the boundary a
 * maintainer would bind is its Python class API.  Each entry point below names the reference method whose body
 * it replaces (paths relative to the reference root, `pls/` = src/projected_langevin_sampling/).  The host-side
 * mirror of that API that calls these functions lives in projected_langevin_sampling_b200/ (ctypes).
 *
 * Conventions
 *   - every matrix is row-major float64 in DEVICE memory, described by (pointer, leading dimension in elements);
 *   - N training points, M inducing points, M_k kept eigen-directions, J particles, D input dimension;
 *   - all work is enqueued on the caller's stream (`stream` = cudaStream_t passed as void*); nothing synchronises
 *     unless stated; buffers are caller-owned, the library allocates nothing on the device;
 *   - every function returns 0 on success, non-zero otherwise (message via pls_last_error); nothing throws,
 *     nothing exits;
 *   - entry points are not re-entrant on the same ctx.
 *
 * Layout "augmented points" (built once by pls_prepare_points_f64): a point set (n x D) is stored as n rows of
 * SP = pls_point_stride(D) doubles: [ (x_d - centre_d) / lengthscale_d  (D) | c | 1 | 0 ... ] with
 * c = -0.5 * |scaled x|^2 (+ log outputscale for the inducing set), so that one dot product of a row-side vector
 * with a column-side vector whose c/1 entries are swapped is the exponent of the RBF/ARD x Scale kernel.  For the
 * linear (test-double) kernel the row is [x | 0 | 0 | 0 ...] and the Gram entry is the dot product itself.
 */
#ifndef PLS_B200_H_
#define PLS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLS_ABI_VERSION 2

/* base kernel k(x, x') -- pls/kernel.py:21-29 `.base_kernel`; gpytorch ScaleKernel(RBFKernel(ard_num_dims=D)) at the
 * reference call sites pls/basis/orthonormal.py:36-41, src/inducing_point_selectors/conditional_variance.py:66-89;
 * LINEAR is the reference's test double mockers/kernel.py:8-23. */
enum { PLS_KERNEL_RBF = 0, PLS_KERNEL_LINEAR = 1 };

/* costs -- pls/costs/{gaussian,bernoulli,poisson,multimodal,student_t}.py */
enum { PLS_COST_GAUSSIAN = 0, PLS_COST_BERNOULLI = 1, PLS_COST_POISSON = 2, PLS_COST_MULTIMODAL = 3, PLS_COST_STUDENT_T = 4 };
/* link functions -- pls/link_functions.py:30-80 */
enum { PLS_LINK_IDENTITY = 0, PLS_LINK_SIGMOID = 1, PLS_LINK_PROBIT = 2, PLS_LINK_SQUARE = 3 };
/* what pls_forward_f64 writes */
enum {
  PLS_EPI_PREDICTION = 0,      /* F = k(X,Z) W                      (n x J)        orthonormal.py:98-108            */
  PLS_EPI_COST_DERIVATIVE = 1, /* d_2 c(y, F)                       (n x J)        PLS.calculate_cost_derivative    */
  PLS_EPI_COST = 2,            /* per row tile: sum_n c(y_n,F)      (tiles x J)    PLS.calculate_cost               */
  PLS_EPI_COST_DERIVATIVE_AND_COST = 3 /* both of the above in one pass (pls_forward_step_f64 only)                  */
};
/* Langevin noise source for pls_project_update_f64 */
enum {
  PLS_NOISE_NONE = 0,   /* xi = 0 (deterministic drift only; used by tests)                                       */
  PLS_NOISE_GIVEN = 1,  /* xi read from a caller buffer (parity mode: the reference's torch.normal stream)        */
  PLS_NOISE_PHILOX = 2  /* xi = Philox4x32-10 + Box-Muller keyed on (seed, step, global row, global particle)     */
};

typedef struct pls_ctx pls_ctx;

/* A likelihood cost with its link function.
 *   closed_form != 0 selects the reference's hand-written derivative where it has one (gaussian/identity
 *   gaussian.py:75-88, bernoulli/sigmoid bernoulli.py:64-77, poisson/square poisson.py:68-82, student_t/identity
 *   student_t.py:74-88); otherwise the derivative is the chain rule (d cost/d mu) * link'(F), i.e. what the
 *   reference's autograd fallback costs/base.py:68-84 evaluates (multimodal.py:79-91 is autograd-only).
 *   observation_noise: gaussian = variance (gaussian.py:71,86); multimodal = std, squared inside (multimodal.py:56).
 *   probit_divisor: the value the reference divides by inside erf (link_functions.py:42: sqrt(tensor(2.0)) in the
 *   default dtype at call time). */
typedef struct pls_cost {
  int32_t cost_id;
  int32_t link_id;
  int32_t closed_form;
  int32_t reserved;
  double observation_noise;
  double shift;
  double bernoulli_noise;
  double degrees_of_freedom;
  double scale;
  double link_jitter;
  double probit_divisor;
  /* multimodal only: log(bernoulli_noise), log(1 - bernoulli_noise) and log(sqrt(2 pi s^2)) as the reference forms them
   * (multimodal.py:55-72 builds them with torch.tensor(...) in the default dtype at call time) */
  double log_weight_1;
  double log_weight_2;
  double log_normaliser;
} pls_cost;

/* ---- context ------------------------------------------------------------------------------------------------ */
int pls_abi_version(void);
int pls_ctx_create(int device, pls_ctx** ctx);
void pls_ctx_destroy(pls_ctx* ctx);
/* last error message of `ctx` (or of the failed pls_ctx_create when ctx == NULL); never NULL */
const char* pls_last_error(const pls_ctx* ctx);
/* number of SMs of the ctx device (grids are sized from it) */
int pls_sm_count(const pls_ctx* ctx);

/* ---- layout helpers (host only, no device work) --------------------------------------------------------------- */
/* SP: doubles per augmented point row for input dimension d (d >= 1; returns -1 if d is unsupported) */
int pls_point_stride(int d);
/* number of N-splits pls_backward_f64 wants for an (m x j) gradient reduced over n_rows training rows */
int pls_backward_splits(const pls_ctx* ctx, int64_t n_rows, int64_t m, int64_t j);

/* rows per forward tile for a launch over j particle columns: the row granularity of the PLS_EPI_COST partial sums
 * (64 for the 64 x 256 tile, 128 for the 128 x 128 tile used when j <= 128) */
int pls_forward_tile_rows(const pls_ctx* ctx, int64_t j);
/* force the CTA tile shape of pls_forward_f64 / pls_backward_f64: rt = 1 (64 rows x 256 particles), 2 (128 x 128),
 * 0 = choose per launch (default).  Benchmarks and tests only; the environment variable PLS_B200_TILE_RT does the same. */
void pls_set_tile_shape(pls_ctx* ctx, int rt);
/* accumulator sets per CTA tile of the generated-Gram kernels: ns = 0 (default) parks a second 256-column accumulator set in
 * tensor memory (64 x 512 tiles: every generated Gram value feeds 512 particles) when the particle slice is an even number of
 * 256-column tiles and the launch is large enough for the wider tile to pay (>= 8 chunks per tile, enough tiles for every SM);
 * ns = 1 never does; ns = 2 does whenever the shape allows it.  Results are bit-identical either way (each column is accumulated
 * in the same order).  Benchmarks and tests only; the environment variable PLS_B200_TILE_NS does the same. */
void pls_set_tile_sets(pls_ctx* ctx, int ns);
/* pairs of CTAs (a thread-block cluster of 2) on adjacent 512-column tiles sharing the generated Gram values through distributed
 * shared memory: mode = 0 (default) the backward role when the launch is large enough, 1 never, 2 both roles whenever the particle
 * slice is a whole number of 1024-column tiles.  Bit-identical results.  Environment: PLS_B200_CLUSTER. */
void pls_set_tile_cluster(pls_ctx* ctx, int mode);

/* ---- one-time setup ------------------------------------------------------------------------------------------ */
/* Builds the augmented layout of a point set.  x: n x d (ldx).  inv_lengthscale, centre: HOST arrays of d doubles
 * (ignored for PLS_KERNEL_LINEAR).  c_extra is added to the c entry (log(outputscale) for the inducing set, 0 for
 * the training set).  out: n x pls_point_stride(d).
 * Replaces the `x.div(lengthscale)` + centring + norm padding of the gpytorch kernel call in
 * pls/basis/orthonormal.py:36-41. */
int pls_prepare_points_f64(pls_ctx* ctx, int kernel_id, const double* x, int64_t n, int d, int64_t ldx,
                           const double* inv_lengthscale, const double* centre, double c_extra, double* out,
                           void* stream);
/* Dense Gram out[i][j] = k(rows_i, cols_j) from two augmented sets (setup K_zz for the eigendecomposition,
 * pls/basis/orthonormal.py:36-38,46-48; k(x*, Z) at predict time, orthonormal.py:232-235). */
int pls_gram_f64(pls_ctx* ctx, int kernel_id, const double* rows_aug, int64_t n_rows, const double* cols_aug,
                 int64_t n_cols, int d, double* out, int64_t ldo, void* stream);

/* ---- the Langevin step, piece by piece ------------------------------------------------------------------------ */
/* C (rows x j) = op(A) * B, op(A) = A (rows x k, lda) or A^T (A is k x rows, lda) when trans_a != 0.
 * Used for W = V~ P (pls/basis/orthonormal.py:106-108, re-associated) and wherever a small dense product is needed. */
int pls_gemm_f64(pls_ctx* ctx, int trans_a, const double* a, int64_t lda, const double* b, int64_t ldb, double* c,
                 int64_t ldc, int64_t rows, int64_t j, int64_t k, void* stream);

/* Forward contraction with on-the-fly Gram tiles: for training rows [0, n) of xa,
 *   F[n][j] = sum_m k(x_n, z_m) W[m][j],      W = V~ P  (m x j, ldw, ldw even, 16-byte aligned)
 * and the epilogue selected by `epilogue`.  y (n doubles) and cost are read for the cost epilogues only.
 * out: n x j (ldo) for PREDICTION / COST_DERIVATIVE (ldo even, 16-byte aligned), ceil(n / pls_forward_tile_rows(ctx, j)) x j (ldo) for COST.
 * Replaces OrthonormalBasis.calculate_untransformed_train_prediction_samples (pls/basis/orthonormal.py:98-108) fused
 * with <Cost>.calculate_cost_derivative / calculate_cost (pls/costs/*.py); the N x M Gram is never materialised. */
int pls_forward_f64(pls_ctx* ctx, int kernel_id, const double* xa, int64_t n, const double* za, int64_t m, int d,
                    const double* w, int64_t ldw, int64_t j, int epilogue, const pls_cost* cost, const double* y,
                    double* out, int64_t ldo, void* stream);

/* pls_forward_f64 with the COST_DERIVATIVE epilogue that ALSO writes the per-row-tile cost sums of the COST epilogue
 * (cost_partial: ceil(n / pls_forward_tile_rows(ctx, j)) x j, ldcp) from the same F tile: one forward per Langevin step
 * serves both the gradient and the energy potential the training loop reads every step
 * (experiments/trainers.py:153-158 runs PLS.calculate_particle_update AND PLS.calculate_energy_potential, i.e. two
 * forwards, per epoch). */
int pls_forward_step_f64(pls_ctx* ctx, int kernel_id, const double* xa, int64_t n, const double* za, int64_t m, int d,
                         const double* w, int64_t ldw, int64_t j, const pls_cost* cost, const double* y, double* dc,
                         int64_t lddc, double* cost_partial, int64_t ldcp, void* stream);

/* Back-projection with on-the-fly Gram tiles, split `splits` ways over the training rows [0, n) of xa:
 *   gp[s][m][j] (+)= sum_{n in split s} k(z_m, x_n) dc[n][j]
 * dc: n x j (lddc even, 16-byte aligned).  gp: splits x m x ldg.  accumulate != 0 adds to gp instead of overwriting
 * (used when the training set is processed in row chunks).
 * Replaces the `k(Z,X) @ cost_derivative` product of OrthonormalBasis._calculate_particle_update
 * (pls/basis/orthonormal.py:151-155). */
int pls_backward_f64(pls_ctx* ctx, int kernel_id, const double* za, int64_t m, const double* xa, int64_t n, int d,
                     const double* dc, int64_t lddc, int64_t j, double* gp, int64_t ldg, int splits, int accumulate,
                     void* stream);

/* ---- the same contractions with the Gram kept resident in HBM ----------------------------------------------------
 * The reference holds k(Z, X) as a (lazy) tensor for the whole run (pls/basis/orthonormal.py:36-41,104-108,151-155).  When
 * n x m doubles fit in device memory the caller may do the same: k = k(X, Z) from pls_gram_f64, row-major,
 *   ldk = pls_gram_cache_ld(m) (m rounded up to 128), readable for pls_gram_cache_rows(n) rows (n rounded up to 128;
 *   the padding -- columns [m, ldk) and the rows past n -- is read and multiplied by zeros: it MUST be finite),
 * and the three entry points below run the contraction kernels with their Gram values LOADED (one streaming 8-byte read
 * per value and tile) instead of generated, which takes the exponent DMMAs and the exp off the FP64 pipe.  `k` may point
 * at any row of a larger cache (row chunks).  Arguments and results are otherwise those of pls_forward_f64,
 * pls_forward_step_f64 and pls_backward_f64 (equal to round-off: pls_gram_f64's values and the kernels' generated ones are
 * two evaluations of the same exponent, each within an ulp). */
/* pls_gram_f64 at stream speed for filling such a buffer (whole cache, or a row chunk re-filled every step): same arguments,
 * the exp is the contraction kernels' own table-driven routine (within 2 ulp of pls_gram_f64's); at most 2 097 120 rows a call. */
int pls_gram_fill_f64(pls_ctx* ctx, int kernel_id, const double* rows_aug, int64_t n_rows, const double* cols_aug,
                      int64_t n_cols, int d, double* out, int64_t ldo, void* stream);
int64_t pls_gram_cache_ld(int64_t m);
int64_t pls_gram_cache_rows(int64_t n);
int pls_forward_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t n, int64_t m, const double* w, int64_t ldw,
                           int64_t j, int epilogue, const pls_cost* cost, const double* y, double* out, int64_t ldo, void* stream);
int pls_forward_step_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t n, int64_t m, const double* w, int64_t ldw,
                                int64_t j, const pls_cost* cost, const double* y, double* dc, int64_t lddc, double* cost_partial,
                                int64_t ldcp, void* stream);
int pls_backward_cached_f64(pls_ctx* ctx, const double* k, int64_t ldk, int64_t m, int64_t n, const double* dc, int64_t lddc,
                            int64_t j, double* gp, int64_t ldg, int splits, int accumulate, void* stream);

/* out[r][c] = sum_s gp[s][r][c] in increasing s (deterministic). */
int pls_reduce_splits_f64(pls_ctx* ctx, const double* gp, int splits, int64_t rows, int64_t j, int64_t ldg,
                          double* out, int64_t ldo, void* stream);

/* delta = -eta * V~^T gm - eta * diag(inv_lambda) P + sqrt(2 eta) xi      (pls/basis/orthonormal.py:151-158)
 * vt: M x M_k (ldv) scaled eigenvectors; gm: M x J (ldg) = k(Z,X) dc; p: M_k x J (ldp).
 * in_place == 0: out (M_k x J, ldo) = delta  (PLS.calculate_particle_update, pls/projected_langevin_sampling.py:107-123)
 * in_place != 0: out must alias p; p += delta  (the caller's `particles += particle_update`,
 *                experiments/trainers.py:153-157).
 * noise_mode GIVEN reads xi (M_k x J, ldxi) -- the reference's draw, src/samplers.py:30-35 via orthonormal.py:141-145;
 * PHILOX generates xi[r][c] from (seed, step, r, j_global_offset + c), independent of how J is sharded. */
int pls_project_update_f64(pls_ctx* ctx, const double* vt, int64_t ldv, int64_t m, int64_t m_k, const double* gm,
                           int64_t ldg, const double* p, int64_t ldp, int64_t j, const double* inv_lambda, double eta,
                           int noise_mode, const double* xi, int64_t ldxi, uint64_t seed, uint64_t step,
                           int64_t j_global_offset, int in_place, double* out, int64_t ldo, void* stream);

/* CUDA-graph support for the Philox stream: a captured pls_project_update_f64 freezes its `step` argument, so a replayed
 * graph would repeat its noise.  With a counter installed, the step index the kernel uses is `step + *counter_dev` (read on
 * the device at run time); pls_advance_step_counter enqueues `*counter_dev += increment` and is captured with the step.
 * pls_set_step_counter(ctx, NULL) restores the plain behaviour. */
void pls_set_step_counter(pls_ctx* ctx, const uint64_t* counter_dev);
int pls_advance_step_counter(pls_ctx* ctx, uint64_t* counter_dev, uint64_t increment, void* stream);

/* Elementwise d_2 c(y, F) on a caller-provided F (n x j): <Cost>.calculate_cost_derivative, pls/costs/*.py. */
int pls_cost_derivative_f64(pls_ctx* ctx, const pls_cost* cost, const double* y, const double* f, int64_t ldf,
                            int64_t n, int64_t j, double* out, int64_t ldo, void* stream);
/* out[j] = sum_n c(y_n, F[n][j]): <Cost>.calculate_cost, pls/costs/*.py.  partial: workspace of
 * ceil(n/128) x j doubles. */
int pls_cost_value_f64(pls_ctx* ctx, const pls_cost* cost, const double* y, const double* f, int64_t ldf, int64_t n,
                       int64_t j, double* partial, double* out, void* stream);
/* out[c] = sum_t partial[t][c] (+ 0.5 * sum_r p[r][c]^2 * inv_lambda[r] when p != NULL): the per-particle energy of
 * OrthonormalBasis.calculate_energy_potential (pls/basis/orthonormal.py:110-126) before the mean. */
int pls_energy_terms_f64(pls_ctx* ctx, const double* partial, int64_t tiles, int64_t ldpart, const double* p,
                         int64_t ldp, int64_t m_k, const double* inv_lambda, int64_t j, double* out, void* stream);
/* Standard-normal Philox stream exactly as pls_project_update_f64 generates it (for tests and for callers that
 * want the noise explicitly). */
int pls_philox_normal_f64(pls_ctx* ctx, uint64_t seed, uint64_t step, int64_t rows, int64_t j,
                          int64_t j_global_offset, double* out, int64_t ldo, void* stream);

/* out[i] = exp(x[i]) with the exponent routine the kernels use for Gram entries: fast != 0 is the table-driven
 * branch-free variant of the hot loop (csrc/pls_common.cuh gram_exp_fast), fast == 0 the CUDA library exp used by the
 * dense Gram and the selector.  Exposed so the tests can bound the hot loop's exponent error in ulps. */
int pls_gram_exp_f64(pls_ctx* ctx, const double* x, int64_t n, int fast, double* out, void* stream);

/* out = (base ? base : 0) + a x + b y + c z on (rows x j) matrices.  The particle update of the InducingPointBasis,
 * -eta k(Z,X) Dc - eta M k(Z,Z)^{-1} P + sqrt(2 eta) e (pls/basis/inducing_point.py:140-149), from its three terms. */
int pls_lincomb3_f64(pls_ctx* ctx, int64_t rows, int64_t j, double a, const double* x, int64_t ldx, double b, const double* y,
                     int64_t ldy, double c, const double* z, int64_t ldz, const double* base, int64_t ldb, double* out, int64_t ldo,
                     void* stream);

/* out[i] = a[i] / b[i] (op 0), log a[i] (op 1) or exp a[i] (op 2) with the BRANCH-FREE routines the register epilogue of the
 * hot kernel uses for the cost functors (csrc/pls_cost.cuh FlatMath).  Exposed so the tests can bound their error in ulps. */
int pls_flat_math_f64(pls_ctx* ctx, int op, const double* a, const double* b, int64_t n, double* out, void* stream);

/* ---- one call per Langevin step ---------------------------------------------------------------------------------
 * PLS.calculate_particle_update (pls/projected_langevin_sampling.py:107-123 -> pls/basis/orthonormal.py:98-108,
 * pls/costs/<cost>.py, orthonormal.py:128-159) as ONE call over caller-owned workspaces: W = V~ P, the row-chunk loop
 * (forward with the cost derivative in its epilogue, backward), the split reduction and -- pls_step_f64 -- the update.
 * pls_step_plan_f64 (host only) fixes the chunking of the training rows and the layout of the workspace:
 *   dc_budget_bytes  bound on the only N-sized intermediate, the Dc chunk (<= 0: 32 GiB -- the headline N = 1M x J = 4096 in one chunk);
 *   gram_mode        PLS_GRAM_GENERATED: Gram tiles regenerated inside the kernels, nothing N x M in memory (default path);
 *                    PLS_GRAM_STAGED:    k(X_c, Z) of the chunk in flight re-formed every step into a chunk-sized region of the
 *                                        workspace (pls_gram_fill_f64) and streamed by that chunk's two launches (chunks are then
 *                                        also bounded by 2 GiB of Gram values: 262 144 rows at M = 1024);
 *                    PLS_GRAM_CACHED:    the caller keeps k(X, Z) (pls_gram_f64 layout of the pls_*_cached_f64 calls) and passes it;
 *   with_cost        reserve room for the per-row-tile cost sums (needed when cost_sums / energy_out is requested).
 * The workspace is plan->workspace_bytes bytes of device memory, 256-byte aligned; after pls_grad_f64 the gradient
 * G' = k(Z,X) d_2 c(y, k(X,Z) V~ P) (M x J, leading dimension plan->ldj) sits at byte offset plan->off_gm and W at plan->off_w. */
enum { PLS_GRAM_GENERATED = 0, PLS_GRAM_STAGED = 1, PLS_GRAM_CACHED = 2 };
typedef struct pls_step_plan {
  int64_t n, m, m_k, j, ldj;
  int64_t chunk_rows;   /* training rows per forward/backward launch pair */
  int64_t cost_tiles;   /* rows of the cost-partial region: sum over chunks of ceil(rows / tile_rows) */
  int32_t n_chunks, splits, tile_rows, gram_mode, with_cost, reserved;
  int64_t off_w, off_gm, off_dc, off_gp, off_cost_partial, off_kstage; /* byte offsets into the workspace */
  int64_t workspace_bytes;
} pls_step_plan;
int pls_step_plan_f64(const pls_ctx* ctx, int64_t n, int64_t m, int64_t m_k, int64_t j, int64_t dc_budget_bytes, int gram_mode,
                      int with_cost, pls_step_plan* plan);
/* The gradient G' into the workspace (see above).  xa: n x SP, za: m x SP augmented points; vt: M x M_k (ldv) or NULL when `p`
 * already holds W (M x J, ldp even, 16-byte aligned: the InducingPointBasis passes k(Z,Z)^{-1} P); p: M_k x J particles (ldp);
 * gram/ldk: the resident Gram for PLS_GRAM_CACHED, else NULL/0; cost_sums (nullable): J doubles, sum_n c(y_n, F[n][j]) of the SAME
 * forward pass (PLS.calculate_cost, pls/projected_langevin_sampling.py:75-88).  Row-sharded callers all-reduce the gradient (and
 * cost_sums) over their row group and then call pls_project_update_f64. */
int pls_grad_f64(pls_ctx* ctx, const pls_step_plan* plan, int kernel_id, int d, const double* xa, const double* za, const double* vt,
                 int64_t ldv, const double* p, int64_t ldp, const pls_cost* cost, const double* y, const double* gram, int64_t ldk,
                 void* workspace, double* cost_sums, void* stream);
/* pls_grad_f64 followed by pls_project_update_f64 (same noise / in_place / out arguments).  energy_out (nullable): J doubles, the
 * per-particle energy potential of the INPUT particles, cost + 1/2 sum_m P_mj^2 / lambda_m (orthonormal.py:110-126 before its mean),
 * from the same forward pass -- what experiments/trainers.py:153-158 pays a second forward for. */
int pls_step_f64(pls_ctx* ctx, const pls_step_plan* plan, int kernel_id, int d, const double* xa, const double* za, const double* vt,
                 int64_t ldv, const double* inv_lambda, double* p, int64_t ldp, const pls_cost* cost, const double* y,
                 const double* gram, int64_t ldk, double eta, int noise_mode, const double* xi, int64_t ldxi, uint64_t seed,
                 uint64_t step, int64_t j_global_offset, int in_place, double* out, int64_t ldo, double* energy_out, void* workspace,
                 void* stream);

/* ---- per-role kernel timing ---------------------------------------------------------------------------------------
 * Between pls_profile_begin and pls_profile_end every launch of the contraction kernel (through any entry point above) is
 * bracketed by a CUDA-event pair on the launching stream.  pls_profile_end synchronises those events and returns
 * out6 = {forward ms, forward launches, forward algorithmic flops (2 n m j each), backward ms, launches, flops}.
 * bench.py's roofline figures come from this; it costs two event records per launch and is off by default. */
int pls_profile_begin(pls_ctx* ctx);
int pls_profile_end(pls_ctx* ctx, double* out6);

/* ---- ConditionalVariance inducing-point selector ------------------------------------------------------------- */
/* Greedy pivoted Cholesky (src/inducing_point_selectors/conditional_variance.py:27-120) on the ALREADY PERMUTED
 * points xp_aug (augmented layout, n rows; the numpy permutation of :60 stays on the host).
 *   kdiag: the value diag k(x, x) takes (outputscale for RBF -- exact, as gpytorch returns it), ignored for LINEAR
 *          where the diagonal is computed;
 *   ci: workspace (m-1) x n doubles; di: workspace n doubles; scratch: workspace pls_cv_scratch_doubles(n, d, m) doubles;
 *   indices_out: m int64 in DEVICE memory, pre-filled by the caller with the sentinel n (as :63); receives positions in
 *          the permuted order, entries never reached keep the sentinel;
 *   n_selected_out (host int*): how many entries were filled.  Synchronises the stream before returning.
 *
 * Exact ties.  The reference takes "the last entry of np.argsort(d) that is not chosen yet" (:105-109).  numpy's default
 * argsort is not stable, so when the largest conditional variance is attained by SEVERAL points (1-D inputs with a short
 * lengthscale do this: every point far from all pivots keeps d = outputscale + jitter to the last bit) the reference's choice
 * is whatever numpy's sort routine makes of that array.  The kernels count how often the maximum is attained:
 *   PLS_CV_TIES_HOST          when it is attained more than once the selector pauses, copies d (n doubles) and the pivots
 *                             chosen so far to the host and calls tie_fn, which returns the position of the next pivot
 *                             (the Python layer passes exactly the reference's two lines); unambiguous pivots never
 *                             leave the device;
 *   PLS_CV_TIES_HIGHEST_INDEX the highest permuted index among the tied points (what a stable argsort gives); no host
 *                             round trip at all; tie_fn is ignored.
 * After the call the scratch header holds, as doubles / int64 bit patterns: [3] n_selected, [4] sum(d) after the last update,
 * [8] the smallest relative gap (max - runner-up) / max seen at any pivot choice (0 if a tie occurred: how close the
 * selection came to depending on round-off), [9] the number of pivots chosen among tied maxima. */
#define PLS_CV_HEADER_DOUBLES 16
enum { PLS_CV_TIES_HIGHEST_INDEX = 0, PLS_CV_TIES_HOST = 1 };
typedef int64_t (*pls_cv_tie_fn)(void* user, const double* d_host, int64_t n, const int64_t* chosen, int n_chosen);
int64_t pls_cv_scratch_doubles(int64_t n, int d, int m);
int pls_cv_select_f64(pls_ctx* ctx, int kernel_id, const double* xp_aug, int64_t n, int d, double kdiag, int m,
                      double jitter, double threshold, int has_threshold, int tie_mode, pls_cv_tie_fn tie_fn, void* tie_user,
                      double* ci, double* di, double* scratch, int64_t* indices_out, int* n_selected_out, void* stream);

/* ---- the selector with the points sharded by rows over several GPUs -------------------------------------------- */
/* Rank r holds rows [n_offset, n_offset + n_local) of the permuted point set, its slice of C ((m-1) x n_local) and of d.
 * Per pivot every rank publishes ONE candidate record (pls_cv_candidate_doubles(d, m) doubles: value, global index,
 * local sum of d, multiplicity, runner-up, the candidate's augmented point and its column of C); the HOST all-gathers the
 * records of all ranks (torch.distributed / NCCL, the only exchange: <= (m + SP + 8) doubles per rank and pivot) and every
 * rank picks the same pivot from them with the single-GPU rules, so the selected indices equal pls_cv_select_f64's.
 *   pls_cv_shard_begin_f64 : d = diag + jitter; candidate for pivot 0
 *   [all-gather]  pls_cv_shard_pick_f64(slot = 0)
 *   for i in 0 .. m-2:  pls_cv_shard_update_f64(iter = i)  (rank-1 update with the published pivot; candidate for pivot i+1)
 *                       [all-gather]  pls_cv_shard_pick_f64(slot = i + 1)
 *   pls_cv_shard_finish    : synchronises, returns how many pivots were chosen
 * Ties (tie_mode as above): with PLS_CV_TIES_HOST a pick that finds the maximum attained more than once (over all ranks)
 * publishes nothing and sets the pause flag; later update / pick calls are no-ops until the host has decided.  The host
 * reads pls_cv_shard_status (synchronises; status4 = {n_selected, stopped, tie pending, slot of the tie}), gathers the
 * ranks' slices of d, decides, and pushes the pivot back on every rank with
 *   pls_cv_shard_force_f64(slot, pivot)  [all-gather]  pls_cv_shard_pick_f64(slot, forced = 1)
 * and resumes with pls_cv_shard_update_f64(iter = slot).
 * indices_out: m int64 GLOBAL positions in the permuted order, pre-filled with the sentinel n_total by the caller.
 * scratch: pls_cv_shard_scratch_doubles(n_local, d, m) doubles, private to the rank. */
int64_t pls_cv_shard_scratch_doubles(int64_t n_local, int d, int m);
int64_t pls_cv_candidate_doubles(int d, int m);
int pls_cv_shard_begin_f64(pls_ctx* ctx, int kernel_id, const double* xa_local, int64_t n_local, int64_t n_offset, int d,
                           double kdiag, int m, double jitter, double* di, double* scratch, double* candidate, void* stream);
int pls_cv_shard_pick_f64(pls_ctx* ctx, const double* candidates, int world, int slot, int d, int m, double threshold,
                          int has_threshold, int tie_mode, int forced, int64_t n_local, int64_t n_offset, double* scratch,
                          int64_t* indices_out, void* stream);
int pls_cv_shard_update_f64(pls_ctx* ctx, int kernel_id, const double* xa_local, int64_t n_local, int64_t n_offset, int d,
                            int iter, int m, double jitter, double* ci, double* di, double* scratch, double* candidate,
                            void* stream);
int pls_cv_shard_force_f64(pls_ctx* ctx, const double* xa_local, int64_t n_local, int64_t n_offset, int d, int m, int slot,
                           int64_t pivot, const double* ci, const double* di, double* scratch, double* candidate, void* stream);
int pls_cv_shard_status(pls_ctx* ctx, const double* scratch, int64_t* status4, void* stream);
int pls_cv_shard_finish(pls_ctx* ctx, const double* scratch, int* n_selected_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLS_B200_H_ */
