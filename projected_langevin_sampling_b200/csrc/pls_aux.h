// Internal declarations of the auxiliary kernels (pls_aux.cu) and the selector (pls_selector.cu).
#pragma once
#include "pls_internal.h"

namespace pls {

constexpr int MAX_D = 26;  // D + 2 <= 28 = 4 * MAX_NKD

struct DimVec {
  double v[MAX_D];
};

struct SmallGemmParams {
  const double* a;
  int64_t lda;
  const double* b;
  int64_t ldb;
  double* c;
  int64_t ldc;
  int64_t rows, j, k;
  // Langevin-update epilogue (orthonormal.py:151-158)
  const double* particles;
  int64_t ldp;
  const double* inv_lambda;
  const double* xi;
  int64_t ldxi;
  double eta;
  int noise_mode;
  int in_place;
  uint64_t seed, step;
  const uint64_t* step_counter;  // optional device counter added to `step` (CUDA-graph replay: the host value is frozen)
  int64_t j_global_offset;
};

cudaError_t launch_prepare_points(int kernel_id, const double* x, int64_t n, int d, int64_t ldx, const DimVec& inv_ls,
                                  const DimVec& centre, double c_extra, int sp, double* out, cudaStream_t stream);
cudaError_t launch_gram_fill(int kernel_id, const double* ra, int64_t nr, const double* ca, int64_t nc, int d, int sp, double* out,
                             int64_t ldo, cudaStream_t stream);
cudaError_t launch_gram(int kernel_id, const double* ra, int64_t nr, const double* ca, int64_t nc, int d, int sp,
                        double* out, int64_t ldo, cudaStream_t stream);
cudaError_t launch_philox_fill(uint64_t seed, uint64_t step, int64_t rows, int64_t j, int64_t joff, double* out,
                               int64_t ldo, cudaStream_t stream);
cudaError_t launch_small_gemm(const SmallGemmParams& p, bool trans_a, bool update, cudaStream_t stream);
cudaError_t launch_reduce_splits(const double* gp, int splits, int64_t rows, int64_t j, int64_t ldg, double* out,
                                 int64_t ldo, cudaStream_t stream);
cudaError_t launch_cost_derivative(const pls_cost& cost, const double* y, const double* f, int64_t ldf, int64_t n,
                                   int64_t j, double* out, int64_t ldo, int sm_count, cudaStream_t stream);
cudaError_t launch_cost_value(const pls_cost& cost, const double* y, const double* f, int64_t ldf, int64_t n, int64_t j,
                              double* partial, cudaStream_t stream);
cudaError_t launch_energy_terms(const double* partial, int64_t tiles, int64_t ldpart, const double* p, int64_t ldp,
                                int64_t m_k, const double* inv_lambda, int64_t j, double* out, cudaStream_t stream);

cudaError_t launch_advance_counter(uint64_t* counter, uint64_t increment, cudaStream_t stream);
cudaError_t launch_lincomb3(int64_t rows, int64_t j, double a, const double* x, int64_t ldx, double b, const double* y, int64_t ldy,
                            double c, const double* z, int64_t ldz, const double* base, int64_t ldb, double* out, int64_t ldo,
                            cudaStream_t stream);
cudaError_t launch_flat_math(int op, const double* a, const double* b, int64_t n, double* out, cudaStream_t stream);
cudaError_t launch_gram_exp(const double* x, int64_t n, int fast, double* out, cudaStream_t stream);

// ConditionalVariance selector (pls_selector.cu)
int64_t cv_scratch_doubles(int64_t n, int d, int m);
cudaError_t run_cv_select(const pls_ctx* ctx, int kernel_id, const double* xp_aug, int64_t n, int d, double kdiag, int m,
                          double jitter, double threshold, int has_threshold, int tie_mode, pls_cv_tie_fn tie_fn, void* tie_user,
                          double* ci, double* di, double* scratch, int64_t* indices_out, int* n_selected_out, cudaStream_t stream);

// row-sharded selector (one candidate record per rank and pivot; the host all-gathers the records between the calls)
int64_t cv_shard_scratch_doubles(int64_t n_local, int d, int m);
int64_t cv_candidate_doubles(int d, int m);
cudaError_t cv_shard_begin(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, double kdiag, int m,
                           double jitter, double* di, double* scratch, double* cand, cudaStream_t stream);
cudaError_t cv_shard_pick(const double* cands, int world, int slot, int d, int m, double threshold, int has_threshold, int tie_mode,
                          int forced, int64_t n_local, int64_t n_offset, double* scratch, int64_t* indices, cudaStream_t stream);
cudaError_t cv_shard_update(int kernel_id, const double* xa, int64_t n_local, int64_t n_offset, int d, int iter, int m,
                            double jitter, double* ci, double* di, double* scratch, double* cand, cudaStream_t stream);
cudaError_t cv_shard_force(const double* xa, int64_t n_local, int64_t n_offset, int d, int m, int slot, int64_t pivot,
                           const double* ci, const double* di, double* scratch, double* cand, cudaStream_t stream);
cudaError_t cv_shard_status(const double* scratch, int64_t* status4, cudaStream_t stream);

}  // namespace pls
