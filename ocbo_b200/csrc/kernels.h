// Internal launch interface between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocbo {

struct NtParams {  // C (=|-=) P Q^T ; all pointers pre-offset to the region origin
  const double* P; int ldp; int64_t sP;
  const double* Q; int ldq; int64_t sQ;
  double* C; int ldc; int64_t sC;
  int ntc;         // column tiles (of the kernel's BN) of a rectangular region
  int ktiles;      // k extent / 16
  int row0, col0;  // global element coordinates of the region origin (diagonal test)
  int tri;         // grid.x enumerates the lower-triangular tiles of a square region
  int lower_skip;  // rectangular region: skip tiles entirely above the diagonal
  int mode;        // 0: C = acc, 1: C -= acc
  const int32_t* info;
  int tma_free;    // TMA main loop without the per-k-tile CTA barrier (set by the launcher)
};

struct NnParams {  // C (=|-=|= rowvec +) P R
  const double* P; int ldp; int64_t sP;
  const double* R; int ldr; int64_t sR;
  double* C; int ldc; int64_t sC;
  const double* rowvec; int64_t sRowvec;
  int ktiles;
  int ntr;           // row tiles of the product (pair_rows)
  int tri_k;         // k stops at the diagonal block of the row tile (L lower)
  int pair_rows;     // CTA y computes row tiles y and ntr-1-y (balanced triangular product)
  int mode;          // 0 set, 1 sub, 2 rowvec + acc
  const int32_t* info;
  // fused sampling tail (sample_argmax_kernel): per (row tile, column) max / first arg-max
  double* tile_max; int32_t* tile_arg; int out_ld;
  int tma_free;
};

struct CovParams {  // C = K(Xa, Xb) - V1^T V2
  const double* Xa; int64_t sXa; int ma; const int32_t* mav;
  const double* Xb; int64_t sXb; int mb; const int32_t* mbv;
  const double* V1; int ldv1; int64_t sV1;
  const double* V2; int ldv2; int64_t sV2;
  const double* theta; int64_t sTheta;
  double* C; int ldc; int64_t sC;
  int ntc, ktiles, dim, kind, sym, tri;
  int tma_free;
};

// panel = 0: 128 x 64 tiles (updates); panel = 1: 64 x 128 tiles (in-place panel solve)
cudaError_t launch_gemm_nt(const NtParams& p, int ntiles, int batch, cudaStream_t st, int panel = 0);
cudaError_t launch_gemm_nn(const NnParams& p, int grid_rows, int ntc, int batch, cudaStream_t st);
cudaError_t launch_postcov(const CovParams& p, int ntiles, int batch, cudaStream_t st);
cudaError_t launch_sample_argmax(const NnParams& p, int ntc, int batch, cudaStream_t st);

cudaError_t launch_gram(int kind, int dim, const double* X1, int n1, int64_t sX1, const int32_t* nv1,
                        const double* X2, int n2, int64_t sX2, const int32_t* nv2,
                        const double* theta, int64_t sTheta, double* K, int ldk, int64_t sK,
                        int batch, int flags, cudaStream_t st);
cudaError_t launch_add_diag(double* A, int n, int lda, int64_t sA, const int32_t* nv,
                            const double* val, int batch, cudaStream_t st);
cudaError_t launch_diag_max(const double* A, int n, int lda, int64_t sA, const int32_t* nv,
                            double* out, int batch, cudaStream_t st);
cudaError_t launch_symmetrize(double* A, int n, int lda, int64_t sA, int batch, cudaStream_t st);

// factor the 128x128 diagonal block at (j0, j0) of every problem, write inv(L_jj)
// (nv: optional per-problem valid size -- elimination steps on identity padding are skipped)
cudaError_t launch_potrf_diag(double* A, int lda, int64_t sA, int j0, double* invdiag,
                              int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                              const int32_t* nv = nullptr);
cudaError_t launch_post_var_ei(const double* V, int n, int ldv, int64_t sV, const int32_t* nv,
                               const double* mean, int64_t sMean, const double* theta,
                               int64_t sTheta, const double* best, int m, const int32_t* mv,
                               double* var, int64_t sVar, double* ei, int64_t sEi, int batch,
                               cudaStream_t st, const double* w = nullptr, int64_t sW = 0,
                               double* mean_out = nullptr, int64_t sMeanOut = 0);
cudaError_t launch_potrf_diag_reg(double* A, int lda, int64_t sA, int j0, double* invdiag,
                                  int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                                  const int32_t* nv = nullptr);
cudaError_t launch_potrf_diag_blk(double* A, int lda, int64_t sA, int j0, double* invdiag,
                                  int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                                  const int32_t* nv = nullptr);
// jitter ladder on the device (ladder.cu)
cudaError_t launch_ladder_restore(double* A, const double* A0, int n, int lda, int64_t sA, int lda0,
                                  int64_t sA0, const int32_t* nv, const int32_t* info,
                                  double factor, int batch, cudaStream_t st);
cudaError_t launch_ladder_advance(int32_t* info, int32_t* rung, int p, int finish, int batch,
                                  cudaStream_t st);
cudaError_t launch_ladder_adopt(double* A, const double* A1, int n, int lda, int64_t sA, double* invd,
                                const double* invd1, int64_t sInv, int32_t* info,
                                const int32_t* info1, int32_t* rung, int p, int batch,
                                cudaStream_t st);
cudaError_t launch_solve_alpha_lml(const double* L, int n, int ldl, int64_t sL,
                                   const double* invdiag, int64_t sInv, const double* y, int64_t sY,
                                   const int32_t* nv, const double* theta, int64_t sTheta,
                                   double* alpha, int64_t sAlpha, double* lml,
                                   const int32_t* info, int batch, cudaStream_t st,
                                   double* w = nullptr, int64_t sW = 0);
cudaError_t launch_post_mean(int kind, int dim, const double* Xs, int m, int64_t sXs,
                             const int32_t* mv, const double* X, int n, int64_t sX,
                             const int32_t* nv, const double* theta, int64_t sTheta,
                             const double* alpha, int64_t sAlpha, double* mean, int64_t sMean,
                             int batch, cudaStream_t st);

cudaError_t launch_kg(const double* means, int ldm, const double* inter, int64_t ldi,
                      const double* cand_var, double noise, int C, int G, int E, double* out,
                      cudaStream_t st);
cudaError_t launch_philox_normal(double* Z, int m, int S, int ldz, int64_t sZ, uint64_t seed,
                                 const int32_t* trial_ids, int batch, cudaStream_t st);
cudaError_t launch_sample_rowdot(const double* mean, int64_t sMean, const double* L, int m, int ldl,
                                 int64_t sL, const double* Z, int S, int ldz, int64_t sZ,
                                 double* F, int ldf, int64_t sF, const int32_t* info, int batch,
                                 cudaStream_t st);
cudaError_t launch_segment_argmax(const double* F, int S, int ldf, int64_t sF,
                                  const int32_t* seg_ptr, int nseg, double* out_max,
                                  int32_t* out_arg, int batch, cudaStream_t st);
cudaError_t launch_joint_pick(const double* seg_max, const int32_t* seg_arg, int nseg,
                              int n_old_seg, int T, int S, int32_t* out_task, int32_t* out_action,
                              double* out_gain, int batch, cudaStream_t st);
cudaError_t launch_probe_dmma(double* sink, int blocks, int iters, cudaStream_t st);
cudaError_t launch_probe_dfma(double* sink, int blocks, int iters, cudaStream_t st);
cudaError_t launch_probe_kern_se(const double* d2, double scale, int64_t n, double* lockstep, double* plain,
                                 cudaStream_t st);
// fp64 rank-k updates through the int8 tensor cores (ozaki.cu): split a panel into int8 slices, then
// C (region) -= P Q^T from the split workspace
size_t ozaki_ws_bytes(int rows, int kw, int batch);
int ozaki_slice_pairs();  // int8 products issued per fp64 product
cudaError_t launch_ozaki_split(const double* P, int ldp, int64_t sP, int rows, int kw, void* ws, const int32_t* info,
                               int batch, cudaStream_t st, int transposed = 0, int tri = 0);
size_t ozaki_sample_ws_bytes(int m, int S, int batch);
cudaError_t launch_ozaki_sample(const double* mean, int64_t sMean, const double* L, int m, int ldl, int64_t sL,
                                const double* Z, int S, int ldz, int64_t sZ, double* tile_max, int32_t* tile_arg,
                                const int32_t* info, int batch, void* ws, cudaStream_t st, int l_in_ws = 0);
// factor-wide workspace (row scales fixed from the diagonal before the factorisation; every element of L split once)
cudaError_t launch_ozaki_diag_scale(const double* A, int lda, int64_t sA, int n, void* ws, const int32_t* info, int batch,
                                    cudaStream_t st);
cudaError_t launch_ozaki_split_fixed(const double* P, int ldp, int64_t sP, int rows, int kw, int row0, int kb0, int n,
                                     void* ws, const int32_t* info, int batch, cudaStream_t st);
cudaError_t launch_ozaki_update_range(const void* ws, int n, int kb0, int nkb, int rowA0, int rowB0, double* C, int ldc,
                                      int64_t sC, int R, int ntc, int tri, int lower_skip, int row0, int col0,
                                      const int32_t* info, int batch, cudaStream_t st);
struct OzGram {  // C = K(X, X) - P Q^T: points [rows][dim] per problem, kernel kind, theta
  const double* X; int64_t sX; int dim, kind; const double* theta; int64_t sTheta;
};
int ozaki_gram_max_dim();
cudaError_t launch_ozaki_update(const void* ws, int rows, int kw, int rowA0, int rowB0, double* C, int ldc, int64_t sC,
                                int R, int ntc, int tri, int lower_skip, int row0, int col0, const int32_t* info,
                                int batch, cudaStream_t st, const OzGram* gram = nullptr);
cudaError_t launch_probe_igemm(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C, int64_t ldc,
                               int m, int n, int k, cudaStream_t st);

}  // namespace ocbo

// process-wide count of kernel launches issued by the library (bench.py's
// "gpu_launches"); statistics only, no behaviour depends on it
namespace ocbo {
extern unsigned long long g_launch_count;
inline cudaError_t counted(cudaError_t e) {
  __atomic_fetch_add(&g_launch_count, 1ULL, __ATOMIC_RELAXED);
  return e;
}
}  // namespace ocbo
