/*
 * libocbo_b200 -- C ABI of the B200-native GP posterior + Thompson-sampling
 * path behind OCBO's Multi-task Thompson Sampling (MTS).
 *
 * The reference (fusion-ml/OCBO) has no FFI: its boundary for this path is the
 * Python GP interface in src/gp/gp_interface.py.  Each entry point below names
 * the reference call it replaces (file:line under /root/reference); the Python
 * binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - matrices are row-major fp64 with an explicit leading dimension (elements)
 *     and a batch stride (elements); `batch` problems are processed per call
 *     (grid.z), a stride of 0 shares the operand between problems;
 *   - dimensions that feed the DMMA tiles (n, m, nrhs) must be multiples of
 *     OCBO_TILE (128); callers pad.  Padding is exact: padded rows/cols of a
 *     Gram matrix are identity, of a cross-covariance zero, so results on the
 *     valid part are unchanged (DESIGN.md "padding");
 *   - theta[b] = { scale, noise_var, mu0, bw_1 .. bw_D } (stride `s_theta`);
 *   - calls are asynchronous on `stream`, never synchronise, never allocate;
 *   - return 0 on success, -k if argument k (1-based) is invalid,
 *     1000 + cudaError_t on a CUDA failure (see ocbo_last_error_string);
 *   - numerical failure of a Cholesky is reported per problem in `info[b]`
 *     (device int32): 0 ok, k>0 = first non-positive / NaN pivot at column k,
 *     LAPACK dpotrf convention, so the host can walk Dragonfly's jitter ladder
 *     (SURVEY.md Appendix A stable_cholesky).
 */
#ifndef OCBO_B200_H
#define OCBO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCBO_TILE 128
#define OCBO_MAX_DIM 16

/* kernel kinds: SE and Matern nu = 1/2, 3/2, 5/2 */
#define OCBO_KERNEL_SE 0
#define OCBO_KERNEL_MATERN12 1
#define OCBO_KERNEL_MATERN32 2
#define OCBO_KERNEL_MATERN52 3

/* ocbo_gram flags */
#define OCBO_GRAM_SYM 1        /* X1 == X2: pad diagonal = 1, symmetric semantics   */
#define OCBO_GRAM_LOWER 2      /* with SYM: only tiles on/below the diagonal written */
#define OCBO_GRAM_ADD_NOISE 4  /* with SYM: add theta.noise_var to valid diagonal    */

typedef void* ocbo_stream_t; /* cudaStream_t */

const char* ocbo_last_error_string(void);
int ocbo_version(void);
/* Kernel launches issued by this process through the library (statistics). */
unsigned long long ocbo_launch_count(void);

/* Workspace (bytes) for the inverse diagonal blocks that ocbo_potrf writes and
 * ocbo_trsm_lower / ocbo_solve_vec read: batch * (n/128) * 128*128 doubles. */
int64_t ocbo_invdiag_elems(int n);

/* K[b][i][j] = scale_b * phi(|| (x1_i - x2_j) / bw_b ||) for i < nv1[b], j < nv2[b];
 * padding per the conventions above.  nv1/nv2 may be NULL (= n1/n2).
 * Replaces gp_core.kernel(A, B)  (src/gp/gp_interface.py:113-117) and the Gram
 * assembly inside Dragonfly's build_posterior / eval. */
int ocbo_gram(int kind, int dim,
              const double* X1, int n1, int64_t s_x1, const int32_t* nv1,
              const double* X2, int n2, int64_t s_x2, const int32_t* nv2,
              const double* theta, int64_t s_theta,
              double* K, int ldk, int64_t s_k,
              int batch, int flags, ocbo_stream_t stream);

/* A[b][i][i] += val[b] for i < nv[b] (jitter ladder step). */
int ocbo_add_diag(double* A, int n, int lda, int64_t s_a, const int32_t* nv,
                  const double* val, int batch, ocbo_stream_t stream);

/* out[b] = max_i A[b][i][i], i < nv[b]  (the ladder's max(diag M)). */
int ocbo_diag_max(const double* A, int n, int lda, int64_t s_a, const int32_t* nv,
                  double* out, int batch, ocbo_stream_t stream);

/* In-place lower Cholesky A = L L^T (row-major, lower triangle referenced and
 * overwritten; the strictly upper part of the diagonal 128-blocks is zeroed,
 * blocks above the diagonal are not touched).  Also writes inv(L_jj) of every
 * diagonal block to invdiag.  info must be zeroed by the caller.
 * Replaces np.linalg.cholesky inside stable_cholesky
 * (GP.build_posterior / draw_gaussian_samples; src/gp/gp_interface.py:78,87-88,132). */
int ocbo_potrf(double* A, int n, int lda, int64_t s_a,
               double* invdiag, int32_t* info, int batch, ocbo_stream_t stream);

/* ocbo_potrf with a device workspace (256-byte aligned, ocbo_potrf_ws_bytes(n, batch) bytes; NULL or
 * too small: same as ocbo_potrf).  With it the rank-k trailing updates of long k run on the 5th-generation
 * tensor cores: the finished panel is split error-free into 7 int8 slices (balanced radix-256 digits) with
 * row-wise exponents, the slice products are formed by tcgen05.mma kind::i8 with exact int32 accumulation in tensor
 * memory and recombined in fp64 in the epilogue (csrc/ozaki.cu).  Slice pairs below 2^-56 of the row maxima
 * are dropped: the update differs from the fp64 one by <= ~1e-14 |P||P|^T element-wise (fp64's own bound
 * for k = 2048 is 2e-13), so factors agree to rounding level but not bit for bit with ocbo_potrf.
 * The row scales are fixed before the factorisation from its diagonal (|L_ik| <= sqrt(A_ii)), so every element of
 * the factor is split once and the slices stay in the workspace (ocbo_sample_tile_argmax_ws can reuse them).
 * The library never allocates: ocbo_potrf_ws_bytes returns 0 when the shape would not use the workspace. */
int64_t ocbo_potrf_ws_bytes(int n, int batch);
int ocbo_potrf_ws(double* A, int n, int lda, int64_t s_a,
                  double* invdiag, int32_t* info, int batch,
                  void* ws, int64_t ws_bytes, ocbo_stream_t stream);

/* ocbo_potrf for identity-padded problems: nv[b] (device, may be NULL) is the valid
 * size of problem b; elimination steps that would only touch the identity padding
 * are skipped (bit-identical result, shorter serial chain for small n). */
int ocbo_potrf_nv(double* A, int n, int lda, int64_t s_a, const int32_t* nv,
                  double* invdiag, int32_t* info, int batch, ocbo_stream_t stream);

/* Device-side walk of stable_cholesky's jitter ladder (Dragonfly
 * draw_gaussian_samples / build_posterior) for a whole batch without a host round
 * trip per rung.  While a ladder is walked info[b] means: 0 attempt pending or
 * succeeded, k>0 last attempt failed at column k, -1 parked (succeeded earlier).
 * One rung p = ocbo_ladder_restore(factor = 10^p) ; ocbo_ladder_advance(p, finish=0) ;
 * ocbo_potrf[_nv].  After the last rung ocbo_ladder_advance(finish=1) maps -1 -> 0.
 *   restore: for problems with info[b] > 0, A[b] <- A0[b] + factor * max_i<nv A0[b][i][i]
 *            on the valid diagonal (tiles on/below the diagonal of 128-blocks);
 *   advance: info > 0 -> rung[b] = p, info[b] = 0;  info == 0 -> info[b] = -1. */
int ocbo_ladder_restore(double* A, const double* A0, int n, int lda, int64_t s_a,
                        int lda0, int64_t s_a0, const int32_t* nv, const int32_t* info,
                        double factor, int batch, ocbo_stream_t stream);
int ocbo_ladder_advance(int32_t* info, int32_t* rung, int p, int finish, int batch,
                        ocbo_stream_t stream);
/* Speculative first rung factored CONCURRENTLY with the plain attempt (on another stream, on a
 * copy A1 = A0 + 10^p max(diag) I with its own invdiag1 / info1): problems whose plain attempt
 * failed (info[b] > 0) while the speculative one succeeded (info1[b] == 0) adopt it --
 * A[b], invdiag[b] <- A1[b], invdiag1[b]; info[b] = 0; rung[b] = p.  Same layout (lda, s_a) for
 * A and A1.  MTS samples at the queried points, so the plain attempt fails on almost every step
 * and the first rung almost always succeeds: the decision's latency is max, not sum. */
int ocbo_ladder_adopt(double* A, const double* A1, int n, int lda, int64_t s_a,
                      double* invdiag, const double* invdiag1,
                      int32_t* info, const int32_t* info1, int32_t* rung, int p, int batch,
                      ocbo_stream_t stream);

/* B <- L^{-1} B  (left, lower, non-transposed), B is n x nrhs row-major.
 * Replaces solve_lower_triangular(L, K(X*, X)^T)  (src/gp/gp_interface.py:116,118;
 * Dragonfly eval 'covar'). */
int ocbo_trsm_lower(const double* L, int n, int ldl, int64_t s_l,
                    const double* invdiag,
                    double* B, int nrhs, int ldb, int64_t s_b,
                    const int32_t* info, int batch, ocbo_stream_t stream);

/* alpha = L^{-T} L^{-1} (y - mu0) on the nv[b] valid rows (0 elsewhere) and
 * lml[b] = -1/2 (y-mu0)^T alpha - sum log L_ii - nv/2 log 2 pi.
 * Either output may be NULL.  Replaces the alpha solve of build_posterior and
 * the LML evaluator behind fit_gp (src/gp/gp_interface.py:87-88,123-125). */
int ocbo_solve_alpha_lml(const double* L, int n, int ldl, int64_t s_l,
                         const double* invdiag,
                         const double* y, int64_t s_y, const int32_t* nv,
                         const double* theta, int64_t s_theta,
                         double* alpha, int64_t s_alpha, double* lml,
                         const int32_t* info, int batch, ocbo_stream_t stream);

/* Same solve with all three outputs optional: alpha, w = L^{-1}(y - mu0) and the LML
 * (w feeds ocbo_post_mean_var_from_v: posterior mean without a second pass over the kernel). */
int ocbo_solve_alpha_w_lml(const double* L, int n, int ldl, int64_t s_l,
                           const double* invdiag,
                           const double* y, int64_t s_y, const int32_t* nv,
                           const double* theta, int64_t s_theta,
                           double* alpha, int64_t s_alpha, double* w, int64_t s_w,
                           double* lml, const int32_t* info, int batch, ocbo_stream_t stream);

/* mean[b][i] = mu0 + sum_j k(xs_i, x_j) alpha_j, i < mv[b] (0 on padding).
 * Replaces the mean half of GP.eval (src/gp/gp_interface.py:69-74). */
int ocbo_post_mean(int kind, int dim,
                   const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                   const double* X, int n, int64_t s_x, const int32_t* nv,
                   const double* theta, int64_t s_theta,
                   const double* alpha, int64_t s_alpha,
                   double* mean, int64_t s_mean, int batch, ocbo_stream_t stream);

/* Sigma = K(Xs, Xs) - V^T V with V = L^{-1} K(X, Xs) (n x m, from
 * ocbo_trsm_lower).  lower_only != 0 writes tiles on/below the diagonal only
 * (diagonal tiles full); padding rows/cols of Sigma are identity.
 * Replaces the covariance half of GP.eval(X, 'covar') and the in-tree
 * get_pt_relations (src/gp/gp_interface.py:72,112-119). */
int ocbo_post_cov(int kind, int dim,
                  const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                  const double* V, int n, int ldv, int64_t s_v,
                  const double* theta, int64_t s_theta,
                  double* Sigma, int lds, int64_t s_sigma,
                  int lower_only, int batch, ocbo_stream_t stream);

/* ocbo_post_cov with a device workspace (256-byte aligned, ocbo_post_cov_ws_bytes(m, n, batch) bytes; NULL, too
 * small, lower_only == 0 or a small shape: same as ocbo_post_cov).  With it V^T V runs on the int8 tensor cores
 * (see ocbo_potrf_ws): K(Xs, Xs) is written first, V is split column-wise into int8 slices and the fused update
 * subtracts the product; V's padded columns must be zero (they are after ocbo_gram / ocbo_trsm_lower). */
int64_t ocbo_post_cov_ws_bytes(int m, int n, int batch);
int ocbo_post_cov_ws(int kind, int dim, const double* Xs, int m, int64_t s_xs, const int32_t* mv,
                     const double* V, int n, int ldv, int64_t s_v, const double* theta,
                     int64_t s_theta, double* Sigma, int lds, int64_t s_sigma, int lower_only,
                     int batch, void* ws, int64_t ws_bytes, ocbo_stream_t stream);

/* Cross form of the above for two point sets (get_pt_relations, :112-119):
 * C = K(A, B) - V1^T V2, V1 (n x ma), V2 (n x mb). */
int ocbo_post_cross_cov(int kind, int dim,
                        const double* Xa, int ma, int64_t s_xa, const int32_t* mav,
                        const double* Xb, int mb, int64_t s_xb, const int32_t* mbv,
                        const double* V1, int ldv1, int64_t s_v1,
                        const double* V2, int ldv2, int64_t s_v2, int n,
                        const double* theta, int64_t s_theta,
                        double* C, int ldc, int64_t s_c,
                        int batch, ocbo_stream_t stream);

/* Forward half of the above: w = L^{-1}(y - mu0) and the LML (= -1/2 w^T w - ...).
 * With V = L^{-1} K(X, Xs) already at hand, the posterior mean is mu0 + V^T w
 * (ocbo_post_mean_var_from_v): no back-substitution, no second pass over the kernel. */
int ocbo_solve_forward_lml(const double* L, int n, int ldl, int64_t s_l,
                           const double* invdiag,
                           const double* y, int64_t s_y, const int32_t* nv,
                           const double* theta, int64_t s_theta,
                           double* w, int64_t s_w, double* lml,
                           const int32_t* info, int batch, ocbo_stream_t stream);

/* mean[b][i] = mu0 + sum_k V[b][k][i] w[b][k],  var[b][i] = scale - sum_k V[b][k][i]^2
 * in one HBM-bound pass over V (either output may be NULL).  Same values as
 * GP.eval(X, 'std')'s mean / std^2 (src/gp/gp_interface.py:69-74). */
int ocbo_post_mean_var_from_v(const double* V, int n, int ldv, int64_t s_v, const int32_t* nv,
                              const double* w, int64_t s_w,
                              const double* theta, int64_t s_theta,
                              int m, const int32_t* mv,
                              double* mean, int64_t s_mean, double* var, int64_t s_var,
                              int batch, ocbo_stream_t stream);

/* Diagonal of Sigma and expected improvement, fused, without forming Sigma:
 *   var[b][i] = scale_b - sum_k V[b][k][i]^2,
 *   ei[b][i]  = std (z Phi(z) + phi(z)),  z = (mean[b][i] - best[b]) / std   (-inf on padding).
 * var or ei may be NULL.  Replaces eval(pts, True) + covmat.diagonal() + the EI formula of
 * the EI-family acquisitions (src/util/misc_util.py:368-385, src/strategies/joint_mei.py:39-48,
 * src/cstrats/profile_cts.py:54-67) and eval(x, 'std') (misc_util.py:321,326). */
int ocbo_post_var_ei(const double* V, int n, int ldv, int64_t s_v, const int32_t* nv,
                     const double* mean, int64_t s_mean,
                     const double* theta, int64_t s_theta, const double* best,
                     int m, const int32_t* mv, double* var, int64_t s_var,
                     double* ei, int64_t s_ei, int batch, ocbo_stream_t stream);

/* Copy the lower triangle onto the upper one (full symmetric covar for eval). */
int ocbo_symmetrize_lower(double* A, int n, int lda, int64_t s_a, int batch,
                          ocbo_stream_t stream);

/* Z[b] (m x S row-major, ld = ldz) <- standard normals from Philox4x32-10
 * keyed by (seed, trial_ids[b]); layout contract in oracle/philox.py.
 * Replaces np.random.normal(size=(m, S)) in draw_gaussian_samples. */
int ocbo_philox_normal(double* Z, int m, int S, int ldz, int64_t s_z,
                       uint64_t seed, const int32_t* trial_ids, int batch,
                       ocbo_stream_t stream);

/* F = mean + L Z  (m x S; S a multiple of 128 uses the DMMA path, S <= 8 the
 * HBM-bound row-dot path).  Replaces (L.dot(U)).T + mu in
 * draw_gaussian_samples (src/gp/gp_interface.py:76-79). */
int ocbo_sample(const double* mean, int64_t s_mean,
                const double* L, int m, int ldl, int64_t s_l,
                const double* Z, int S, int ldz, int64_t s_z,
                double* F, int ldf, int64_t s_f,
                const int32_t* info, int batch, ocbo_stream_t stream);

/* Fused sampling tail: per 128-row tile t of F = mean + L Z and per sample column s,
 *   tile_max[b][t][s] = max_i F[b][128 t + i][s],  tile_arg[b][t][s] = first arg-max i in [0,128)
 * (np.max / np.argmax semantics, NaN included), WITHOUT writing F: the m x S sample matrix of
 * joint_mts.py:42 never exists in memory.  With task segments of exactly 128 rows (the sweep:
 * A = 128 actions per task) tile_max / tile_arg ARE the seg_max / seg_arg of ocbo_joint_pick.
 * m a multiple of 128, S a multiple of 32; values bit-identical to ocbo_sample followed by
 * ocbo_segment_argmax.  Replaces draw_sample + np.max/np.argmax (src/gp/gp_interface.py:76-79,
 * src/strategies/joint_mts.py:42-53). */
int ocbo_sample_tile_argmax(const double* mean, int64_t s_mean,
                            const double* L, int m, int ldl, int64_t s_l,
                            const double* Z, int S, int ldz, int64_t s_z,
                            double* tile_max, int32_t* tile_arg,
                            const int32_t* info, int batch, ocbo_stream_t stream);

/* ocbo_sample_tile_argmax with a device workspace (256-byte aligned, ocbo_sample_ws_bytes(m, S, batch) bytes; NULL,
 * too small or a small shape: same as ocbo_sample_tile_argmax).  With it L (lower triangular: only the blocks up to
 * the diagonal) and Z^T are split into int8 slices and mean + L Z runs on the int8 tensor cores (see ocbo_potrf_ws);
 * max / first arg-max per (128-row tile, column) come out of the tensor-memory epilogue, same np.argmax rule.
 * l_in_ws != 0: L was factored by ocbo_potrf_ws with THIS workspace and nothing has written to it since -- its
 * slices (and row scales) are still at the head of the workspace and are not computed again. */
int64_t ocbo_sample_ws_bytes(int m, int S, int batch);
int ocbo_sample_tile_argmax_ws(const double* mean, int64_t s_mean, const double* L, int m, int ldl,
                               int64_t s_l, const double* Z, int S, int ldz, int64_t s_z,
                               double* tile_max, int32_t* tile_arg, const int32_t* info, int batch,
                               void* ws, int64_t ws_bytes, int l_in_ws, ocbo_stream_t stream);

/* Segmented max / first-index argmax per sample column:
 * seg_ptr[b][0..nseg] row offsets; out_max/out_arg are [batch][nseg][S]
 * (arg relative to the segment start; empty segment -> -inf / -1).
 * Replaces np.max / np.argmax in src/strategies/mts.py:37-38,
 * joint_mts.py:46-53, src/cstrats/profile_cts.py:101-104. */
int ocbo_segment_argmax(const double* F, int S, int ldf, int64_t s_f,
                        const int32_t* seg_ptr, int nseg,
                        double* out_max, int32_t* out_arg,
                        int batch, ocbo_stream_t stream);

/* Joint-MTS pick from the segment reductions (joint_mts.py:43-57):
 * segments 0..T-1 are the old points per task (pass n_old_seg = 0 for the
 * grid-only sweep: old best = 0), segments n_old_seg.. are the T new-point
 * segments.  Per (b, s): task = first argmax_t (max new_t - max old_t),
 * action = arg within new_t.  Outputs [batch][S]. */
int ocbo_joint_pick(const double* seg_max, const int32_t* seg_arg, int nseg,
                    int n_old_seg, int T, int S,
                    int32_t* out_task, int32_t* out_action, double* out_gain,
                    int batch, ocbo_stream_t stream);

/* Knowledge gradient of G x C sets of E lines (REVI):
 *   kg[c][g] = E[max_i (a_gi + b_cgi Z)] - max_i a_gi,  Z ~ N(0,1),
 *   a_gi = judge_means[g * ld_means + i],  b_cgi = inter[c * ld_inter + g * E + i] / sqrt(noise + cand_var[c]).
 * Sort by (slope, intercept), upper envelope with a stack, segment sum in order -- the
 * reference's knowledge_gradient (src/util/misc_util.py:180-212) as called from REVI's loops
 * (src/strategies/corr_opt.py:103-119, src/cstrats/postmax_cts.py:147-160).  2 <= E <= 1024. */
int ocbo_kg(const double* judge_means, int ld_means, const double* inter, int64_t ld_inter,
            const double* cand_var, double noise, int C, int G, int E, double* kg,
            ocbo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * OCBO_BENCH section -- measurement helpers for bench.py, NOT part of the drop-in path and
 * exempt from the "never synchronise, never allocate" contract above: they create CUDA events
 * and wait for the stream.  No product code calls them.
 * ------------------------------------------------------------------------------------------ */
#ifndef OCBO_NO_BENCH

/* ocbo_potrf with CUDA events around every launch, summed per kernel class into
 * ms_host[4] = { diagonal-block, panel-solve, in-panel rank-128 update,
 * outer rank-k trailing update } milliseconds.
 * SYNCHRONISES the stream and creates / destroys CUDA events. */
int ocbo_potrf_profile(double* A, int n, int lda, int64_t s_a,
                       double* invdiag, int32_t* info, int batch, ocbo_stream_t stream,
                       float* ms_host);

/* ocbo_potrf_ws instrumented the same way: ms_host[7] = { diagonal-block, panel-solve, in-panel fp64 update,
 * outer fp64 update, int8 split, int8-tensor-core outer update, int8-tensor-core middle update }. */
int ocbo_potrf_ws_profile(double* A, int n, int lda, int64_t s_a,
                          double* invdiag, int32_t* info, int batch,
                          void* ws, int64_t ws_bytes, ocbo_stream_t stream, float* ms_host);

/* FP64 throughput probes used by bench.py to measure the DMMA / DFMA rate the
 * roofline is quoted against (MEASURED_PEAKS.json has no FP64 row).
 * Each thread block runs `iters` dependent-free DMMA.8x8x4 / DFMA chains. */
int ocbo_probe_dmma(double* sink, int blocks, int iters, ocbo_stream_t stream);
int ocbo_probe_dfma(double* sink, int blocks, int iters, ocbo_stream_t stream);
/* SE kernel values scale * exp(-d2 / 2) of n scaled squared distances, twice: `lockstep` through the branch-free
 * exp chain the covariance epilogue of the int8 kernel evaluates eight at a time (csrc/common.cuh: exp_core), `plain`
 * through libdevice's exp, one call per element, like every other Gram kernel.  The two must be bit-identical
 * (tests/test_gpu_ozaki.py); replaces nothing in the reference. */
int ocbo_probe_kern_se(const double* d2, double scale, int64_t n, double* lockstep, double* plain,
                       ocbo_stream_t stream);
/* Plain C = A B^T DGEMM (n x n x n, multiples of 128) through the library's
 * DMMA tile core -- the GEMM-shaped roofline reference for the core itself. */
int ocbo_probe_dgemm_nt(const double* A, const double* B, double* C, int n,
                        ocbo_stream_t stream);

/* tcgen05 probe: C (m x n, int32, ld = ldc) = A (m x k, int8, ld = lda) B^T (B: n x k, int8, ld = ldb) on the
 * 5th-generation tensor cores (tcgen05.mma kind::i8, TMA operands, TMEM accumulator) -- the building block of an
 * Ozaki-split fp64 update; m, n, k multiples of 128, lda / ldb multiples of 16. */
int ocbo_probe_igemm_nt(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C, int64_t ldc,
                        int m, int n, int k, ocbo_stream_t stream);

/* Ozaki probe: C (rows x rows, fp64) -= P P^T with P rows x kw fp64, computed on the int8 tensor cores from 8
 * error-free int8 slices of P (csrc/ozaki.cu); tri != 0: only the 128 x 64 tiles on and below the diagonal.
 * ws: device workspace of ocbo_ozaki_ws_bytes(rows, kw, 1) bytes, 256-byte aligned. */
int64_t ocbo_ozaki_ws_bytes(int rows, int kw, int batch);
int ocbo_ozaki_slice_pairs(void);  /* int8 slice-pair products issued per fp64 product (28) */
int ocbo_probe_ozaki_update(const double* P, int ldp, int rows, int kw, double* C, int ldc, void* ws,
                            int64_t ws_bytes, int tri, ocbo_stream_t stream);

#endif /* OCBO_NO_BENCH */

#ifdef __cplusplus
}
#endif
#endif /* OCBO_B200_H */
