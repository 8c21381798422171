// Posterior variance (diagonal of Sigma) and expected improvement, fused:
//   var_i = scale - sum_k V[k][i]^2          (stationary kernel: k(x,x) = scale)
//   ei_i  = std_i * (z Phi(z) + phi(z)),  z = (mean_i - best_b) / std_i
// Replaces `eval(pts, include_covar=True)` + `covmat.diagonal()` + the EI formula of
// the reference's EI-family acquisitions (src/util/misc_util.py:368-385,
// src/strategies/joint_mei.py:39-48, src/cstrats/profile_cts.py:54-67) without ever
// forming the m x m covariance: O(m n) instead of O(m^2 n).  HBM-bound: V is read once,
// coalesced (thread = candidate, rows of V are contiguous in the candidate index).
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

struct EiArgs {
  const double* V; int n; int ldv; int64_t sV; const int32_t* nv;
  const double* mean; int64_t sMean;
  const double* theta; int64_t sTheta;
  const double* best;  // per problem, may be NULL (variance only)
  int m; const int32_t* mv;
  double* var; int64_t sVar;
  double* ei; int64_t sEi;
  const double* w; int64_t sW;          // optional: mean_out_i = mu0 + sum_k V[k][i] w[k]
  double* mean_out; int64_t sMeanOut;
};

__global__ void __launch_bounds__(128) post_var_ei_kernel(EiArgs a) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= a.m) return;
  const int nv = a.nv ? a.nv[b] : a.n;
  const int mv = a.mv ? a.mv[b] : a.m;
  const double* V = a.V + (size_t)b * a.sV + i;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
  const double* w = a.w ? a.w + (size_t)b * a.sW : nullptr;
  int k = 0;
  for (; k + 3 < nv; k += 4) {
    const double v0 = V[(size_t)k * a.ldv], v1 = V[(size_t)(k + 1) * a.ldv];
    const double v2 = V[(size_t)(k + 2) * a.ldv], v3 = V[(size_t)(k + 3) * a.ldv];
    s0 = fma(v0, v0, s0);
    s1 = fma(v1, v1, s1);
    s2 = fma(v2, v2, s2);
    s3 = fma(v3, v3, s3);
    if (w) {
      d0 = fma(v0, w[k], d0);
      d1 = fma(v1, w[k + 1], d1);
      d2 = fma(v2, w[k + 2], d2);
      d3 = fma(v3, w[k + 3], d3);
    }
  }
  for (; k < nv; ++k) {
    const double v = V[(size_t)k * a.ldv];
    s0 = fma(v, v, s0);
    if (w) d0 = fma(v, w[k], d0);
  }
  if (a.mean_out)
    a.mean_out[(size_t)b * a.sMeanOut + i] =
        (i < mv) ? a.theta[(size_t)b * a.sTheta + TH_MU0] + ((d0 + d1) + (d2 + d3)) : 0.0;
  const double scale = a.theta[(size_t)b * a.sTheta + TH_SCALE];
  const double var = (i < mv) ? scale - ((s0 + s1) + (s2 + s3)) : 1.0;
  if (a.var) a.var[(size_t)b * a.sVar + i] = var;
  if (a.ei) {
    double ei = -CUDART_INF;  // padding never wins an arg-max
    if (i < mv) {
      const double sd = sqrt(var);
      const double z = (a.mean[(size_t)b * a.sMean + i] - a.best[b]) / sd;
      const double cdf = 0.5 * erfc(-z * 0.7071067811865476);
      const double pdf = 0.3989422804014327 * exp(-0.5 * z * z);
      ei = sd * (z * cdf + pdf);
    }
    a.ei[(size_t)b * a.sEi + i] = ei;
  }
}

cudaError_t launch_post_var_ei(const double* V, int n, int ldv, int64_t sV, const int32_t* nv,
                               const double* mean, int64_t sMean, const double* theta,
                               int64_t sTheta, const double* best, int m, const int32_t* mv,
                               double* var, int64_t sVar, double* ei, int64_t sEi, int batch,
                               cudaStream_t st, const double* w, int64_t sW, double* mean_out,
                               int64_t sMeanOut) {
  EiArgs a{V, n, ldv, sV, nv, mean, sMean, theta, sTheta, best, m, mv, var, sVar, ei, sEi,
           w, sW, mean_out, sMeanOut};
  post_var_ei_kernel<<<dim3((m + 127) / 128, batch), 128, 0, st>>>(a);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
