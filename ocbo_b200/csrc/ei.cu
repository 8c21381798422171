// Posterior variance (diagonal of Sigma) and expected improvement, fused:
//   var_i = scale - sum_k V[k][i]^2          (stationary kernel: k(x,x) = scale)
//   ei_i  = std_i * (z Phi(z) + phi(z)),  z = (mean_i - best_b) / std_i
// Replaces `eval(pts, include_covar=True)` + `covmat.diagonal()` + the EI formula of
// the reference's EI-family acquisitions (src/util/misc_util.py:368-385,
// src/strategies/joint_mei.py:39-48, src/cstrats/profile_cts.py:54-67) without ever
// forming the m x m covariance: O(m n) instead of O(m^2 n).  HBM-bound: V is read once,
// coalesced (lane = candidate, rows of V are contiguous in the candidate index).
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

// One CTA = 32 candidates x 16 row slices: warp w walks rows k = w, w+16, ... of V (each row
// access is one 256-byte line per warp), 8 loads in flight per thread; the 16 partial sums are
// combined through shared memory in a fixed order, so results do not depend on scheduling.
// (The first version used one thread per candidate and one CTA per 128 candidates: 64 CTAs
// for the sweep's m = 8192 and 354 GB/s.  8 slices: 2.65 TB/s with 256 CTAs of 8 warps on 148
// SMs -- 16-32 KB in flight per SM, under Little's law for HBM3e; 16 slices: 3.33 TB/s (20.2 us
// for the sweep's 67 MB, ncu); 32 slices: 23.0 us.)
constexpr int EI_COLS = 32, EI_SLICES = 16;

__global__ void __launch_bounds__(EI_COLS * EI_SLICES) post_var_ei_kernel(EiArgs a) {
  __shared__ double ps[EI_SLICES][EI_COLS], pd[EI_SLICES][EI_COLS];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * EI_COLS + lane;
  const int nv = a.nv ? a.nv[b] : a.n;
  const int mv = a.mv ? a.mv[b] : a.m;
  const bool in = i < a.m;
  const double* V = a.V + (size_t)b * a.sV + (in ? i : 0);
  const double* w = a.w ? a.w + (size_t)b * a.sW : nullptr;
  double s = 0.0, d = 0.0;
  int k = slice;
  for (; k + 7 * EI_SLICES < nv; k += 8 * EI_SLICES) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = in ? V[(size_t)(k + u * EI_SLICES) * a.ldv] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = fma(v[u], v[u], s);
      if (w) d = fma(v[u], w[k + u * EI_SLICES], d);
    }
  }
  for (; k < nv; k += EI_SLICES) {
    const double v = in ? V[(size_t)k * a.ldv] : 0.0;
    s = fma(v, v, s);
    if (w) d = fma(v, w[k], d);
  }
  ps[slice][lane] = s;
  pd[slice][lane] = d;
  __syncthreads();
  if (slice != 0 || !in) return;
#pragma unroll
  for (int u = 1; u < EI_SLICES; ++u) {
    s += ps[u][lane];
    d += pd[u][lane];
  }
  if (a.mean_out)
    a.mean_out[(size_t)b * a.sMeanOut + i] =
        (i < mv) ? a.theta[(size_t)b * a.sTheta + TH_MU0] + d : 0.0;
  const double scale = a.theta[(size_t)b * a.sTheta + TH_SCALE];
  const double var = (i < mv) ? scale - s : 1.0;
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
  post_var_ei_kernel<<<dim3((m + EI_COLS - 1) / EI_COLS, batch), EI_COLS * EI_SLICES, 0, st>>>(a);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
