// Gram / cross-covariance assembly and the fused posterior-mean GEMV.
// HBM-bound pieces of the path: one pass, 16-byte vector stores, points staged
// (pre-scaled by 1/bandwidth) in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

constexpr int GT = 64;  // Gram tile edge

struct GramArgs {
  const double* X1; int n1; int64_t sX1; const int32_t* nv1;
  const double* X2; int n2; int64_t sX2; const int32_t* nv2;
  const double* theta; int64_t sTheta;
  double* K; int ldk; int64_t sK;
  int kind, dim, flags;
};

// KIND is a template parameter and every kernel value of the thread's 8 x 2 patch is evaluated
// unconditionally (padding is selected afterwards): 16 independent exp() chains per thread in
// straight-line code.  With the kind tested at run time inside the loop the compiler kept the
// evaluations serial and the kernel sat at 20 % FP64-pipe utilisation (85 us for 1024 x 8192).
template <int KIND>
__global__ void __launch_bounds__(256) gram_kernel(GramArgs a) {
  __shared__ double x1s[MAXD][GT];
  __shared__ __align__(16) double x2s[MAXD][GT];
  __shared__ double n1s[GT];
  __shared__ __align__(16) double n2s[GT];
  const int b = blockIdx.z, tj = blockIdx.x, ti = blockIdx.y;
  const bool sym = a.flags & OCBO_GRAM_SYM;
  if (sym && (a.flags & OCBO_GRAM_LOWER)) {
    // keep everything a 128-blocked lower-triangular consumer may touch
    if (tj * GT > ((ti * GT) / TILE) * TILE + TILE - 1) return;
  }
  const int D = a.dim;
  const double* th = a.theta + (size_t)b * a.sTheta;
  const double scale = th[TH_SCALE];
  const int nv1 = a.nv1 ? a.nv1[b] : a.n1;
  const int nv2 = a.nv2 ? a.nv2[b] : a.n2;
  const int row0 = ti * GT, col0 = tj * GT;
  const double* X1 = a.X1 + (size_t)b * a.sX1;
  const double* X2 = a.X2 + (size_t)b * a.sX2;
  for (int e = threadIdx.x; e < GT * D; e += 256) {
    const int r = e / D, d = e % D;
    const double bw = th[TH_BW + d];
    x1s[d][r] = (row0 + r < nv1) ? X1[(size_t)(row0 + r) * D + d] / bw : 0.0;
    x2s[d][r] = (col0 + r < nv2) ? X2[(size_t)(col0 + r) * D + d] / bw : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 2 * GT) {  // squared norms of the scaled points, numpy's summation order
    const int r = threadIdx.x % GT;
    double (*xs)[GT] = threadIdx.x < GT ? x1s : x2s;
    double s = 0.0;
    for (int d = 0; d < D; ++d) s = sqnorm_np(s, xs[d][r]);
    (threadIdx.x < GT ? n1s : n2s)[r] = s;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const double2 nc = *reinterpret_cast<const double2*>(&n2s[2 * tx]);
  double xc0[MAXD], xc1[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) {
    if (d < D) {
      const double2 v = *reinterpret_cast<const double2*>(&x2s[d][2 * tx]);
      xc0[d] = v.x;
      xc1[d] = v.y;
    }
  }
  const double noise = (a.flags & OCBO_GRAM_ADD_NOISE) ? th[TH_NOISE] : 0.0;
  double* K = a.K + (size_t)b * a.sK;
  const int gj = col0 + 2 * tx;
  double kv[GT / 8][2];
#pragma unroll
  for (int k = 0; k < GT / 8; ++k) {
    const int r = ty + 8 * k;
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
      if (d < D) {
        const double xr = x1s[d][r];
        d0 = fma(xr, xc0[d], d0);
        d1 = fma(xr, xc1[d], d1);
      }
    }
    kv[k][0] = dist2_np(n1s[r], nc.x, d0);
    kv[k][1] = dist2_np(n1s[r], nc.y, d1);
  }
  if constexpr (KIND == OCBO_KERNEL_SE) {  // the exp chains in lockstep (common.cuh: same bits as kern_eval)
#pragma unroll
    for (int k = 0; k < GT / 8; k += 4) kern_eval_se<8>(scale, &kv[k][0]);
  } else {
#pragma unroll
    for (int k = 0; k < GT / 8; ++k) {
      kv[k][0] = kern_eval(KIND, scale, kv[k][0]);
      kv[k][1] = kern_eval(KIND, scale, kv[k][1]);
    }
  }
#pragma unroll
  for (int k = 0; k < GT / 8; ++k) {
    const int gi = row0 + ty + 8 * k;
    const bool vr = gi < nv1;
    double o0 = (vr && gj < nv2) ? kv[k][0] : 0.0;
    double o1 = (vr && gj + 1 < nv2) ? kv[k][1] : 0.0;
    if (sym && gi == gj) o0 = (vr && gj < nv2) ? o0 + noise : 1.0;
    if (sym && gi == gj + 1) o1 = (vr && gj + 1 < nv2) ? o1 + noise : 1.0;
    *reinterpret_cast<double2*>(K + (size_t)gi * a.ldk + gj) = make_double2(o0, o1);
  }
}

cudaError_t launch_gram(int kind, int dim, const double* X1, int n1, int64_t sX1, const int32_t* nv1,
                        const double* X2, int n2, int64_t sX2, const int32_t* nv2,
                        const double* theta, int64_t sTheta, double* K, int ldk, int64_t sK,
                        int batch, int flags, cudaStream_t st) {
  GramArgs a{X1, n1, sX1, nv1, X2, n2, sX2, nv2, theta, sTheta, K, ldk, sK, kind, dim, flags};
  dim3 grid(n2 / GT, n1 / GT, batch);
  switch (kind) {
    case OCBO_KERNEL_SE: gram_kernel<OCBO_KERNEL_SE><<<grid, 256, 0, st>>>(a); break;
    case OCBO_KERNEL_MATERN12: gram_kernel<OCBO_KERNEL_MATERN12><<<grid, 256, 0, st>>>(a); break;
    case OCBO_KERNEL_MATERN32: gram_kernel<OCBO_KERNEL_MATERN32><<<grid, 256, 0, st>>>(a); break;
    default: gram_kernel<OCBO_KERNEL_MATERN52><<<grid, 256, 0, st>>>(a); break;
  }
  return counted(cudaGetLastError());
}

// --------------------------------------------------------------------------
__global__ void add_diag_kernel(double* A, int n, int lda, int64_t sA, const int32_t* nv,
                                const double* val) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lim = nv ? nv[b] : n;
  if (i < lim) A[(size_t)b * sA + (size_t)i * lda + i] += val[b];
}

cudaError_t launch_add_diag(double* A, int n, int lda, int64_t sA, const int32_t* nv,
                            const double* val, int batch, cudaStream_t st) {
  add_diag_kernel<<<dim3((n + 255) / 256, batch), 256, 0, st>>>(A, n, lda, sA, nv, val);
  return counted(cudaGetLastError());
}

__global__ void diag_max_kernel(const double* A, int n, int lda, int64_t sA, const int32_t* nv,
                                double* out) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  const int lim = nv ? nv[b] : n;
  double m = -CUDART_INF;
  for (int i = threadIdx.x; i < lim; i += blockDim.x)
    m = fmax(m, A[(size_t)b * sA + (size_t)i * lda + i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -CUDART_INF;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) out[b] = m;
  }
}

cudaError_t launch_diag_max(const double* A, int n, int lda, int64_t sA, const int32_t* nv,
                            double* out, int batch, cudaStream_t st) {
  diag_max_kernel<<<batch, 256, 0, st>>>(A, n, lda, sA, nv, out);
  return counted(cudaGetLastError());
}

// mirror the lower triangle onto the upper one, 32x32 tiles through smem
__global__ void symmetrize_kernel(double* A, int n, int lda, int64_t sA) {
  __shared__ double t[32][33];
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (tj > ti) return;
  double* M = A + (size_t)blockIdx.z * sA;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = 0; k < 32; k += 8) t[ty + k][tx] = M[(size_t)(ti * 32 + ty + k) * lda + tj * 32 + tx];
  __syncthreads();
  for (int k = 0; k < 32; k += 8) {
    const int r = tj * 32 + ty + k, c = ti * 32 + tx;  // transposed position
    if (c > r) M[(size_t)r * lda + c] = t[tx][ty + k];
  }
}

cudaError_t launch_symmetrize(double* A, int n, int lda, int64_t sA, int batch, cudaStream_t st) {
  symmetrize_kernel<<<dim3(n / 32, n / 32, batch), dim3(32, 8), 0, st>>>(A, n, lda, sA);
  return counted(cudaGetLastError());
}

// --------------------------------------------------------------------------
// mean_i = mu0 + sum_j k(xs_i, x_j) alpha_j ; K(Xs, X) is never materialised.
struct MeanArgs {
  const double* Xs; int m; int64_t sXs; const int32_t* mv;
  const double* X; int n; int64_t sX; const int32_t* nv;
  const double* theta; int64_t sTheta;
  const double* alpha; int64_t sAlpha;
  double* mean; int64_t sMean;
  int kind, dim;
};

template <int KIND>
__global__ void __launch_bounds__(128) post_mean_kernel(MeanArgs a) {
  __shared__ double xs[MAXD][128];
  __shared__ double ns[128];
  __shared__ double al[128];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int D = a.dim;
  const double* th = a.theta + (size_t)b * a.sTheta;
  const double scale = th[TH_SCALE];
  const int mv = a.mv ? a.mv[b] : a.m;
  const int nv = a.nv ? a.nv[b] : a.n;
  double xi[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d)
    if (d < D) xi[d] = (i < mv) ? a.Xs[(size_t)b * a.sXs + (size_t)i * D + d] / th[TH_BW + d] : 0.0;
  double ni = 0.0;
#pragma unroll
  for (int d = 0; d < MAXD; ++d)
    if (d < D) ni = sqnorm_np(ni, xi[d]);
  const double* X = a.X + (size_t)b * a.sX;
  const double* alpha = a.alpha + (size_t)b * a.sAlpha;
  double acc = 0.0;
  for (int j0 = 0; j0 < nv; j0 += 128) {
    __syncthreads();
    for (int e = threadIdx.x; e < 128 * D; e += 128) {
      const int r = e / D, d = e % D;
      xs[d][r] = (j0 + r < nv) ? X[(size_t)(j0 + r) * D + d] / th[TH_BW + d] : 0.0;
    }
    al[threadIdx.x] = (j0 + threadIdx.x < nv) ? alpha[j0 + threadIdx.x] : 0.0;
    __syncthreads();
    {
      double s = 0.0;
      for (int d = 0; d < D; ++d) s = sqnorm_np(s, xs[d][threadIdx.x]);
      ns[threadIdx.x] = s;
    }
    __syncthreads();
    // padded training points carry alpha = 0, so whole groups of 4 are evaluated: four
    // independent exp() chains per step instead of one (the sum order is unchanged)
    const int lim = (min(128, nv - j0) + 3) & ~3;
    for (int j = 0; j < lim; j += 4) {
      double kv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
          if (d < D) dot = fma(xi[d], xs[d][j + u], dot);
        kv[u] = dist2_np(ni, ns[j + u], dot);
      }
      if constexpr (KIND == OCBO_KERNEL_SE) {
        kern_eval_se<4>(scale, kv);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) kv[u] = kern_eval(KIND, scale, kv[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fma(kv[u], al[j + u], acc);
    }
  }
  if (i < a.m) a.mean[(size_t)b * a.sMean + i] = (i < mv) ? th[TH_MU0] + acc : 0.0;
}

cudaError_t launch_post_mean(int kind, int dim, const double* Xs, int m, int64_t sXs,
                             const int32_t* mv, const double* X, int n, int64_t sX,
                             const int32_t* nv, const double* theta, int64_t sTheta,
                             const double* alpha, int64_t sAlpha, double* mean, int64_t sMean,
                             int batch, cudaStream_t st) {
  MeanArgs a{Xs, m, sXs, mv, X, n, sX, nv, theta, sTheta, alpha, sAlpha, mean, sMean, kind, dim};
  const dim3 grid((m + 127) / 128, batch);
  switch (kind) {
    case OCBO_KERNEL_SE: post_mean_kernel<OCBO_KERNEL_SE><<<grid, 128, 0, st>>>(a); break;
    case OCBO_KERNEL_MATERN12: post_mean_kernel<OCBO_KERNEL_MATERN12><<<grid, 128, 0, st>>>(a); break;
    case OCBO_KERNEL_MATERN32: post_mean_kernel<OCBO_KERNEL_MATERN32><<<grid, 128, 0, st>>>(a); break;
    default: post_mean_kernel<OCBO_KERNEL_MATERN52><<<grid, 128, 0, st>>>(a); break;
  }
  return counted(cudaGetLastError());
}

}  // namespace ocbo
