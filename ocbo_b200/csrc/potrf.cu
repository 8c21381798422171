// Diagonal-block Cholesky + triangular inverse, and the alpha / LML solve.
// These are the latency-bound (one CTA per problem) pieces of the path; the
// flops live in the DMMA kernels of gemm_kernels.cu.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

constexpr int NB = TILE;        // 128
constexpr int LDS_ = NB + 1;    // padded smem leading dimension
constexpr int PD_THREADS = 4 * NB;

// Factor the NB x NB diagonal block at (j0, j0): A = L L^T (lower referenced).
// Output: L in place (strictly upper part of the block zeroed) and inv(L) to
// invdiag (full NB x NB, upper part zero).
// Shared-memory layout: one (NB x NB+1) array; the lower triangle holds L, the
// strictly upper triangle holds inv(L)^T, diag of inv(L) in dinv[].
// Failure (pivot <= 0 or NaN, LAPACK dpotrf convention) -> info[b] = column+1.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* A, int lda, int64_t sA, int j0, double* invdiag, int64_t sInv,
                  int32_t* info) {
  extern __shared__ double sm[];
  double* As = sm;                  // NB * LDS_
  double* col = sm + NB * LDS_;     // NB
  double* dinv = col + NB;          // NB
  const int b = blockIdx.x;
  if (info[b] != 0) return;
  const int tid = threadIdx.x;
  double* Ab = A + (size_t)b * sA + (size_t)j0 * lda + j0;

  for (int e = tid; e < NB * NB; e += PD_THREADS) {
    const int r = e / NB, c = e % NB;
    As[r * LDS_ + c] = (c <= r) ? Ab[(size_t)r * lda + c] : 0.0;
  }
  const int tx = tid & 31, ty = tid >> 5;  // 32 x 16
  int fail = 0;
  for (int k = 0; k < NB; ++k) {
    __syncthreads();
    const double akk = As[k * LDS_ + k];
    if (!(akk > 0.0)) {  // also catches NaN
      fail = k + 1;
      break;  // uniform: every thread read the same value
    }
    const double d = sqrt(akk);
    const double rinv = 1.0 / d;
    if (tid > k && tid < NB) {
      const double l = As[tid * LDS_ + k] * rinv;
      col[tid] = l;
      As[tid * LDS_ + k] = l;
    }
    if (tid == k) As[k * LDS_ + k] = d;
    __syncthreads();
    for (int i = k + 1 + ty; i < NB; i += 16) {
      const double li = col[i];
      double* row = As + i * LDS_;
      for (int j = k + 1 + tx; j <= i; j += 32) row[j] = fma(-li, col[j], row[j]);
    }
  }
  if (fail) {
    if (tid == 0) info[b] = j0 + fail;
    return;
  }
  __syncthreads();
  // L back to global, strictly upper part of the block zeroed
  for (int e = tid; e < NB * NB; e += PD_THREADS) {
    const int r = e / NB, c = e % NB;
    Ab[(size_t)r * lda + c] = (c <= r) ? As[r * LDS_ + c] : 0.0;
  }
  if (tid < NB) dinv[tid] = 1.0 / As[tid * LDS_ + tid];
  __syncthreads();
  // inverse: quad (4 lanes) per column c; X[i][c] stored at As[c][i], i > c
  {
    const int c = tid >> 2, q = tid & 3;
    const int c0 = (tid >> 5) * 8;  // first column of this warp
    const double dc = dinv[c];
    for (int i = c0 + 1; i < NB; ++i) {
      double part = 0.0;
      if (i > c) {
        const double* Li = As + i * LDS_;
        const double* Xc = As + c * LDS_;
        for (int k = c + q; k < i; k += 4) {
          const double xk = (k == c) ? dc : Xc[k];
          part = fma(Li[k], xk, part);
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (i > c && q == 0) As[c * LDS_ + i] = -part * dinv[i];
      __syncwarp();
    }
  }
  __syncthreads();
  double* Inv = invdiag + (size_t)b * sInv + (size_t)(j0 / NB) * NB * NB;
  for (int e = tid; e < NB * NB; e += PD_THREADS) {
    const int r = e / NB, c = e % NB;
    double v = 0.0;
    if (c == r) v = dinv[r];
    else if (c < r) v = As[c * LDS_ + r];
    Inv[e] = v;
  }
}

cudaError_t launch_potrf_diag(double* A, int lda, int64_t sA, int j0, double* invdiag,
                              int64_t sInv, int32_t* info, int batch, cudaStream_t st) {
  constexpr int SMEM = (NB * LDS_ + 2 * NB) * 8;
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  potrf_diag_kernel<<<batch, PD_THREADS, SMEM, st>>>(A, lda, sA, j0, invdiag, sInv, info);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// w = L^{-1} (y - mu0);  alpha = L^{-T} w;  lml = -1/2 w^T w - sum log L_ii
//                                                - nv/2 log(2 pi)
// Block forward / backward substitution with the inverted diagonal blocks.
// One CTA (256 threads) per problem.
// ---------------------------------------------------------------------------
struct SolveArgs {
  const double* L; int n; int ldl; int64_t sL;
  const double* invdiag; int64_t sInv;
  const double* y; int64_t sY; const int32_t* nv;
  const double* theta; int64_t sTheta;
  double* alpha; int64_t sAlpha; double* lml;
  const int32_t* info;
};

__global__ void __launch_bounds__(256) solve_alpha_lml_kernel(SolveArgs a) {
  extern __shared__ double sm[];
  double* r = sm;            // n
  double* tmp = sm + a.n;    // NB
  __shared__ double red[8];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.info && a.info[b] != 0) {
    if (a.lml && tid == 0) a.lml[b] = CUDART_NAN;
    return;
  }
  const int n = a.n, nblk = n / NB;
  const int nv = a.nv ? a.nv[b] : n;
  const double mu0 = a.theta[(size_t)b * a.sTheta + TH_MU0];
  const double* L = a.L + (size_t)b * a.sL;
  const double* Inv = a.invdiag + (size_t)b * a.sInv;
  const double* y = a.y + (size_t)b * a.sY;
  for (int i = tid; i < n; i += 256) r[i] = (i < nv) ? y[i] - mu0 : 0.0;
  __syncthreads();
  // forward
  for (int I = 0; I < nblk; ++I) {
    const int base = I * NB;
    for (int rr = warp; rr < NB; rr += 8) {
      const double* Lrow = L + (size_t)(base + rr) * a.ldl;
      double s = 0.0;
      for (int k = lane; k < base; k += 32) s = fma(Lrow[k], r[k], s);
      s = warp_sum(s);
      if (lane == 0) tmp[rr] = r[base + rr] - s;
    }
    __syncthreads();
    const double* Ib = Inv + (size_t)I * NB * NB;
    for (int rr = warp; rr < NB; rr += 8) {
      double s = 0.0;
      for (int k = lane; k <= rr; k += 32) s = fma(Ib[rr * NB + k], tmp[k], s);
      s = warp_sum(s);
      if (lane == 0) r[base + rr] = s;
    }
    __syncthreads();
  }
  // quadratic form and log-determinant
  double quad = 0.0, logdet = 0.0;
  for (int i = tid; i < nv; i += 256) {
    quad = fma(r[i], r[i], quad);
    logdet += log(L[(size_t)i * a.ldl + i]);
  }
  quad = warp_sum(quad);
  logdet = warp_sum(logdet);
  if (lane == 0) red[warp] = -0.5 * quad - logdet;
  __syncthreads();
  if (tid == 0 && a.lml) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    a.lml[b] = s - 0.5 * (double)nv * 1.8378770664093453;  // log(2 pi)
  }
  if (!a.alpha) return;
  // backward: alpha = L^{-T} w
  for (int I = nblk - 1; I >= 0; --I) {
    const int base = I * NB;
    // tmp_i = w_i - sum_{j >= base+NB} L[j][base+i] alpha_j ; two half-ranges of j per column
    {
      const int i = tid & (NB - 1), half = tid >> 7;
      double s = 0.0;
      for (int j = base + NB + half; j < n; j += 2) s = fma(L[(size_t)j * a.ldl + base + i], r[j], s);
      if (half == 1) tmp[i] = s;
      __syncthreads();
      if (half == 0) tmp[i] = r[base + i] - s - tmp[i];
      __syncthreads();
    }
    const double* Ib = Inv + (size_t)I * NB * NB;
    if (tid < NB) {
      double s = 0.0;
      for (int k = tid; k < NB; ++k) s = fma(Ib[k * NB + tid], tmp[k], s);
      r[base + tid] = s;
    }
    __syncthreads();
  }
  double* alpha = a.alpha + (size_t)b * a.sAlpha;
  for (int i = tid; i < n; i += 256) alpha[i] = (i < nv) ? r[i] : 0.0;
}

cudaError_t launch_solve_alpha_lml(const double* L, int n, int ldl, int64_t sL,
                                   const double* invdiag, int64_t sInv, const double* y, int64_t sY,
                                   const int32_t* nv, const double* theta, int64_t sTheta,
                                   double* alpha, int64_t sAlpha, double* lml,
                                   const int32_t* info, int batch, cudaStream_t st) {
  SolveArgs a{L, n, ldl, sL, invdiag, sInv, y, sY, nv, theta, sTheta, alpha, sAlpha, lml, info};
  const int smem = (n + NB) * 8;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(solve_alpha_lml_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  solve_alpha_lml_kernel<<<batch, 256, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ocbo
