// Diagonal-block Cholesky + triangular inverse, and the alpha / LML solve.
// These are the latency-bound (one CTA per problem) pieces of the path; the
// flops live in the DMMA kernels of gemm_kernels.cu.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ocbo {

constexpr int NB = TILE;        // 128

// The NB x NB diagonal block at (j0, j0): A = L L^T (lower referenced), L in place and inv(L)
// to invdiag.  Failure (pivot <= 0 or NaN, LAPACK dpotrf convention) -> info[b] = column + 1.
// Two register-resident kernels implement it: the blocked one (potrf_diag_blk.cu, default) and
// the column-per-barrier one it is bit-identical to (potrf_diag_reg.cu, OCBO_DIAG_KERNEL=reg,
// kept for that test).  The first shared-memory version (140 us per block) is gone.
cudaError_t launch_potrf_diag(double* A, int lda, int64_t sA, int j0, double* invdiag,
                              int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                              const int32_t* nv) {
  static int which = -1;
  if (which < 0) {
    const char* e = getenv("OCBO_DIAG_KERNEL");
    which = (e && e[0] == 'r') ? 1 : 2;
  }
  if (which == 1) return launch_potrf_diag_reg(A, lda, sA, j0, invdiag, sInv, info, batch, st, nv);
  return launch_potrf_diag_blk(A, lda, sA, j0, invdiag, sInv, info, batch, st, nv);
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
  double* w; int64_t sW;  // optional: w = L^{-1}(y - mu0) (mean = mu0 + V^T w needs no back-solve)
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
  // Each warp owns rows warp, warp + 8, ... of a block and walks RU of them at a time: the row
  // dot products are independent, so their loads are in flight together (one row at a time was
  // one exposed HBM latency per row and 8 warps x 256 B in flight per SM: 254 us for 64 problems
  // of n = 768, 0.8 TB/s, "long scoreboard" 31 cycles per issue in ncu).  Measured for (rows at a
  // time, k chunks per step): (16, 1) 127 us, (4, 4) 145, (16, 2) 181, (4, 2) 186, (2, 4) 192,
  // (8, 2) 209, (8, 4) 314.  Every row keeps its own lane-strided accumulation order, so the
  // values are those of the one-row loop bit for bit.
#ifndef OCBO_SOLVE_RU
#define OCBO_SOLVE_RU 16
#endif
  constexpr int RU = OCBO_SOLVE_RU;
#ifndef OCBO_SOLVE_KQ
#define OCBO_SOLVE_KQ 1
#endif
  constexpr int KQ = OCBO_SOLVE_KQ;  // k chunks of 32 per step (base is a multiple of 128)
  for (int I = 0; I < nblk; ++I) {
    const int base = I * NB;
    for (int rr0 = warp; rr0 < NB; rr0 += 8 * RU) {
      double s[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) s[u] = 0.0;
      const double* Lrow = L + (size_t)(base + rr0) * a.ldl;
      // (all RU x KQ loads of a step are issued before the first use -- written out, the
      // compiler otherwise keeps only two or three in flight)
      for (int k = lane; k < base; k += 32 * KQ) {
        double v[KQ][RU];
#pragma unroll
        for (int q = 0; q < KQ; ++q)
#pragma unroll
          for (int u = 0; u < RU; ++u) v[q][u] = Lrow[(size_t)(8 * u) * a.ldl + k + 32 * q];
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
          const double rk = r[k + 32 * q];
#pragma unroll
          for (int u = 0; u < RU; ++u) s[u] = fma(v[q][u], rk, s[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const double t = warp_sum(s[u]);
        if (lane == 0) tmp[rr0 + 8 * u] = r[base + rr0 + 8 * u] - t;
      }
    }
    __syncthreads();
    const double* Ib = Inv + (size_t)I * NB * NB;
    for (int rr0 = warp; rr0 < NB; rr0 += 8 * RU) {
      double s[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) s[u] = 0.0;
      for (int k = lane; k <= rr0 + 8 * (RU - 1); k += 32) {
        const double tk = tmp[k];
#pragma unroll
        for (int u = 0; u < RU; ++u)
          if (k <= rr0 + 8 * u) s[u] = fma(Ib[(rr0 + 8 * u) * NB + k], tk, s[u]);
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const double t = warp_sum(s[u]);
        if (lane == 0) r[base + rr0 + 8 * u] = t;
      }
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
  if (a.w) {
    double* wout = a.w + (size_t)b * a.sW;
    for (int i = tid; i < n; i += 256) wout[i] = r[i];
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
                                   const int32_t* info, int batch, cudaStream_t st, double* w,
                                   int64_t sW) {
  SolveArgs a{L, n, ldl, sL, invdiag, sInv, y, sY, nv, theta, sTheta, alpha, sAlpha, lml, info,
              w, sW};
  const int smem = (n + NB) * 8;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(solve_alpha_lml_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  solve_alpha_lml_kernel<<<batch, 256, smem, st>>>(a);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
