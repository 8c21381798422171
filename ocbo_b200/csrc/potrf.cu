// Diagonal-block Cholesky + triangular inverse, and the alpha / LML solve.
// These are the latency-bound (one CTA per problem) pieces of the path; the
// flops live in the DMMA kernels of gemm_kernels.cu.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ocbo {

constexpr int NB = TILE;        // 128
constexpr int LDS_ = NB + 1;    // padded smem leading dimension
constexpr int PD_THREADS = 4 * NB;
constexpr int MB = 16;          // micro-block of the blocked triangular inverse

// Factor the NB x NB diagonal block at (j0, j0): A = L L^T (lower referenced).
// Output: L in place (strictly upper part of the block zeroed) and inv(L) to
// invdiag (full NB x NB, upper part zero).
// Shared-memory layout: one (NB x NB+1) array; the lower triangle holds L, the
// strictly upper triangle holds inv(L)^T, diag of inv(L) in dinv[].
// Failure (pivot <= 0 or NaN, LAPACK dpotrf convention) -> info[b] = column+1.
//
// Phase 1 (factor): right-looking, one column per step, two barriers per step;
//   the rank-1 update keeps l_i / l_j in registers and only walks the rows and
//   column passes that are still inside the trailing triangle (warp-uniform
//   bounds), so the work shrinks with the triangle.
// Phase 2 (inverse): 16 x 16 diagonal micro-blocks are inverted by forward
//   substitution in registers (8 blocks in parallel), then block row I of
//   X = L^{-1} is  X_I,: = -inv(L_II) * (L_I,<I * X_<I,:)  -- two short dot
//   products per element, 7 sequential block rows instead of 127 scalar rows.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* A, int lda, int64_t sA, int j0, double* invdiag, int64_t sInv,
                  int32_t* info) {
  extern __shared__ double sm[];
  double* As = sm;                  // NB * LDS_
  double* col = sm + NB * LDS_;     // NB
  double* dinv = col + NB;          // NB
  double* Tm = dinv + NB;           // MB * LDS_  (temporary of the blocked inverse)
  const int b = blockIdx.x;
  if (info[b] != 0) return;
  const int tid = threadIdx.x;
  double* Ab = A + (size_t)b * sA + (size_t)j0 * lda + j0;

  for (int e = tid; e < NB * NB; e += PD_THREADS) {
    const int r = e / NB, c = e % NB;
    As[r * LDS_ + c] = (c <= r) ? Ab[(size_t)r * lda + c] : 0.0;
  }
  const int tx = tid & 31, ty = tid >> 5;  // 32 lanes x 16 warps
  int fail = 0;
  for (int k = 0; k < NB; ++k) {
    __syncthreads();
    const double akk = As[k * LDS_ + k];
    if (!(akk > 0.0)) {  // also catches NaN
      fail = k + 1;
      break;  // uniform: every thread read the same value
    }
    const double rinv = rsqrt(akk);
    if (tid > k && tid < NB) {
      const double l = As[tid * LDS_ + k] * rinv;
      col[tid] = l;
      As[tid * LDS_ + k] = l;
    }
    if (tid == k) As[k * LDS_ + k] = akk * rinv;
    __syncthreads();
    const int base = k + 1;
    const int rem = NB - base;                       // trailing rows / cols
    const int na = (rem > ty) ? (rem - ty + 15) >> 4 : 0;  // rows of this warp
    const int nbp = (rem + 31) >> 5;                 // 32-wide column passes
    double lj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = base + tx + 32 * q;
      lj[q] = (q < nbp && j < NB) ? col[j] : 0.0;
    }
    for (int a = 0; a < na; ++a) {
      const int i = base + ty + 16 * a;
      const double li = col[i];
      double* row = As + i * LDS_;
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = base + tx + 32 * q;
        v[q] = (q < nbp && j <= i) ? row[j] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = base + tx + 32 * q;
        if (q < nbp && j <= i) row[j] = fma(-li, lj[q], v[q]);
      }
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
  // ---- phase 2a: invert the 16 x 16 diagonal micro-blocks (thread = column) ----
  if (tid < NB) {
    const int c = tid, blk = c / MB, cl = c % MB, r0 = blk * MB;
    double x[MB];
#pragma unroll
    for (int i = 0; i < MB; ++i) {
      double s = (i == cl) ? 1.0 : 0.0;
#pragma unroll
      for (int kk = 0; kk < MB; ++kk)
        if (kk < i) s = fma(-As[(r0 + i) * LDS_ + r0 + kk], (kk >= cl) ? x[kk] : 0.0, s);
      x[i] = (i >= cl) ? s / As[(r0 + i) * LDS_ + r0 + i] : 0.0;
    }
    // every thread must finish READING the lower micro-block before the upper
    // part is written: the upper storage of this block is disjoint from the
    // lower triangle, so plain stores are safe.
#pragma unroll
    for (int i = 0; i < MB; ++i) {
      if (i == cl) dinv[c] = x[i];
      else if (i > cl) As[c * LDS_ + r0 + i] = x[i];
    }
  }
  __syncthreads();
  // ---- phase 2b: block rows I = 1 .. 7 ------------------------------------------
  for (int I = 1; I < NB / MB; ++I) {
    const int r0 = I * MB;         // first row of block row I; columns c < r0
    // T[il][c] = sum_{k=c}^{r0-1} L[r0+il][k] * X[k][c]
    for (int e = tid; e < MB * r0; e += PD_THREADS) {
      const int il = e / r0, c = e % r0;
      const double* Li = As + (r0 + il) * LDS_;
      const double* Xc = As + c * LDS_;   // X[k][c] at Xc[k], k > c
      double s0 = Li[c] * dinv[c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = c + 1;
      for (; k + 3 < r0; k += 4) {
        s0 = fma(Li[k], Xc[k], s0);
        s1 = fma(Li[k + 1], Xc[k + 1], s1);
        s2 = fma(Li[k + 2], Xc[k + 2], s2);
        s3 = fma(Li[k + 3], Xc[k + 3], s3);
      }
      for (; k < r0; ++k) s0 = fma(Li[k], Xc[k], s0);
      Tm[il * LDS_ + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    // X[r0+il][c] = - sum_{i2 <= il} invLII[il][i2] * T[i2][c]
    for (int e = tid; e < MB * r0; e += PD_THREADS) {
      const int il = e / r0, c = e % r0;
      double s = dinv[r0 + il] * Tm[il * LDS_ + c];
      for (int i2 = 0; i2 < il; ++i2) s = fma(As[(r0 + i2) * LDS_ + r0 + il], Tm[i2 * LDS_ + c], s);
      As[c * LDS_ + r0 + il] = -s;
    }
    __syncthreads();
  }
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
                              int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                              const int32_t* nv) {
  // default: blocked register-resident kernel (potrf_diag_blk.cu); OCBO_DIAG_KERNEL=reg keeps
  // the column-per-barrier register kernel and =smem the shared-memory version selectable
  // for A/B measurements (blk and reg are bit-identical)
  static int which = -1;
  if (which < 0) {
    const char* e = getenv("OCBO_DIAG_KERNEL");
    which = !e ? 2 : (e[0] == 's' ? 0 : (e[0] == 'r' ? 1 : 2));
  }
  if (which == 2) return launch_potrf_diag_blk(A, lda, sA, j0, invdiag, sInv, info, batch, st, nv);
  if (which == 1) return launch_potrf_diag_reg(A, lda, sA, j0, invdiag, sInv, info, batch, st, nv);
  constexpr int SMEM = (NB * LDS_ + 2 * NB + MB * LDS_) * 8;
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  potrf_diag_kernel<<<batch, PD_THREADS, SMEM, st>>>(A, lda, sA, j0, invdiag, sInv, info);
  return counted(cudaGetLastError());
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
