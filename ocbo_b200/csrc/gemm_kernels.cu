// DMMA tile kernels: C (-)= P Q^T, C (-)= P R, Sigma = K(Xa,Xb) - V1^T V2.
#include "gemm_core.cuh"
#include "kernels.h"

namespace ocbo {

// ---------------------------------------------------------------------------
// C_tile (mode 0: =, mode 1: -=)  P_tile * Q_tile^T
//   P: rows follow C rows, k contiguous;  Q: rows follow C cols, k contiguous.
// Used for the Cholesky trailing update (P = Q = L panel, mode 1) and the
// panel solve X <- X * inv(L_jj)^T (P = X itself, Q = inv(L_jj), mode 0).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nt_kernel(NtParams p) {
  using G = Gemm<TILE, TILE, true, true>;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.z;
  if (p.info && p.info[b] != 0) return;
  int ti, tj;
  if (p.tri) {
    tri_index(blockIdx.x, ti, tj);
  } else {
    ti = blockIdx.x / p.ntc;
    tj = blockIdx.x % p.ntc;
    if (p.lower_skip && p.row0 + ti * TILE + TILE - 1 < p.col0 + tj * TILE) return;
  }
  const double* P = p.P + (size_t)b * p.sP + (size_t)ti * TILE * p.ldp;
  const double* Q = p.Q + (size_t)b * p.sQ + (size_t)tj * TILE * p.ldq;
  double* C = p.C + (size_t)b * p.sC + (size_t)ti * TILE * p.ldc + (size_t)tj * TILE;

  double acc[G::MI][G::NI][2];
  G::zero(acc);
  G::mainloop(P, p.ldp, Q, p.ldq, p.ktiles, smem, acc);
  const int ldc = p.ldc;
  if (p.mode == 0) {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(v0, v1);
    });
  } else {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      double2* ptr = reinterpret_cast<double2*>(C + (size_t)r * ldc + c);
      double2 o = *ptr;
      o.x -= v0;
      o.y -= v1;
      *ptr = o;
    });
  }
}

// ---------------------------------------------------------------------------
// C_tile (mode 0: =, 1: -=, 2: = rowvec + )  P_tile * R_tile
//   P: rows follow C rows, k contiguous;  R: k rows, cols follow C cols.
// Used for the left triangular solve (update + multiply by inv(L_II), in
// place) and the sampling product F = mean + L Z (tri_k: k stops at the
// diagonal block of the row tile).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nn_kernel(NnParams p) {
  using G = Gemm<TILE, TILE, true, false>;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.z;
  if (p.info && p.info[b] != 0) return;
  const int tj = blockIdx.x;
  const int ti = p.reverse_rows ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
  const double* P = p.P + (size_t)b * p.sP + (size_t)ti * TILE * p.ldp;
  const double* R = p.R + (size_t)b * p.sR + (size_t)tj * TILE;
  double* C = p.C + (size_t)b * p.sC + (size_t)ti * TILE * p.ldc + (size_t)tj * TILE;
  int ktiles = p.ktiles;
  if (p.tri_k) ktiles = min(ktiles, (ti + 1) * (TILE / BK));

  double acc[G::MI][G::NI][2];
  G::zero(acc);
  G::mainloop(P, p.ldp, R, p.ldr, ktiles, smem, acc);
  const int ldc = p.ldc;
  if (p.mode == 0) {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(v0, v1);
    });
  } else if (p.mode == 1) {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      double2* ptr = reinterpret_cast<double2*>(C + (size_t)r * ldc + c);
      double2 o = *ptr;
      o.x -= v0;
      o.y -= v1;
      *ptr = o;
    });
  } else {
    const double* rv = p.rowvec + (size_t)b * p.sRowvec + (size_t)ti * TILE;
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      const double mu = rv[r];
      *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(mu + v0, mu + v1);
    });
  }
}

// ---------------------------------------------------------------------------
// C_tile = K(Xa_i, Xb_j) - sum_k V1[k][i] V2[k][j]     (posterior covariance)
//   V1: n x ma, V2: n x mb row-major (k rows).  sym: Xa == Xb, V1 == V2,
//   padding diagonal = 1; lower_only skips tiles above the diagonal.
// The Gram term is evaluated in the epilogue from points staged (pre-scaled by
// 1/bandwidth) in the drained cp.async ring, so K(Xs,Xs) never exists in HBM.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) postcov_kernel(CovParams p) {
  using G = Gemm<TILE, TILE, false, false>;
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.z;
  int ti, tj;
  if (p.tri) {
    tri_index(blockIdx.x, ti, tj);
  } else {
    ti = blockIdx.x / p.ntc;
    tj = blockIdx.x % p.ntc;
  }
  const double* V1 = p.V1 + (size_t)b * p.sV1 + (size_t)ti * TILE;
  const double* V2 = p.V2 + (size_t)b * p.sV2 + (size_t)tj * TILE;
  double* C = p.C + (size_t)b * p.sC + (size_t)ti * TILE * p.ldc + (size_t)tj * TILE;

  double acc[G::MI][G::NI][2];
  G::zero(acc);
  G::mainloop(V1, p.ldv1, V2, p.ldv2, p.ktiles, smem, acc);

  // stage scaled points: xa[d][128], xb[d][128]
  const int D = p.dim;
  const double* th = p.theta + (size_t)b * p.sTheta;
  const double scale = th[TH_SCALE];
  double* xa = smem;
  double* xb = smem + MAXD * TILE;
  const double* Xa = p.Xa + (size_t)b * p.sXa + (size_t)ti * TILE * D;
  const double* Xb = p.Xb + (size_t)b * p.sXb + (size_t)tj * TILE * D;
  for (int e = threadIdx.x; e < TILE * D; e += GEMM_THREADS) {
    const int r = e / D, d = e % D;
    const double bw = th[TH_BW + d];
    xa[d * TILE + r] = Xa[e] / bw;
    xb[d * TILE + r] = Xb[e] / bw;
  }
  __syncthreads();
  const int mav = p.mav ? p.mav[b] : p.ma;
  const int mbv = p.mbv ? p.mbv[b] : p.mb;
  const int row0 = ti * TILE, col0 = tj * TILE;
  const int kind = p.kind, sym = p.sym, ldc = p.ldc;
  G::foreach(acc, [&](int r, int c, double v0, double v1) {
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
      if (d < D) {
        const double xr = xa[d * TILE + r];
        const double2 xc = *reinterpret_cast<const double2*>(xb + d * TILE + c);
        const double e0 = xr - xc.x, e1 = xr - xc.y;
        d0 = fma(e0, e0, d0);
        d1 = fma(e1, e1, d1);
      }
    }
    const int gi = row0 + r, gj = col0 + c;
    double o0, o1;
    if (gi < mav && gj < mbv) o0 = kern_eval(kind, scale, d0) - v0;
    else o0 = (sym && gi == gj) ? 1.0 : 0.0;
    if (gi < mav && gj + 1 < mbv) o1 = kern_eval(kind, scale, d1) - v1;
    else o1 = (sym && gi == gj + 1) ? 1.0 : 0.0;
    *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(o0, o1);
  });
}

// ------------------------------- launchers ---------------------------------
template <class K>
static cudaError_t ensure_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t launch_gemm_nt(const NtParams& p, int ntiles, int batch, cudaStream_t st) {
  using G = Gemm<TILE, TILE, true, true>;
  static bool init = false;
  if (!init) {
    cudaError_t e = ensure_smem(gemm_nt_kernel, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    init = true;
  }
  if (ntiles <= 0 || batch <= 0) return cudaSuccess;
  gemm_nt_kernel<<<dim3(ntiles, 1, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(p);
  return counted(cudaGetLastError());
}

cudaError_t launch_gemm_nn(const NnParams& p, int ntr, int ntc, int batch, cudaStream_t st) {
  using G = Gemm<TILE, TILE, true, false>;
  static bool init = false;
  if (!init) {
    cudaError_t e = ensure_smem(gemm_nn_kernel, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    init = true;
  }
  if (ntr <= 0 || ntc <= 0 || batch <= 0) return cudaSuccess;
  gemm_nn_kernel<<<dim3(ntc, ntr, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(p);
  return counted(cudaGetLastError());
}

cudaError_t launch_postcov(const CovParams& p, int ntiles, int batch, cudaStream_t st) {
  using G = Gemm<TILE, TILE, false, false>;
  static_assert(G::SMEM_DOUBLES >= 2 * MAXD * TILE, "ring too small for the point staging");
  static bool init = false;
  if (!init) {
    cudaError_t e = ensure_smem(postcov_kernel, G::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    init = true;
  }
  if (ntiles <= 0 || batch <= 0) return cudaSuccess;
  postcov_kernel<<<dim3(ntiles, 1, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(p);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
