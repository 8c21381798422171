// Register-resident factorisation + inversion of one 128 x 128 diagonal block.
//
// 256 threads form a 16 x 16 grid; thread (ty, tx) owns the elements
// A[ty + 16a][tx + 16b] (2-D cyclic layout, 8 x 8 blocks of 16 x 16), keeping the
// 36 lower blocks (b <= a) of A and of W (the running right-hand side I of
// L X = I) in registers.  One column per step, ONE barrier per step:
//
//   before the barrier  owners of column k publish A[k:, k] (raw) and owners
//                       of row k publish W[k, :k+1] (raw) in shared memory;
//   after the barrier   everyone forms rinv = rsqrt(A[k][k]), its l_i, l_j and
//                       x_c = W[k][c] * rinv, then applies the rank-1 updates
//                         A[i][j] -= l_i l_j      (i >= j > k)
//                         W[i][c] -= l_i x_c      (i > k >= c)
//                       entirely in registers.
//
// The column loop is unrolled over the 8 block columns (template parameter KB)
// so that every block-level test is resolved at compile time: step k only
// touches the (8-KB)(9-KB)/2 + (8-KB)(KB+1) blocks that are still live.  (A first
// version with run-time block tests executed 1150 instructions per step and was
// slower than the shared-memory kernel it replaces: 164 us vs 140 us per block.)
// The diagonal chain is the serial critical path of the blocked Cholesky
// (64 blocks for m = 8192).
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

namespace {
constexpr int NBR = TILE;      // 128
constexpr int G = 16;          // thread grid edge
constexpr int NBLK = NBR / G;  // 8 blocks per dimension

struct DiagSmem {
  double colA[2][NBR];
  double rowX[2][NBR];
};

// Steps k = 16*KB .. 16*KB+15.  Returns 0 or the 1-based failing column.
template <int KB>
__device__ __forceinline__ int diag_steps(double (&a)[NBLK][NBLK], double (&w)[NBLK][NBLK],
                                          DiagSmem& sm, int tx, int ty) {
#pragma unroll 1
  for (int kc = 0; kc < G; ++kc) {
    const int k = KB * G + kc, buf = k & 1;
    __syncthreads();
    const double akk = sm.colA[buf][k];
    if (!(akk > 0.0)) return k + 1;  // also NaN; uniform across the block
    const double rinv = rsqrt(akk);
    double li[NBLK], lj[NBLK], xk[NBLK];
#pragma unroll
    for (int q = KB; q < NBLK; ++q) {
      const int i = ty + G * q, j = tx + G * q;
      li[q] = (i > k) ? sm.colA[buf][i] * rinv : 0.0;
      lj[q] = (j > k) ? sm.colA[buf][j] * rinv : 0.0;
    }
#pragma unroll
    for (int q = 0; q <= KB; ++q) xk[q] = sm.rowX[buf][tx + G * q] * rinv;
    // rank-1 updates, live blocks only
#pragma unroll
    for (int ia = KB; ia < NBLK; ++ia) {
#pragma unroll
      for (int ib = KB; ib <= ia; ++ib) a[ia][ib] = fma(-li[ia], lj[ib], a[ia][ib]);
#pragma unroll
      for (int ib = 0; ib <= KB; ++ib) w[ia][ib] = fma(-li[ia], xk[ib], w[ia][ib]);
    }
    // finalise column k of L (owners tx == kc) and row k of X (owners ty == kc)
    if (tx == kc) {
#pragma unroll
      for (int ia = KB; ia < NBLK; ++ia) {
        const int i = ty + G * ia;
        if (i > k) a[ia][KB] = li[ia];
        else if (i == k) a[ia][KB] = akk * rinv;
      }
    }
    if (ty == kc) {
#pragma unroll
      for (int ib = 0; ib <= KB; ++ib) w[KB][ib] = xk[ib];
    }
    // publish column k+1 / row k+1 for the next step
    const int nb = buf ^ 1;
    if (kc + 1 < G) {
      if (tx == kc + 1) {
#pragma unroll
        for (int ia = KB; ia < NBLK; ++ia) sm.colA[nb][ty + G * ia] = a[ia][KB];
      }
      if (ty == kc + 1) {
#pragma unroll
        for (int ib = 0; ib <= KB; ++ib) sm.rowX[nb][tx + G * ib] = w[KB][ib];
      }
    } else if (KB + 1 < NBLK) {
      constexpr int K1 = (KB + 1 < NBLK) ? KB + 1 : KB;  // next block column
      if (tx == 0) {
#pragma unroll
        for (int ia = K1; ia < NBLK; ++ia) sm.colA[nb][ty + G * ia] = a[ia][K1];
      }
      if (ty == 0) {
#pragma unroll
        for (int ib = 0; ib <= K1; ++ib) sm.rowX[nb][tx + G * ib] = w[K1][ib];
      }
    }
  }
  return 0;
}

__global__ void __launch_bounds__(G * G, 1)
potrf_diag_reg_kernel(double* A, int lda, int64_t sA, int j0, double* invdiag, int64_t sInv,
                      int32_t* info, const int32_t* nv) {
  __shared__ DiagSmem sm;
  const int b = blockIdx.x;
  if (info[b] != 0) return;
  const int tid = threadIdx.x;
  const int tx = tid & (G - 1), ty = tid >> 4;
  double* Ab = A + (size_t)b * sA + (size_t)j0 * lda + j0;

  double a[NBLK][NBLK];  // only b <= a used
  double w[NBLK][NBLK];
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia)
#pragma unroll
    for (int ib = 0; ib <= ia; ++ib) {
      const int i = ty + G * ia, j = tx + G * ib;
      a[ia][ib] = Ab[(size_t)i * lda + j];
      w[ia][ib] = (i == j) ? 1.0 : 0.0;
    }
  // publish column 0 / row 0
  if (tx == 0) {
#pragma unroll
    for (int ia = 0; ia < NBLK; ++ia) sm.colA[0][ty + G * ia] = a[ia][0];
  }
  if (ty == 0) sm.rowX[0][tx] = w[0][0];

  // Rows/columns >= nv[b] of the matrix are padding (identity, see the header): their
  // elimination steps are no-ops (pivot 1, l = 0), so only the 16-column stages that
  // hold valid columns are run.  Bit-identical to running all eight.
  int stages = NBLK;
  if (nv) {
    int nvl = nv[b] - j0;
    nvl = nvl < 0 ? 0 : (nvl > NBR ? NBR : nvl);
    stages = (nvl + G - 1) / G;
  }
  int fail = 0;
  if (stages > 0) fail = diag_steps<0>(a, w, sm, tx, ty);
  if (!fail && stages > 1) fail = diag_steps<1>(a, w, sm, tx, ty);
  if (!fail && stages > 2) fail = diag_steps<2>(a, w, sm, tx, ty);
  if (!fail && stages > 3) fail = diag_steps<3>(a, w, sm, tx, ty);
  if (!fail && stages > 4) fail = diag_steps<4>(a, w, sm, tx, ty);
  if (!fail && stages > 5) fail = diag_steps<5>(a, w, sm, tx, ty);
  if (!fail && stages > 6) fail = diag_steps<6>(a, w, sm, tx, ty);
  if (!fail && stages > 7) fail = diag_steps<7>(a, w, sm, tx, ty);
  if (fail) {
    if (tid == 0) info[b] = j0 + fail;
    return;
  }
  // L in place (strictly upper part zeroed) and inv(L) to the workspace
  double* Inv = invdiag + (size_t)b * sInv + (size_t)(j0 / NBR) * NBR * NBR;
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia)
#pragma unroll
    for (int ib = 0; ib < NBLK; ++ib) {
      const int i = ty + G * ia, j = tx + G * ib;
      double lv = 0.0, xv = 0.0;
      if (ib <= ia) {
        if (j <= i) {
          lv = a[ia][ib];
          xv = w[ia][ib];
        }
      }
      Ab[(size_t)i * lda + j] = lv;
      Inv[i * NBR + j] = xv;
    }
}
}  // namespace

cudaError_t launch_potrf_diag_reg(double* A, int lda, int64_t sA, int j0, double* invdiag,
                                  int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                                  const int32_t* nv) {
  potrf_diag_reg_kernel<<<batch, G * G, 0, st>>>(A, lda, sA, j0, invdiag, sInv, info, nv);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
