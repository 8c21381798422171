// Blocked, register-resident factorisation + inversion of one 128 x 128 diagonal block.
//
// Same data layout as potrf_diag_reg.cu (256 threads = 16 x 16 grid, thread (ty, tx) owns
// A[ty + 16a][tx + 16b] and the matching element of W, the running right-hand side I of
// L X = I, 36 lower blocks each, in registers), same arithmetic in the same order -- the
// result is BIT-IDENTICAL -- but the 128 column steps are regrouped into 8 block steps of 16
// columns, so the serial rsqrt chain runs inside ONE warp (shuffles, no block barrier) and
// the threads only meet at 3 barriers per block step instead of 16:
//
//   gather   owners publish the 16 x 16 diagonal block D, the 16-column panel below it and
//            the 16 rows of W that this block step finalises            (shared memory)
//   factor   lanes 0-15 of every warp: D = L_JJ L_JJ^T.  Lane i keeps row i of the SYMMETRIC
//            block; step k broadcasts raw row k from lane k, every lane forms
//            rinv = rsqrt(a_kk), l_i = a_ik rinv and a_ij -= l_i l_j.  Chain per column:
//            shfl + rsqrt (MUFU + 4 dependent fp64 ops) + mul + fma = 230 clocks measured.
//   solve    lanes 16-31 of every warp ride the same instruction stream: panel rows
//            (P = X L_JJ^-T) and the columns of W_J (W_J <- L_JJ^-1 W_J), see blk_factor_solve
//   update   everyone: rank-16 update of its registers, A -= P P^T, W -= P W_J, operands
//            read from shared memory (broadcast / conflict-free strides).
//
// Measured on B200 (tests/tools/bench_diag.py, diag_trace.cu): 40 us per block against 60 us for
// the column-per-barrier kernel; per block step 4.35 k clocks in the factor+solve chain
// (270 per column with all 8 warps riding), 2.5-4.3 k in the update, 8 k once for the final
// stores.  Tried and dropped: a look-ahead variant (rest of the update interleaved with the next
// block step's chain in one uniform loop) -- bit-identical too, but the loop form of the chain
// (runtime pivot lane, full-width shuffles) cost 450 clocks per column: 53 us per block.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

namespace {
constexpr int NBR = TILE;      // 128
constexpr int G = 16;          // thread grid edge = block step width
constexpr int NBLK = NBR / G;  // 8
constexpr int PLD = 17;        // panel buffer row stride (odd: column access conflict-free)
constexpr int XLD = NBR + 2;   // W-row buffer row stride

struct BlkSmem {
  double D[G][G + 1];     // diagonal block in, L_JJ out (lower incl. diagonal)
  double P[NBR][PLD];     // panel rows (absolute row index), 16 columns
  double X[G][XLD];       // the 16 rows of W finalised by this block step
  int fail;
};

// ---- gather (registers -> shared), block step J ------------------------------------
template <int J>
__device__ __forceinline__ void blk_gather(const double (&a)[NBLK][NBLK],
                                           const double (&w)[NBLK][NBLK], BlkSmem& sm, int tx,
                                           int ty) {
  sm.D[ty][tx] = a[J][J];
#pragma unroll
  for (int ia = J + 1; ia < NBLK; ++ia) sm.P[ty + G * ia][tx] = a[ia][J];
#pragma unroll
  for (int ib = 0; ib <= J; ++ib) sm.X[ty][tx + G * ib] = w[J][ib];
}

// ---- factor + solve, every warp -------------------------------------------------------
// Lanes 0-15 of EVERY warp factor the 16 x 16 block redundantly (lane i keeps row i of the
// symmetric block); lanes 16-31 carry 16 "riders" per warp through the very same
// instruction stream: a rider is either a panel row x (P = X L_JJ^-T: p_c = x_c / L_cc,
// x_c2 -= p_c L_c2,c) or a column x of W_J (W_J <- L_JJ^-1 W_J: v_k = x_k / L_kk,
// x_k2 -= L_k2,k v_k) -- with l_i = x_k rinv and L_j,k = a_kj rinv both are
// x_j -= l_i (a_kj rinv), exactly the diagonal rows' update.  8 warps x 16 riders = 128 =
// 16 (7 - J) panel rows + 16 (J + 1) columns of W_J for every J, so the panel solve and the
// W solve cost no time of their own: they ride in the stall slots of the rsqrt chain.
__device__ __forceinline__ void blk_factor_solve(BlkSmem& sm, int J, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int i = lane & (G - 1);
  const bool rider = lane >= G;
  const int r = warp * G + i;             // rider index 0 .. 127
  const int npanel = G * (NBLK - 1 - J);  // riders < npanel are panel rows, the rest W columns
  const bool is_row = r < npanel;
  double* prow = sm.P[G * (J + 1) + (is_row ? r : 0)];
  const int wc = is_row ? 0 : r - npanel;
  double a[G], l[G];
  if (!rider) {
#pragma unroll
    for (int j = 0; j < G; ++j) a[j] = (j <= i) ? sm.D[i][j] : sm.D[j][i];  // symmetric row i
  } else if (is_row) {
#pragma unroll
    for (int j = 0; j < G; ++j) a[j] = prow[j];
  } else {
#pragma unroll
    for (int j = 0; j < G; ++j) a[j] = sm.X[j][wc];
  }
  int fail = 0;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    double rk[G];
#pragma unroll
    for (int j = k; j < G; ++j) rk[j] = __shfl_sync(0xffffffffu, a[j], k);  // raw row k
    const double akk = rk[k];
    if (!fail && !(akk > 0.0)) fail = k + 1;  // also NaN; identical in every lane
    const double rinv = rsqrt(akk);
    const double li = a[k] * rinv;  // L_ik (L_kk for i == k) | p_k of a panel row | v_k of a W column
    l[k] = li;
#pragma unroll
    for (int j = k + 1; j < G; ++j) a[j] = fma(-li, rk[j] * rinv, a[j]);
  }
  __syncwarp();
  if (!rider) {
    if (warp == 0) {
#pragma unroll
      for (int j = 0; j < G; ++j)
        if (j <= i) sm.D[i][j] = l[j];
    }
  } else if (is_row) {
#pragma unroll
    for (int j = 0; j < G; ++j) prow[j] = l[j];
  } else {
#pragma unroll
    for (int j = 0; j < G; ++j) sm.X[j][wc] = l[j];
  }
  if (tid == 0) sm.fail = fail;
}

// ---- update (shared -> registers), block step J ---------------------------------------
template <int J>
__device__ __forceinline__ void blk_update(double (&a)[NBLK][NBLK], double (&w)[NBLK][NBLK],
                                           const BlkSmem& sm, int tx, int ty, double* Ab, int lda,
                                           double* Inv) {
  // Block column J of L and block row J of W are final now: they go out to global memory here
  // (fire-and-forget stores that overlap the update below) instead of in one 8 k-clock burst of
  // 128 stores per thread at the end of the kernel.
  {
    const int iJ = ty + G * J, jJ = tx + G * J;
    const bool low = tx <= ty;
    Ab[(size_t)iJ * lda + jJ] = low ? sm.D[ty][tx] : 0.0;
#pragma unroll
    for (int ia = J + 1; ia < NBLK; ++ia) Ab[(size_t)(ty + G * ia) * lda + jJ] = sm.P[ty + G * ia][tx];
#pragma unroll
    for (int ib = 0; ib < J; ++ib) Inv[iJ * NBR + tx + G * ib] = sm.X[ty][tx + G * ib];
    Inv[iJ * NBR + jJ] = low ? sm.X[ty][tx + G * J] : 0.0;
  }
  if (J + 1 >= NBLK) return;
#pragma unroll 1
  for (int k = 0; k < G; ++k) {
    double pi[NBLK], pj[NBLK], xk[NBLK];
#pragma unroll
    for (int q = J + 1; q < NBLK; ++q) {
      pi[q] = sm.P[ty + G * q][k];
      pj[q] = sm.P[tx + G * q][k];
    }
#pragma unroll
    for (int q = 0; q <= J; ++q) xk[q] = sm.X[k][tx + G * q];
#pragma unroll
    for (int ia = J + 1; ia < NBLK; ++ia) {
#pragma unroll
      for (int ib = J + 1; ib <= ia; ++ib) a[ia][ib] = fma(-pi[ia], pj[ib], a[ia][ib]);
#pragma unroll
      for (int ib = 0; ib <= J; ++ib) w[ia][ib] = fma(-pi[ia], xk[ib], w[ia][ib]);
    }
  }
}

#ifdef OCBO_BLK_TRACE
__device__ long long g_blk_trace[8 * 8 + 4];
#define BLK_MARK(slot)                                        \
  do {                                                        \
    if (tid == 0 && blockIdx.x == 0) g_blk_trace[slot] = clock64(); \
  } while (0)
#else
#define BLK_MARK(slot) \
  do {                 \
  } while (0)
#endif

#define OCBO_BLK_DISPATCH(FN, J, ...)         \
  switch (J) {                                \
    case 0: FN<0>(__VA_ARGS__); break;        \
    case 1: FN<1>(__VA_ARGS__); break;        \
    case 2: FN<2>(__VA_ARGS__); break;        \
    case 3: FN<3>(__VA_ARGS__); break;        \
    case 4: FN<4>(__VA_ARGS__); break;        \
    case 5: FN<5>(__VA_ARGS__); break;        \
    case 6: FN<6>(__VA_ARGS__); break;        \
    default: FN<7>(__VA_ARGS__); break;       \
  }

__global__ void __launch_bounds__(G * G, 1)
potrf_diag_blk_kernel(double* A, int lda, int64_t sA, int j0, double* invdiag, int64_t sInv,
                      int32_t* info, const int32_t* nv) {
  __shared__ BlkSmem sm;
  const int b = blockIdx.x;
  if (info[b] != 0) return;
  const int tid = threadIdx.x;
  const int tx = tid & (G - 1), ty = tid >> 4;
  double* Ab = A + (size_t)b * sA + (size_t)j0 * lda + j0;

  double* Inv = invdiag + (size_t)b * sInv + (size_t)(j0 / NBR) * NBR * NBR;
  double a[NBLK][NBLK];  // only ib <= ia used
  double w[NBLK][NBLK];
  BLK_MARK(64);
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia)
#pragma unroll
    for (int ib = 0; ib <= ia; ++ib) {
      const int i = ty + G * ia, j = tx + G * ib;
      a[ia][ib] = Ab[(size_t)i * lda + j];
      w[ia][ib] = (i == j) ? 1.0 : 0.0;
    }
  // blocks above the diagonal of L and of inv(L) are zero (consumers read full 128 x 128 tiles)
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia)
#pragma unroll
    for (int ib = ia + 1; ib < NBLK; ++ib) {
      Ab[(size_t)(ty + G * ia) * lda + tx + G * ib] = 0.0;
      Inv[(ty + G * ia) * NBR + tx + G * ib] = 0.0;
    }
  // identity padding: only the block steps that hold valid columns are run (bit-identical)
  int stages = NBLK;
  if (nv) {
    int nvl = nv[b] - j0;
    nvl = nvl < 0 ? 0 : (nvl > NBR ? NBR : nvl);
    stages = (nvl + G - 1) / G;
  }
#pragma unroll 1
  for (int J = 0; J < stages; ++J) {
    BLK_MARK(8 * J + 0);
    OCBO_BLK_DISPATCH(blk_gather, J, a, w, sm, tx, ty)
    __syncthreads();
    BLK_MARK(8 * J + 1);
    blk_factor_solve(sm, J, tid);
    BLK_MARK(8 * J + 2);
    __syncthreads();
    if (sm.fail) {
      if (tid == 0) info[b] = j0 + G * J + sm.fail;
      return;
    }
    BLK_MARK(8 * J + 3);
    BLK_MARK(8 * J + 4);
    BLK_MARK(8 * J + 5);
    OCBO_BLK_DISPATCH(blk_update, J, a, w, sm, tx, ty, Ab, lda, Inv)
    BLK_MARK(8 * J + 6);
    __syncthreads();
    BLK_MARK(8 * J + 7);
  }
  BLK_MARK(65);
  // block rows that only hold identity padding were not touched: L stays as it came in,
  // inv(L) is the identity there
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia) {
    if (ia >= stages) {
#pragma unroll
      for (int ib = 0; ib <= ia; ++ib)
        Inv[(ty + G * ia) * NBR + tx + G * ib] = (ia == ib && tx == ty) ? 1.0 : 0.0;
    }
  }
  BLK_MARK(66);
}
}  // namespace

cudaError_t launch_potrf_diag_blk(double* A, int lda, int64_t sA, int j0, double* invdiag,
                                  int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                                  const int32_t* nv) {
  potrf_diag_blk_kernel<<<batch, G * G, 0, st>>>(A, lda, sA, j0, invdiag, sInv, info, nv);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
