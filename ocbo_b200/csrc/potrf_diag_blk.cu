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
// Measured on B200 (tests/tools/bench_diag.py, diag_trace.cu): 38-39 us per block against 60 us
// for the column-per-barrier kernel; per block step 4.5 k clocks in the factor+solve chain
// (280 per column with all 8 warps riding: latency), 2.0-3.8 k in the update loop (shared-memory
// bandwidth), ~0.8 k for the stores of the finished block column and the first operand loads.  Tried and dropped: a look-ahead variant (rest of the update interleaved with the next
// block step's chain in one uniform loop) -- bit-identical too, but the loop form of the chain
// (runtime pivot lane, full-width shuffles) cost 450 clocks per column: 53 us per block.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

namespace {
constexpr int NBR = TILE;      // 128
constexpr int G = 16;          // thread grid edge = block step width
constexpr int NBLK = NBR / G;  // 8
constexpr int PLD = 18;        // row stride of the operand buffers: even, so that two consecutive
                               // k of one row are one aligned 16-byte load

struct BlkSmem {
  double D[G][G + 1];     // diagonal block in, L_JJ out (lower incl. diagonal)
  __align__(16) double P[NBR][PLD];  // panel rows (absolute row index) x 16 columns k
  __align__(16) double X[NBR][PLD];  // the 16 rows k of W finalised by this block step, stored
                                     // TRANSPOSED: X[column of W][k]
  int fail;
};

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

// ---- gather (registers -> shared), block step J ------------------------------------
template <int J>
__device__ __forceinline__ void blk_gather(const double (&a)[NBLK][NBLK],
                                           const double (&w)[NBLK][NBLK], BlkSmem& sm, int tx,
                                           int ty) {
  sm.D[ty][tx] = a[J][J];
#pragma unroll
  for (int ia = J + 1; ia < NBLK; ++ia) sm.P[ty + G * ia][tx] = a[ia][J];
#pragma unroll
  for (int ib = 0; ib <= J; ++ib) sm.X[tx + G * ib][ty] = w[J][ib];
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
// (one out-of-line copy, called from every block step: it only touches shared memory, and
// inlined eight times it pushed the kernel out of the instruction cache)
__device__ __noinline__ void blk_factor_solve(BlkSmem& sm, int J, int tid) {
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
    for (int j = 0; j < G; ++j) a[j] = sm.X[wc][j];
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
    for (int j = 0; j < G; ++j) sm.X[wc][j] = l[j];
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
    for (int ib = 0; ib < J; ++ib) Inv[iJ * NBR + tx + G * ib] = sm.X[tx + G * ib][ty];
    Inv[iJ * NBR + jJ] = low ? sm.X[tx + G * J][ty] : 0.0;
  }
#ifdef OCBO_BLK_TRACE
  if (threadIdx.x == 0 && blockIdx.x == 0) g_blk_trace[8 * J + 3] = clock64();
#endif
  if (J + 1 >= NBLK) return;
  // rank-16 update, two columns k per operand load and the next pair prefetched.  Measured
  // (tests/tools/diag_trace.cu): this loop is bound by shared-memory BANDWIDTH -- 128 bytes per
  // clock and SM, counted per lane, broadcast or not: 256 threads x (15 - J) operands x 8 bytes
  // per column = 256 (15 - J) clocks per 16 columns, 3.8 k clocks at J = 0 against 2.2 k of FP64
  // pipe.  8- or 16-byte loads, prefetch or not, 16 / 8 / 4 loop trips all time the same; only a
  // layout with more register reuse per loaded operand (DMMA fragments) would move it.
  // The per-element order of the subtractions (k ascending) is unchanged.
  double2 pi[2][NBLK], pj[2][NBLK], xk[2][NBLK];
  auto load = [&](int k, int s) {
#pragma unroll
    for (int q = J + 1; q < NBLK; ++q) {
      pi[s][q] = *reinterpret_cast<const double2*>(&sm.P[ty + G * q][k]);
      pj[s][q] = *reinterpret_cast<const double2*>(&sm.P[tx + G * q][k]);
    }
#pragma unroll
    for (int q = 0; q <= J; ++q) xk[s][q] = *reinterpret_cast<const double2*>(&sm.X[tx + G * q][k]);
  };
  auto apply = [&](int s) {
#pragma unroll
    for (int ia = J + 1; ia < NBLK; ++ia) {
#pragma unroll
      for (int ib = J + 1; ib <= ia; ++ib) {
        a[ia][ib] = fma(-pi[s][ia].x, pj[s][ib].x, a[ia][ib]);
        a[ia][ib] = fma(-pi[s][ia].y, pj[s][ib].y, a[ia][ib]);
      }
#pragma unroll
      for (int ib = 0; ib <= J; ++ib) {
        w[ia][ib] = fma(-pi[s][ia].x, xk[s][ib].x, w[ia][ib]);
        w[ia][ib] = fma(-pi[s][ia].y, xk[s][ib].y, w[ia][ib]);
      }
    }
  };
  load(0, 0);
#ifdef OCBO_BLK_TRACE
  if (threadIdx.x == 0 && blockIdx.x == 0) g_blk_trace[8 * J + 4] = clock64();
#endif
#pragma unroll 1
  for (int k = 0; k < G; k += 4) {
    load(k + 2, 1);
    apply(0);
    if (k + 4 < G) load(k + 4, 0);
    apply(1);
  }
}

// ---- one block step ---------------------------------------------------------------------
template <int J>
__device__ __forceinline__ int blk_step(double (&a)[NBLK][NBLK], double (&w)[NBLK][NBLK],
                                        BlkSmem& sm, int tx, int ty, int tid, double* Ab, int lda,
                                        double* Inv) {
  // block column J of W comes into existence here (identity on the diagonal block, zero
  // below): until now it was a constant, so it costs no registers before this step -- and
  // block column J of L / block row J of W cost none after it (they are stored in blk_update).
  w[J][J] = (tx == ty) ? 1.0 : 0.0;
#pragma unroll
  for (int ia = J + 1; ia < NBLK; ++ia) w[ia][J] = 0.0;
  BLK_MARK(8 * J + 0);
  blk_gather<J>(a, w, sm, tx, ty);
  __syncthreads();
  BLK_MARK(8 * J + 1);
  blk_factor_solve(sm, J, tid);
  BLK_MARK(8 * J + 2);
  __syncthreads();
  const int fail = sm.fail;
  if (fail) return G * J + fail;
  BLK_MARK(8 * J + 5);
  blk_update<J>(a, w, sm, tx, ty, Ab, lda, Inv);
  BLK_MARK(8 * J + 6);
  __syncthreads();
  BLK_MARK(8 * J + 7);
  return 0;
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
#ifdef OCBO_BLK_REPS  // development: run the body again on its own output to time it with a warm
  for (int rep = 0; rep < OCBO_BLK_REPS; ++rep) {  // instruction cache (tests/tools/diag_trace.cu)
    __syncthreads();
#endif
  double a[NBLK][NBLK];  // only ib <= ia used
  double w[NBLK][NBLK];
  BLK_MARK(64);
#pragma unroll
  for (int ia = 0; ia < NBLK; ++ia)
#pragma unroll
    for (int ib = 0; ib <= ia; ++ib) {
      const int i = ty + G * ia, j = tx + G * ib;
      a[ia][ib] = Ab[(size_t)i * lda + j];
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
  // straight-line sequence of the block steps (no loop, no switch): the register allocator sees
  // which blocks of a and w are live in which step
  int fail = 0;
  if (stages > 0) fail = blk_step<0>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 1) fail = blk_step<1>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 2) fail = blk_step<2>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 3) fail = blk_step<3>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 4) fail = blk_step<4>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 5) fail = blk_step<5>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 6) fail = blk_step<6>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (!fail && stages > 7) fail = blk_step<7>(a, w, sm, tx, ty, tid, Ab, lda, Inv);
  if (fail) {
    if (tid == 0) info[b] = j0 + fail;
    return;
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
#ifdef OCBO_BLK_REPS
  }
#endif
}
}  // namespace

cudaError_t launch_potrf_diag_blk(double* A, int lda, int64_t sA, int j0, double* invdiag,
                                  int64_t sInv, int32_t* info, int batch, cudaStream_t st,
                                  const int32_t* nv) {
  potrf_diag_blk_kernel<<<batch, G * G, 0, st>>>(A, lda, sA, j0, invdiag, sInv, info, nv);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
