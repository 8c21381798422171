// DMMA (FP64 tensor-core) tile GEMM core shared by every dense contraction on
// the path: Cholesky trailing updates, panel / left triangular solves,
// Sigma = K - V^T V, and the L*Z sampling product.
//
// One CTA (256 threads = 2 x 4 warps) computes a BM x BN tile; operands stream
// through a 3-stage cp.async ring in shared memory, 16 k per stage.  Operand
// tiles are stored in shared memory in the orientation they have in global
// memory (no transposes); the leading dimensions (20 resp. R+4 doubles) are
// == 4 (mod 16) so that the 16 lanes of a half-warp, reading the DMMA fragment
// element (row = lane>>2, k = lane&3), hit 16 distinct 8-byte banks.
#pragma once
#include "common.cuh"

namespace ocbo {

// Operand tile of R rows (the M or N extent) by BK k-values.
//   KMAJ  = true : global element (r, k) at g[r * ldg + k]   (k contiguous)
//   KMAJ  = false: global element (r, k) at g[k * ldg + r]   (r contiguous)
template <int R, bool KMAJ>
struct OperandTile {
  static constexpr int LD = KMAJ ? (BK + 4) : (R + 4);
  static constexpr int SIZE = KMAJ ? R * LD : BK * LD;  // doubles per stage

  __device__ __forceinline__ static void load(double* s, const double* g, int ldg, int tid) {
    if (KMAJ) {
      constexpr int CHUNKS = R * (BK / 2);
#pragma unroll
      for (int c = tid; c < CHUNKS; c += GEMM_THREADS) {
        const int r = c / (BK / 2);
        const int kc = (c % (BK / 2)) * 2;
        cp_async16(s + r * LD + kc, g + (size_t)r * ldg + kc);
      }
    } else {
      constexpr int CHUNKS = BK * (R / 2);
#pragma unroll
      for (int c = tid; c < CHUNKS; c += GEMM_THREADS) {
        const int k = c / (R / 2);
        const int rc = (c % (R / 2)) * 2;
        cp_async16(s + k * LD + rc, g + (size_t)k * ldg + rc);
      }
    }
  }
  __device__ __forceinline__ static double frag(const double* s, int r, int k) {
    return KMAJ ? s[r * LD + k] : s[k * LD + r];
  }
  // pointer advance for k-tile kt
  __device__ __forceinline__ static const double* at_k(const double* g, int ldg, int kt) {
    return KMAJ ? g + (size_t)kt * BK : g + (size_t)kt * BK * ldg;
  }
};

template <int BM, int BN, bool A_KMAJ, bool B_KMAJ>
struct Gemm {
  using TA = OperandTile<BM, A_KMAJ>;
  using TB = OperandTile<BN, B_KMAJ>;
  static constexpr int STAGE_DOUBLES = TA::SIZE + TB::SIZE;
  static constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;
  static constexpr int WM = BM / 2, WN = BN / 4;  // warp tile
  static constexpr int MI = WM / 8, NI = WN / 8;  // 8x8 DMMA sub-tiles per warp

  // acc += A(BM x 16*ktiles) * B(16*ktiles x BN).  A, B point at the tile origin.
  __device__ __forceinline__ static void mainloop(const double* __restrict__ A, int lda,
                                                  const double* __restrict__ B, int ldb,
                                                  int ktiles, double* smem,
                                                  double (&acc)[MI][NI][2]) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < ktiles) {
        double* st = smem + s * STAGE_DOUBLES;
        TA::load(st, TA::at_k(A, lda, s), lda, tid);
        TB::load(st + TA::SIZE, TB::at_k(B, ldb, s), ldb, tid);
      }
      cp_async_commit();
    }
    for (int kt = 0; kt < ktiles; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) {
        double* st = smem + (nk % STAGES) * STAGE_DOUBLES;
        TA::load(st, TA::at_k(A, lda, nk), lda, tid);
        TB::load(st + TA::SIZE, TB::at_k(B, ldb, nk), ldb, tid);
      }
      cp_async_commit();
      const double* sa = smem + (kt % STAGES) * STAGE_DOUBLES;
      const double* sb = sa + TA::SIZE;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = TA::frag(sa, wm * WM + mi * 8 + g, kk * 4 + t);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = TB::frag(sb, wn * WN + ni * 8 + g, kk * 4 + t);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the ring: smem may be reused
  }

  __device__ __forceinline__ static void zero(double (&acc)[MI][NI][2]) {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }

  // f(row_in_tile, col_in_tile (even), v0, v1): v0 at (row, col), v1 at (row, col+1)
  template <class F>
  __device__ __forceinline__ static void foreach(double (&acc)[MI][NI][2], F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
        f(wm * WM + mi * 8 + g, wn * WN + ni * 8 + 2 * t, acc[mi][ni][0], acc[mi][ni][1]);
  }
};

// linear index t over the lower triangle (incl. diagonal) -> (ti >= tj)
__device__ __forceinline__ void tri_index(int t, int& ti, int& tj) {
  int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  ti = i;
  tj = t - i * (i + 1) / 2;
}

}  // namespace ocbo
