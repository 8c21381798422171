// DMMA (FP64 tensor-core) tile GEMM core shared by every dense contraction on
// the path: Cholesky rank-k updates, panel / left triangular solves,
// Sigma = K - V^T V, and the L*Z sampling product.
//
// One CTA (256 threads = WARPS_M x WARPS_N warps) computes a BM x BN tile;
// operands stream through a 3-stage cp.async ring in shared memory, 16 k per
// stage.  The production shape is 128 x 64 with 32 x 32 warp tiles: 32 fp64
// accumulators per thread keep the kernel under 128 registers so TWO CTAs are
// resident per SM and the prologue / epilogue of one overlaps the DMMA main
// loop of the other (ncu, round 1: a single 128 x 128 CTA per SM left the DMMA
// pipe idle 21 % of the time at k = 512).
//
// Operand tiles are stored in shared memory in the orientation they have in
// global memory (no transposes); the leading dimensions (20 resp. R+4 doubles)
// are == 4 (mod 16) so that the 16 lanes of a half-warp, reading the DMMA
// fragment element (row = lane>>2, k = lane&3), hit 16 distinct 8-byte banks
// (ncu: 0 shared-memory bank conflicts).
#pragma once
#include "common.cuh"
#include "tma.cuh"

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

template <int BM, int BN, bool A_KMAJ, bool B_KMAJ, int WARPS_M, int WARPS_N>
struct Gemm {
  static_assert(WARPS_M * WARPS_N * 32 == GEMM_THREADS, "8 warps per CTA");
  using TA = OperandTile<BM, A_KMAJ>;
  using TB = OperandTile<BN, B_KMAJ>;
  static constexpr int STAGE_DOUBLES = TA::SIZE + TB::SIZE;
  static constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;  // warp tile
  static constexpr int MI = WM / 8, NI = WN / 8;              // 8x8 DMMA sub-tiles per warp
  using Acc = double[MI][NI][2];

  __device__ __forceinline__ static void issue(const double* A, int lda, const double* B, int ldb,
                                               int kt, double* smem, int tid) {
    double* st = smem + (kt % STAGES) * STAGE_DOUBLES;
    TA::load(st, TA::at_k(A, lda, kt), lda, tid);
    TB::load(st + TA::SIZE, TB::at_k(B, ldb, kt), ldb, tid);
  }

  // Start the operand stream (first STAGES-1 k-tiles) -- callers issue this BEFORE
  // loading the C tile so both latencies overlap.
  __device__ __forceinline__ static void prologue(const double* A, int lda, const double* B,
                                                  int ldb, int ktiles, double* smem) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < ktiles) issue(A, lda, B, ldb, s, smem, tid);
      cp_async_commit();
    }
  }

  // acc += A(BM x 16*ktiles) * B(16*ktiles x BN); prologue() must have been called.
  __device__ __forceinline__ static void mainloop(const double* __restrict__ A, int lda,
                                                  const double* __restrict__ B, int ldb,
                                                  int ktiles, double* smem, Acc& acc) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    for (int kt = 0; kt < ktiles; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) issue(A, lda, B, ldb, nk, smem, tid);
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

  // ---- TMA-fed operand stream (tma.cuh): same ring, same fragment addressing -----------------
  // A tile origin (a_in, a_row) and B tile origin (b_in, b_row) in their tensor maps' frames
  // (inner = contiguous coordinate); the k coordinate advances along the inner axis of a k-major
  // operand and along the row axis otherwise.  ``it`` numbers the k-tiles a CTA has consumed since
  // the barriers were initialised (stage = it % STAGES, phase parity = (it / STAGES) & 1), so a CTA
  // that computes several tiles keeps one running count.
  static constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
  struct TmaTile {
    const CUtensorMap* mA; int a_in, a_row, a_b;
    const CUtensorMap* mB; int b_in, b_row, b_b;
  };
  __device__ __forceinline__ static uint64_t* tma_barriers(double* smem) {
    return reinterpret_cast<uint64_t*>(smem + SMEM_DOUBLES);
  }
  static constexpr int TMA_SMEM_BYTES = SMEM_BYTES + 64;
  __device__ __forceinline__ static void tma_init(double* smem) {
    if (threadIdx.x == 0) {
      uint64_t* full = tma_barriers(smem);
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full + s, 1);                         // "stage filled": one expect_tx arrival
        mbar_init(full + STAGES + s, GEMM_THREADS / 32);  // "stage drained": one arrival per warp
      }
      mbar_init_fence();
    }
    __syncthreads();
  }
  __device__ __forceinline__ static void tma_issue(const TmaTile& t, int kt, int it, double* smem) {
    double* st = smem + (it % STAGES) * STAGE_DOUBLES;
    uint64_t* bar = tma_barriers(smem) + (it % STAGES);
    mbar_expect_tx(bar, STAGE_BYTES);
    if (A_KMAJ) tma_load_3d(st, t.mA, bar, t.a_in + kt * BK, t.a_row, t.a_b);
    else tma_load_3d(st, t.mA, bar, t.a_in, t.a_row + kt * BK, t.a_b);
    if (B_KMAJ) tma_load_3d(st + TA::SIZE, t.mB, bar, t.b_in + kt * BK, t.b_row, t.b_b);
    else tma_load_3d(st + TA::SIZE, t.mB, bar, t.b_in, t.b_row + kt * BK, t.b_b);
  }
  __device__ __forceinline__ static void tma_prologue(const TmaTile& t, int ktiles, int it0, double* smem) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s)
        if (s < ktiles) tma_issue(t, s, it0 + s, smem);
    }
  }
  // acc += A B over ``ktiles`` k-tiles; tma_prologue(t, ktiles, it0) must have been called.
  // ``mode`` 1 (default, "free run"): no CTA-wide barrier per k-tile -- every warp signals "stage
  // drained" on the stage's second mbarrier and only the producing thread waits for it before
  // refilling, so the warps drift up to STAGES-1 k-tiles apart and the DMMA pipe is not emptied at
  // every k-tile (measured on the sweep: 29.3 k -> 30.4 k samples/s).  ``mode`` 0 (OCBO_TMA_FREE=0)
  // keeps the barrier.
  //
  // WHERE the "drained" arrival sits matters.  Placed right after the k-tile's DMMAs in the source,
  // ptxas hoists it above them, directly behind the ISSUE of the last fragment loads and with no
  // scoreboard wait on them (SASS: SYNCS.ARRIVE two slots after the last LDS.64) -- and the refill
  // by the TMA engine then races with loads still in flight: wrong factors in some builds, right
  // ones in others (the parity check in bench.py and three GPU tests caught it).  The arrival for
  // k-tile it-1 is therefore the FIRST instruction of k-tile it, across the loop's back edge, where
  // no scheduler moves it: every DMMA of k-tile it-1 has ISSUED by then (in-order issue), hence
  // every fragment load it consumed has returned.  The last k-tile's arrival follows the closing
  // CTA barrier.
  __device__ __forceinline__ static void tma_mainloop(const TmaTile& t, int ktiles, int it0, double* smem,
                                                      Acc& acc, int mode = 1) {
    const bool free_run = mode != 0;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, tt = lane & 3;
    uint64_t* full = tma_barriers(smem);
    for (int kt = 0; kt < ktiles; ++kt) {
      const int it = it0 + kt;
#ifndef OCBO_ARRIVE_AT_END
      if (kt > 0) {  // this warp is done with k-tile it-1 (first thing after the loop's back edge)
        __syncwarp();
        if (lane == 0) mbar_arrive(full + STAGES + ((it - 1) % STAGES));
      }
#endif
      mbar_wait(full + (it % STAGES), (uint32_t)((it / STAGES) & 1));
      if (!free_run) __syncthreads();  // every warp is past k-tile kt-1: its stage may be refilled
      if (tid == 0 && kt + STAGES - 1 < ktiles) {
        if (free_run && it >= 1)  // the stage of k-tile it-1 has been drained by all warps
          mbar_wait(full + STAGES + ((it - 1) % STAGES), (uint32_t)(((it - 1) / STAGES) & 1));
        tma_issue(t, kt + STAGES - 1, it + STAGES - 1, smem);
      }
      const double* sa = smem + (it % STAGES) * STAGE_DOUBLES;
      const double* sb = sa + TA::SIZE;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = TA::frag(sa, wm * WM + mi * 8 + g, kk * 4 + tt);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = TB::frag(sb, wn * WN + ni * 8 + g, kk * 4 + tt);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
#ifdef OCBO_ARRIVE_AT_END
      // the placement that FAILS (tests/tools/arrive_placement.py builds this variant to show it):
      // straight-line behind the DMMAs, ptxas schedules the arrival above them
      __syncwarp();
      if (lane == 0) mbar_arrive(full + STAGES + (it % STAGES));
#endif
    }
    __syncthreads();  // every warp is done with the ring: smem may be reused
    // (the drained barriers count k-tiles in both modes: their phases stay in step over the passes
    // of a kernel that computes several tiles)
#ifndef OCBO_ARRIVE_AT_END
    if (ktiles > 0 && lane == 0) mbar_arrive(full + STAGES + ((it0 + ktiles - 1) % STAGES));
#endif
  }

  __device__ __forceinline__ static void zero(Acc& acc) {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }

  // f(row_in_tile, col_in_tile (even), v0, v1): v0 at (row, col), v1 at (row, col+1);
  // v0 / v1 are references into the accumulators.
  template <class F>
  __device__ __forceinline__ static void foreach(Acc& acc, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
        f(wm * WM + mi * 8 + g, wn * WN + ni * 8 + 2 * t, acc[mi][ni][0], acc[mi][ni][1]);
  }

  // acc = -C  (all loads issued back to back: one memory latency, overlapped with
  // the cp.async prologue); with the DMMA accumulating P*Q on top, the epilogue
  // stores -acc = C - P*Q without ever reading C again.
  __device__ __forceinline__ static void load_negated(Acc& acc, const double* C, int ldc) {
    foreach(acc, [&](int r, int c, double& v0, double& v1) {
      const double2 o = *reinterpret_cast<const double2*>(C + (size_t)r * ldc + c);
      v0 = -o.x;
      v1 = -o.y;
    });
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

// Same for 2:1 tiles (row tile = 2 column tiles wide): row ti holds the column
// tiles tj = 0 .. 2*ti+1, i.e. everything up to the end of the diagonal block.
__device__ __forceinline__ void tri_index_2to1(int t, int& ti, int& tj) {
  int i = (int)((sqrt(4.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) <= t) ++i;
  while (i * (i + 1) > t) --i;
  ti = i;
  tj = t - i * (i + 1);
}

// Banded order over the same 2:1 lower-triangular tile set, for L2 reuse: the row tiles are
// cut into bands of RB; a band is walked column by column (all its rows for column tile 0, then
// column tile 1, ...), so the CTAs resident at one time share the band's RB row panels of P
// (RB * 128 * k * 8 bytes: 17 MB at k = 2048) and stream each column panel of Q once per BAND.
// Row-major order streamed it once per ~3 row tiles (everything resident at a time): ncu showed
// 1.04 GB of DRAM reads for a rank-2048 update whose operands are 250 MB.
//   band g = rows [r0, r1): rectangle  tj in [0, 2 r0 + 2)  x  rows [r0, r1)   (column-major)
//                           triangle   rows r0+1 .. r1-1, tj in [2 r0 + 2, 2 ti + 1]
template <int RB>
__device__ __forceinline__ void tri_index_2to1_banded(int t, int nrows, int& ti, int& tj) {
  int r0 = 0;
  for (;;) {
    const int r1 = min(r0 + RB, nrows);
    const int count = (r1 - r0) * (r0 + r1 + 1);
    if (t < count || r1 >= nrows) {
      const int rows = r1 - r0, rect = rows * (2 * r0 + 2);
      if (t < rect) {
        tj = t / rows;
        ti = r0 + t % rows;
      } else {
        int u, v;
        tri_index_2to1(t - rect, u, v);
        ti = r0 + 1 + u;
        tj = 2 * r0 + 2 + v;
      }
      return;
    }
    t -= count;
    r0 = r1;
  }
}

}  // namespace ocbo
