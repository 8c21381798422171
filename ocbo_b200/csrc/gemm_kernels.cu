// DMMA tile kernels: C (-)= P Q^T, C (-)= P R, Sigma = K(Xa,Xb) - V1^T V2.
// 128 x 64 tiles, 256 threads, two CTAs per SM (see gemm_core.cuh).
#include <cstdlib>
#include <cstring>

#include "gemm_core.cuh"
#include "kernels.h"

namespace ocbo {

constexpr int TRI_BAND = 8;  // row tiles per band of the triangular tile order (gemm_core.cuh)

// ---------------------------------------------------------------------------
// C_tile (mode 0: =, mode 1: -=)  P_tile * Q_tile^T
//   P: rows follow C rows, k contiguous;  Q: rows follow C cols, k contiguous.
// <128,64>: Cholesky rank-k updates (P = Q = L panel, mode 1).
// <64,128>: panel solve X <- X * inv(L_jj)^T in place (P = X itself, Q = inv(L_jj),
//           mode 0): the CTA owns complete rows of the 128-wide panel, so reading
//           P over the whole k range and then overwriting it is race free.
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, bool TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ, NtParams p) {
  using G = Gemm<BM, BN, true, true, WARPS_M, WARPS_N>;
  extern __shared__ __align__(128) double smem[];
  const int b = blockIdx.z;
  if (p.info && p.info[b] != 0) return;
  int ti, tj;
  if (p.tri) {
    tri_index_2to1_banded<TRI_BAND>(blockIdx.x, p.ntc / (TILE / TN), ti, tj);
  } else {
    ti = blockIdx.x / p.ntc;
    tj = blockIdx.x % p.ntc;
    if (p.lower_skip && p.row0 + ti * BM + BM - 1 < p.col0 + tj * BN) return;
  }
  const double* P = p.P + (size_t)b * p.sP + (size_t)ti * BM * p.ldp;
  const double* Q = p.Q + (size_t)b * p.sQ + (size_t)tj * BN * p.ldq;
  double* C = p.C + (size_t)b * p.sC + (size_t)ti * BM * p.ldc + (size_t)tj * BN;
  const int ldc = p.ldc;

  typename G::Acc acc;
  // operand tiles by TMA (one thread, mbarrier transaction counts) or by cp.async (all threads)
  typename G::TmaTile tile{&mapP, 0, ti * BM, p.sP ? b : 0, &mapQ, 0, tj * BN, p.sQ ? b : 0};
  if (TMA) {
    G::tma_init(smem);
    G::tma_prologue(tile, p.ktiles, 0, smem);
  } else {
    G::prologue(P, p.ldp, Q, p.ldq, p.ktiles, smem);
  }
  if (p.mode == 0) G::zero(acc);
  else G::load_negated(acc, C, ldc);
  if (TMA) { if (p.tma_free) G::tma_mainloop(tile, p.ktiles, 0, smem, acc, 1); else G::tma_mainloop(tile, p.ktiles, 0, smem, acc, 0); }
  else G::mainloop(P, p.ldp, Q, p.ldq, p.ktiles, smem, acc);
  if (p.mode == 0) {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(v0, v1);
    });
  } else {
    G::foreach(acc, [&](int r, int c, double v0, double v1) {
      *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(-v0, -v1);
    });
  }
}

// ---------------------------------------------------------------------------
// C_tile (mode 0: =, 1: -=, 2: = rowvec + )  P_tile * R_tile
//   P: rows follow C rows, k contiguous;  R: k rows, cols follow C cols.
// Used for the left triangular solve (update + multiply by inv(L_II), in
// place: each CTA's R tile is its own C tile) and the sampling product
// F = mean + L Z.  tri_k: k stops at the diagonal block of the row tile;
// pair_rows: the CTA computes row tile y and row tile ntr-1-y, so every CTA
// of the triangular product carries the same amount of work.
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, bool TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_nn_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR, NnParams p) {
  using G = Gemm<BM, BN, true, false, WARPS_M, WARPS_N>;
  extern __shared__ __align__(128) double smem[];
  const int b = blockIdx.z;
  if (p.info && p.info[b] != 0) return;
  const int tj = blockIdx.x;
  const int ldc = p.ldc;
  const int npass = p.pair_rows ? 2 : 1;
  if (TMA) G::tma_init(smem);
  int it0 = 0;  // k-tiles consumed so far (mbarrier phase bookkeeping across the two passes)
  for (int pass = 0; pass < npass; ++pass) {
    const int ti = (pass == 0) ? (int)blockIdx.y : p.ntr - 1 - (int)blockIdx.y;
    if (pass == 1 && ti <= (int)blockIdx.y) break;
    const double* P = p.P + (size_t)b * p.sP + (size_t)ti * BM * p.ldp;
    const double* R = p.R + (size_t)b * p.sR + (size_t)tj * BN;
    double* C = p.C + (size_t)b * p.sC + (size_t)ti * BM * ldc + (size_t)tj * BN;
    int ktiles = p.ktiles;
    if (p.tri_k) ktiles = min(ktiles, (ti + 1) * (BM / BK));

    typename G::Acc acc;
    typename G::TmaTile tile{&mapP, 0, ti * BM, p.sP ? b : 0, &mapR, tj * BN, 0, p.sR ? b : 0};
    if (TMA) G::tma_prologue(tile, ktiles, it0, smem);
    else G::prologue(P, p.ldp, R, p.ldr, ktiles, smem);
    if (p.mode == 1) G::load_negated(acc, C, ldc);
    else G::zero(acc);
    if (TMA) { if (p.tma_free) G::tma_mainloop(tile, ktiles, it0, smem, acc, 1); else G::tma_mainloop(tile, ktiles, it0, smem, acc, 0); }
    else G::mainloop(P, p.ldp, R, p.ldr, ktiles, smem, acc);
    it0 += ktiles;
    if (p.mode == 0) {
      G::foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(v0, v1);
      });
    } else if (p.mode == 1) {
      G::foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(-v0, -v1);
      });
    } else {
      const double* rv = p.rowvec + (size_t)b * p.sRowvec + (size_t)ti * BM;
      G::foreach(acc, [&](int r, int c, double v0, double v1) {
        const double mu = rv[r];
        *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(mu + v0, mu + v1);
      });
    }
  }
}

// ---------------------------------------------------------------------------
// Fused sampling tail (north_star c, joint_mts.py:42-53): the tile of
//   F = mean + L Z      (L lower triangular, k stops at the row tile's diagonal block)
// is reduced in the epilogue to one (max, FIRST arg-max) per sample column over the tile's
// 128 rows, so the m x S sample matrix never exists in HBM.  128 x 32 tiles: S = 256 gives
// 8 column tiles x 32 row-tile pairs = 256 CTAs for ONE trial (the 128 x 64 product had 128
// CTAs on 296 slots: 65 % of the DMMA peak at batch 1); the accumulation order of every
// element is that of gemm_nn_kernel, so the reduced values are bit-identical to
// ocbo_sample + ocbo_segment_argmax.  np.argmax / np.max semantics incl. NaN (first NaN wins).
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, bool TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
sample_argmax_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR,
                     NnParams p) {
  using G = Gemm<BM, BN, true, false, WARPS_M, WARPS_N>;
  extern __shared__ __align__(128) double smem[];
  const int b = blockIdx.z;
  if (p.info && p.info[b] != 0) return;
  const int tj = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  if (TMA) G::tma_init(smem);
  int it0 = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int ti = (pass == 0) ? (int)blockIdx.y : p.ntr - 1 - (int)blockIdx.y;
    if (pass == 1 && ti <= (int)blockIdx.y) break;
    const double* P = p.P + (size_t)b * p.sP + (size_t)ti * BM * p.ldp;
    const double* R = p.R + (size_t)b * p.sR + (size_t)tj * BN;
    const int ktiles = min(p.ktiles, (ti + 1) * (BM / BK));

    typename G::Acc acc;
    typename G::TmaTile tile{&mapP, 0, ti * BM, p.sP ? b : 0, &mapR, tj * BN, 0, p.sR ? b : 0};
    G::zero(acc);
    if (TMA) {
      G::tma_prologue(tile, ktiles, it0, smem);
      { if (p.tma_free) G::tma_mainloop(tile, ktiles, it0, smem, acc, 1); else G::tma_mainloop(tile, ktiles, it0, smem, acc, 0); }  // ends with a barrier: the ring is free
    } else {
      G::prologue(P, p.ldp, R, p.ldr, ktiles, smem);
      G::mainloop(P, p.ldp, R, p.ldr, ktiles, smem, acc);
    }
    it0 += ktiles;

    const double* rv = p.rowvec + (size_t)b * p.sRowvec + (size_t)ti * BM;
    double* red_v = smem;                                   // [WARPS_M][BN]
    int* red_r = reinterpret_cast<int*>(smem + WARPS_M * BN);
#pragma unroll
    for (int ni = 0; ni < G::NI; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double bv = 0.0;
        int br = -1;
#pragma unroll
        for (int mi = 0; mi < G::MI; ++mi) {  // rows ascending: a later equal value never wins
          const int r = wm * G::WM + mi * 8 + g;
          const double v = rv[r] + acc[mi][ni][e];
          if (br < 0 || arg_better(v, r, bv, br)) {
            bv = v;
            br = r;
          }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {  // across the 8 row groups g (lane bits 2..4)
          const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int orr = __shfl_xor_sync(0xffffffffu, br, o);
          if (arg_better(ov, orr, bv, br)) {
            bv = ov;
            br = orr;
          }
        }
        if (g == 0) {
          const int c = wn * G::WN + ni * 8 + 2 * t + e;
          red_v[wm * BN + c] = bv;
          red_r[wm * BN + c] = br;
        }
      }
    __syncthreads();
    if (threadIdx.x < BN) {
      double bv = red_v[threadIdx.x];
      int br = red_r[threadIdx.x];
#pragma unroll
      for (int w = 1; w < WARPS_M; ++w) {
        const double ov = red_v[w * BN + threadIdx.x];
        const int orr = red_r[w * BN + threadIdx.x];
        if (arg_better(ov, orr, bv, br)) {
          bv = ov;
          br = orr;
        }
      }
      const size_t o = ((size_t)b * p.ntr + ti) * p.out_ld + (size_t)tj * BN + threadIdx.x;
      p.tile_max[o] = bv;
      p.tile_arg[o] = br;
    }
    __syncthreads();  // the reduction scratch is the ring of the next pass
  }
}

// ---------------------------------------------------------------------------
// C_tile = K(Xa_i, Xb_j) - sum_k V1[k][i] V2[k][j]     (posterior covariance)
//   V1: n x ma, V2: n x mb row-major (k rows).  sym: Xa == Xb, V1 == V2,
//   padding diagonal = 1; tri: lower-triangular tiles only (2:1 mapping).
// The Gram term is evaluated in the epilogue from points staged (pre-scaled by
// 1/bandwidth) in the drained cp.async ring, so K(Xs,Xs) never exists in HBM.
// ---------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, int KIND, bool TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
postcov_kernel(const __grid_constant__ CUtensorMap mapV1, const __grid_constant__ CUtensorMap mapV2, CovParams p) {
  using G = Gemm<BM, BN, false, false, WARPS_M, WARPS_N>;
  static_assert(G::SMEM_DOUBLES >= (MAXD + 1) * (BM + BN), "ring too small for the point staging");
  extern __shared__ __align__(128) double smem[];
  const int b = blockIdx.z;
  int ti, tj;
  if (p.tri) {
    tri_index_2to1_banded<TRI_BAND>(blockIdx.x, p.ntc / (TILE / TN), ti, tj);
  } else {
    ti = blockIdx.x / p.ntc;
    tj = blockIdx.x % p.ntc;
  }
  const double* V1 = p.V1 + (size_t)b * p.sV1 + (size_t)ti * BM;
  const double* V2 = p.V2 + (size_t)b * p.sV2 + (size_t)tj * BN;
  double* C = p.C + (size_t)b * p.sC + (size_t)ti * BM * p.ldc + (size_t)tj * BN;

  typename G::Acc acc;
  G::zero(acc);
  if (TMA) {
    typename G::TmaTile tile{&mapV1, ti * BM, 0, p.sV1 ? b : 0, &mapV2, tj * BN, 0, p.sV2 ? b : 0};
    G::tma_init(smem);
    G::tma_prologue(tile, p.ktiles, 0, smem);
    { if (p.tma_free) G::tma_mainloop(tile, p.ktiles, 0, smem, acc, 1); else G::tma_mainloop(tile, p.ktiles, 0, smem, acc, 0); }
  } else {
    G::prologue(V1, p.ldv1, V2, p.ldv2, p.ktiles, smem);
    G::mainloop(V1, p.ldv1, V2, p.ldv2, p.ktiles, smem, acc);
  }

  // stage scaled points: xa[d][BM], xb[d][BN]
  const int D = p.dim;
  const double* th = p.theta + (size_t)b * p.sTheta;
  const double scale = th[TH_SCALE];
  double* xa = smem;
  double* xb = smem + MAXD * BM;
  const double* Xa = p.Xa + (size_t)b * p.sXa + (size_t)ti * BM * D;
  const double* Xb = p.Xb + (size_t)b * p.sXb + (size_t)tj * BN * D;
  for (int e = threadIdx.x; e < BM * D; e += GEMM_THREADS) {
    const int r = e / D, d = e % D;
    xa[d * BM + r] = Xa[e] / th[TH_BW + d];
  }
  for (int e = threadIdx.x; e < BN * D; e += GEMM_THREADS) {
    const int r = e / D, d = e % D;
    xb[d * BN + r] = Xb[e] / th[TH_BW + d];
  }
  __syncthreads();
  // squared norms of the scaled points in numpy's summation order (dist2_np, common.cuh)
  double* na = smem + MAXD * (BM + BN);
  double* nb = na + BM;
  for (int r = threadIdx.x; r < BM + BN; r += GEMM_THREADS) {
    const double* xs = r < BM ? xa + r : xb + (r - BM);
    const int ld = r < BM ? BM : BN;
    double s = 0.0;
    for (int d = 0; d < D; ++d) s = sqnorm_np(s, xs[d * ld]);
    na[r] = s;
  }
  __syncthreads();
  const int mav = p.mav ? p.mav[b] : p.ma;
  const int mbv = p.mbv ? p.mbv[b] : p.mb;
  const int row0 = ti * BM, col0 = tj * BN;
  const int sym = p.sym, ldc = p.ldc;
  // Epilogue in three straight-line passes over the accumulators (KIND is a template
  // parameter): squared distances, then all 32 kernel values of the thread unconditionally
  // -- independent exp() chains the scheduler can interleave -- then padding select + store.
  // (With the kind and the padding tested inside one loop the chains ran one after another
  // and the epilogue cost about as much as 10 % of the k = 1024 main loop.)
  // (two halves of the warp tile at a time: 16 values in flight keep the kernel at 128 registers)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NH = 4, HM = G::MI / NH;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      double kv[HM][G::NI][2];
#pragma unroll
      for (int mi = 0; mi < HM; ++mi)
#pragma unroll
        for (int ni = 0; ni < G::NI; ++ni) {
          const int r = wm * G::WM + (h * HM + mi) * 8 + g, c = wn * G::WN + ni * 8 + 2 * t;
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            if (d < D) {
              const double xr = xa[d * BM + r];
              const double2 xc = *reinterpret_cast<const double2*>(xb + d * BN + c);
              d0 = fma(xr, xc.x, d0);
              d1 = fma(xr, xc.y, d1);
            }
          }
          const double2 ncol = *reinterpret_cast<const double2*>(nb + c);
          kv[mi][ni][0] = dist2_np(na[r], ncol.x, d0);
          kv[mi][ni][1] = dist2_np(na[r], ncol.y, d1);
        }
      if constexpr (KIND == OCBO_KERNEL_SE) {  // exp chains in lockstep, one row of fragments at a time
#pragma unroll
        for (int mi = 0; mi < HM; ++mi)
#pragma unroll
          for (int ni = 0; ni < G::NI; ni += 2) kern_eval_se<4>(scale, &kv[mi][ni][0]);  // 8 at a time would spill (128 registers)
      } else {
#pragma unroll
        for (int mi = 0; mi < HM; ++mi)
#pragma unroll
          for (int ni = 0; ni < G::NI; ++ni) {
            kv[mi][ni][0] = kern_eval(KIND, scale, kv[mi][ni][0]);
            kv[mi][ni][1] = kern_eval(KIND, scale, kv[mi][ni][1]);
          }
      }
#pragma unroll
      for (int mi = 0; mi < HM; ++mi)
#pragma unroll
        for (int ni = 0; ni < G::NI; ++ni) {
          const int r = wm * G::WM + (h * HM + mi) * 8 + g, c = wn * G::WN + ni * 8 + 2 * t;
          const int gi = row0 + r, gj = col0 + c;
          const bool vr = gi < mav;
          double o0 = (vr && gj < mbv) ? kv[mi][ni][0] - acc[h * HM + mi][ni][0] : 0.0;
          double o1 = (vr && gj + 1 < mbv) ? kv[mi][ni][1] - acc[h * HM + mi][ni][1] : 0.0;
          if (sym && gi == gj && !(vr && gj < mbv)) o0 = 1.0;
          if (sym && gi == gj + 1 && !(vr && gj + 1 < mbv)) o1 = 1.0;
          *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = make_double2(o0, o1);
        }
    }
  }
}

// ------------------------------- launchers ---------------------------------
template <class K>
static cudaError_t ensure_smem(K kernel, int bytes, bool& done) {
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) done = true;
  return e;
}

// OCBO_TMA=0 keeps every operand load on cp.async (the round-1 path), for A/B measurements
static bool use_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCBO_TMA");
    v = (e ? atoi(e) != 0 : 1) && tma_encoder() != nullptr;
  }
  return v != 0;
}

// OCBO_TMA_FREE=0 keeps a CTA-wide barrier per k-tile in the TMA main loop (A/B measurements)
static int tma_free_run() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCBO_TMA_FREE");
    v = e ? atoi(e) : 1;
  }
  return v;
}

cudaError_t launch_gemm_nt(const NtParams& p0, int ntiles, int batch, cudaStream_t st, int panel) {
  if (ntiles <= 0 || batch <= 0) return cudaSuccess;
  NtParams p = p0;
  p.tma_free = tma_free_run();
  CUtensorMap mP, mQ;
  memset(&mP, 0, sizeof(mP));
  memset(&mQ, 0, sizeof(mQ));
  if (panel) {
    using G = Gemm<TM / 2, TILE, true, true, 2, 4>;
    const uint64_t kext = (uint64_t)p.ktiles * BK;
    if (use_tma() && tma_make_map(&mP, p.P, kext, (int64_t)ntiles * (TM / 2), p.ldp, p.sP, batch, BK + 4, TM / 2) &&
        tma_make_map(&mQ, p.Q, kext, (int64_t)p.ntc * TILE, p.ldq, p.sQ, batch, BK + 4, TILE)) {
      static bool init = false;
      cudaError_t e = ensure_smem(gemm_nt_kernel<TM / 2, TILE, 2, 4, true>, G::TMA_SMEM_BYTES, init);
      if (e != cudaSuccess) return e;
      gemm_nt_kernel<TM / 2, TILE, 2, 4, true><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::TMA_SMEM_BYTES, st>>>(mP, mQ, p);
      return counted(cudaGetLastError());
    }
    static bool init = false;
    cudaError_t e = ensure_smem(gemm_nt_kernel<TM / 2, TILE, 2, 4, false>, G::SMEM_BYTES, init);
    if (e != cudaSuccess) return e;
    gemm_nt_kernel<TM / 2, TILE, 2, 4, false><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(mP, mQ, p);
    return counted(cudaGetLastError());
  }
  using G = Gemm<TM, TN, true, true, 4, 2>;
  // region extents in rows: lower-triangular launches enumerate R x 2R tiles of a square region
  const int64_t rowsP = p.tri ? (int64_t)(p.ntc / (TILE / TN)) * TILE : (int64_t)(ntiles / p.ntc) * TM;
  const int64_t rowsQ = (int64_t)p.ntc * TN;
  const uint64_t kext = (uint64_t)p.ktiles * BK;
  if (use_tma() && tma_make_map(&mP, p.P, kext, rowsP, p.ldp, p.sP, batch, BK + 4, TM) &&
      tma_make_map(&mQ, p.Q, kext, rowsQ, p.ldq, p.sQ, batch, BK + 4, TN)) {
    static bool init = false;
    cudaError_t e = ensure_smem(gemm_nt_kernel<TM, TN, 4, 2, true>, G::TMA_SMEM_BYTES, init);
    if (e != cudaSuccess) return e;
    gemm_nt_kernel<TM, TN, 4, 2, true><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::TMA_SMEM_BYTES, st>>>(mP, mQ, p);
    return counted(cudaGetLastError());
  }
  static bool init = false;
  cudaError_t e = ensure_smem(gemm_nt_kernel<TM, TN, 4, 2, false>, G::SMEM_BYTES, init);
  if (e != cudaSuccess) return e;
  gemm_nt_kernel<TM, TN, 4, 2, false><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(mP, mQ, p);
  return counted(cudaGetLastError());
}

// tensor maps of a C (op)= P R launch: P k-major (rows x k), R k rows x columns
template <int BN>
static bool nn_maps(const NnParams& p, int rows, int ntc, int batch, CUtensorMap* mP, CUtensorMap* mR) {
  memset(mP, 0, sizeof(*mP));
  memset(mR, 0, sizeof(*mR));
  const uint64_t kext = (uint64_t)p.ktiles * BK;
  return use_tma() && kext > 0 && tma_make_map(mP, p.P, kext, rows, p.ldp, p.sP, batch, BK + 4, TM) &&
         tma_make_map(mR, p.R, (uint64_t)ntc * BN, kext, p.ldr, p.sR, batch, BN + 4, BK);
}

cudaError_t launch_gemm_nn(const NnParams& p0, int grid_rows, int ntc, int batch, cudaStream_t st) {
  using G = Gemm<TM, TN, true, false, 4, 2>;
  if (grid_rows <= 0 || ntc <= 0 || batch <= 0) return cudaSuccess;
  NnParams p = p0;
  p.tma_free = tma_free_run();
  CUtensorMap mP, mR;
  const int rows = (p.pair_rows ? p.ntr : grid_rows) * TM;
  if (nn_maps<TN>(p, rows, ntc, batch, &mP, &mR)) {
    static bool init = false;
    cudaError_t e = ensure_smem(gemm_nn_kernel<TM, TN, 4, 2, true>, G::TMA_SMEM_BYTES, init);
    if (e != cudaSuccess) return e;
    gemm_nn_kernel<TM, TN, 4, 2, true><<<dim3(ntc, grid_rows, batch), GEMM_THREADS, G::TMA_SMEM_BYTES, st>>>(mP, mR, p);
    return counted(cudaGetLastError());
  }
  static bool init = false;
  cudaError_t e = ensure_smem(gemm_nn_kernel<TM, TN, 4, 2, false>, G::SMEM_BYTES, init);
  if (e != cudaSuccess) return e;
  gemm_nn_kernel<TM, TN, 4, 2, false><<<dim3(ntc, grid_rows, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(mP, mR, p);
  return counted(cudaGetLastError());
}

cudaError_t launch_sample_argmax(const NnParams& p0, int ntc, int batch, cudaStream_t st) {
  using G = Gemm<TM, 32, true, false, 4, 2>;
  if (ntc <= 0 || batch <= 0) return cudaSuccess;
  NnParams p = p0;
  p.tma_free = tma_free_run();
  CUtensorMap mP, mR;
  const dim3 grid(ntc, (p.ntr + 1) / 2, batch);
  if (nn_maps<32>(p, p.ntr * TM, ntc, batch, &mP, &mR)) {
    static bool init = false;
    cudaError_t e = ensure_smem(sample_argmax_kernel<TM, 32, 4, 2, true>, G::TMA_SMEM_BYTES, init);
    if (e != cudaSuccess) return e;
    sample_argmax_kernel<TM, 32, 4, 2, true><<<grid, GEMM_THREADS, G::TMA_SMEM_BYTES, st>>>(mP, mR, p);
    return counted(cudaGetLastError());
  }
  static bool init = false;
  cudaError_t e = ensure_smem(sample_argmax_kernel<TM, 32, 4, 2, false>, G::SMEM_BYTES, init);
  if (e != cudaSuccess) return e;
  sample_argmax_kernel<TM, 32, 4, 2, false><<<grid, GEMM_THREADS, G::SMEM_BYTES, st>>>(mP, mR, p);
  return counted(cudaGetLastError());
}

template <int KIND>
static cudaError_t launch_postcov_kind(const CovParams& p0, int ntiles, int batch, cudaStream_t st) {
  using G = Gemm<TM, TN, false, false, 4, 2>;
  CovParams p = p0;
  p.tma_free = tma_free_run();
  CUtensorMap m1, m2;
  memset(&m1, 0, sizeof(m1));
  memset(&m2, 0, sizeof(m2));
  const uint64_t kext = (uint64_t)p.ktiles * BK;
  if (kext > 0 && use_tma() && tma_make_map(&m1, p.V1, p.ma, kext, p.ldv1, p.sV1, batch, TM + 4, BK) &&
      tma_make_map(&m2, p.V2, p.mb, kext, p.ldv2, p.sV2, batch, TN + 4, BK)) {
    static bool init = false;
    cudaError_t e = ensure_smem(postcov_kernel<TM, TN, 4, 2, KIND, true>, G::TMA_SMEM_BYTES, init);
    if (e != cudaSuccess) return e;
    postcov_kernel<TM, TN, 4, 2, KIND, true><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::TMA_SMEM_BYTES, st>>>(m1, m2, p);
    return counted(cudaGetLastError());
  }
  static bool init = false;
  cudaError_t e = ensure_smem(postcov_kernel<TM, TN, 4, 2, KIND, false>, G::SMEM_BYTES, init);
  if (e != cudaSuccess) return e;
  postcov_kernel<TM, TN, 4, 2, KIND, false><<<dim3(ntiles, 1, batch), GEMM_THREADS, G::SMEM_BYTES, st>>>(m1, m2, p);
  return counted(cudaGetLastError());
}

cudaError_t launch_postcov(const CovParams& p, int ntiles, int batch, cudaStream_t st) {
  if (ntiles <= 0 || batch <= 0) return cudaSuccess;
  switch (p.kind) {
    case OCBO_KERNEL_SE: return launch_postcov_kind<OCBO_KERNEL_SE>(p, ntiles, batch, st);
    case OCBO_KERNEL_MATERN12: return launch_postcov_kind<OCBO_KERNEL_MATERN12>(p, ntiles, batch, st);
    case OCBO_KERNEL_MATERN32: return launch_postcov_kind<OCBO_KERNEL_MATERN32>(p, ntiles, batch, st);
    default: return launch_postcov_kind<OCBO_KERNEL_MATERN52>(p, ntiles, batch, st);
  }
}

}  // namespace ocbo
