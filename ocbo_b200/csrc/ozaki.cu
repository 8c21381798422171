// FP64 rank-k updates on the 5th-generation tensor cores (tcgen05.mma kind::i8), Ozaki-style:
//
//   P (rows x k, fp64) = scale_r * sum_{t=0..6} 2^(-8 t) s_t        s_t: int8 slices, row-wise exponent
//   P Q^T              = scale_i scale_j * sum_{g=0..6} 2^(-8 g) [ sum_{t+u=g} s_t(P) s_u(Q)^T ]
//
// The slices are the BALANCED radix-256 digits (in [-128, 127], the full int8 range) of the 57-bit integer
// rint(P / 2^E * 2^56), |P / 2^E| <= 0.498: 7 slices carry 56 bits below the row maximum, so 28 slice pairs do
// what 36 pairs of 7-bit slices (8 x 7 = 56 bits) would.  Pairs with t + u >= 7 are dropped: below 2^-55 of the
// product of the row maxima per term (measured <= 1e-14 |P||Q|^T element-wise on Cholesky panels; fp64's own
// bound for k = 2048 is 2e-13).  The integer products are EXACT: every group g is one int32 accumulator in
// tensor memory, |acc_g| <= (g + 1) k 2^14 < 2^31 for k <= 16384.
//
// Two kernels:
//   ozaki_split_kernel   fp64 panel -> int8 slices + per-row scale, written as the shared-memory image of the
//                        operand tiles (k blocks of 32 bytes, 32-byte swizzle applied), HBM bound
//   ozaki_update_kernel  persistent, warp specialised: one thread streams those images with contiguous bulk copies
//                        (cp.async.bulk: all 7 slices of a 128-row P tile in one 28 KiB copy, 7 x 2 KiB for the 64-row
//                        Q tile; 4 stages of 42 KiB on mbarrier transaction counts), one thread issues the 28
//                        slice-pair products of each k step into 7 accumulators of 128 x 64 int32 (448 TMEM
//                        columns), eight warps read the accumulators back (tcgen05.ld), recombine them in fp64
//                        (Horner in 2^-8, exact conversions) into registers, release tensor memory for the next
//                        tile, and apply C -= scale_i scale_j (...) through a shared-memory transposition
//                        (coalesced read-modify-write of C) while the next tile's products are issued.
// What bounds the main loop: shared-memory bandwidth.  Per k step the tensor core reads 10 x 4 KiB of P slices and
// 28 x 2 KiB of Q slices and the bulk copies write 42 KiB: 138 KiB at 128 B/clk = 1.1 k clocks against 0.9 k clocks of
// MMA issue (28 x 32) -- measured 1.19 k per k step (ncu: tensor pipe 57-62 %, tensor-core shared-memory wavefronts 61 %).
// A cta_group::2 pairing (each CTA reads half of the Q rows and receives half of the copies) would make it MMA bound.
// Why a fused kernel: with library int8 GEMMs the split / accumulate passes cost more than the GEMMs save
// (0.78x); here the 8 int32 products never leave the SM and every operand byte staged in shared
// memory feeds 28/14 as many MMAs as in a plain GEMM.
#include <cstdlib>

#include "common.cuh"
#include "gemm_core.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace ocbo {

namespace {
constexpr int OZ_S = 7;                      // slices per operand (8 bits each)
constexpr int OZ_BM = 128, OZ_BN = 64;       // C tile
constexpr int OZ_BK = 32;                    // bytes of k per stage = one 32-byte swizzle row = one tcgen05.mma
constexpr int OZ_STAGES = 4;
constexpr int OZ_A_SLICE = OZ_BM * OZ_BK, OZ_B_SLICE = OZ_BN * OZ_BK;
constexpr int OZ_A_STAGE = OZ_S * OZ_A_SLICE, OZ_B_STAGE = OZ_S * OZ_B_SLICE;
constexpr int OZ_STAGE_BYTES = OZ_A_STAGE + OZ_B_STAGE;                 // 42 KiB
constexpr int OZ_EPI_LD = 17;                                           // 16 columns + 1 pad (doubles)
constexpr int OZ_EPI_WARPS = 8;                                         // two per TMEM lane quarter: 32 columns each
constexpr int OZ_EPI_BYTES = OZ_EPI_WARPS * 32 * OZ_EPI_LD * 8;
constexpr int OZ_BAR_BYTES = 256;
constexpr int OZ_GRAM_D = 8;                                            // covariance mode: point dimension staged at most
// sampling mode: per-warp column maxima; covariance mode: scaled points and squared norms of the tile's rows / columns
constexpr int OZ_RED_BYTES = (OZ_BM + OZ_BN) * (OZ_GRAM_D + 1) * 8;
static_assert(OZ_RED_BYTES >= 2 * OZ_EPI_WARPS * 32 * (8 + 4), "scratch too small for the sampling reduction");
constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE_BYTES + OZ_EPI_BYTES + OZ_BAR_BYTES + OZ_RED_BYTES + 1024;
constexpr int OZ_THREADS = 64 + 32 * OZ_EPI_WARPS;  // warp 0: TMA, warp 1: MMA + TMEM allocation, warps 2..9: epilogue
constexpr uint32_t OZ_TMEM_COLS = 512;       // 7 accumulators x 64 columns (allocations are powers of two)

// instruction descriptor (cute/arch/mma_sm100_desc.hpp): D = S32, A = B = S8, both K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t oz_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
}

// K-major operand tile, 32-byte swizzle: 8-row groups 256 bytes apart (SBO), descriptor version 1, layout type 6
__device__ __forceinline__ uint64_t oz_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
// contiguous global -> shared bulk copy onto an mbarrier transaction count (SASS UBLKCP)
__device__ __forceinline__ void oz_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void oz_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// exact int32 -> fp64 without the conversion pipe: 2^52 + 2^31 + x as a bit pattern, minus the bias
__device__ __forceinline__ double oz_i2d(uint32_t x) {
  return __hiloint2double(0x43300000, (int)(x ^ 0x80000000u)) - 4503601774854144.0;
}

// Row exponent: amax = f 2^e, f in [0.5, 1).  E = e + 1 gives |P| 2^-E < 1/2; the balanced 7-digit radix-256
// range is +-0.498 2^0, so the top 0.4 % of f takes one more bit.  inv = 2^(56 - E), scale = 2^(E - 8).
__device__ __forceinline__ void oz_exponent(double amax, bool bad, double& inv, double& sc) {
  inv = 0.0;
  sc = 0.0;
  if (bad) {
    sc = CUDART_NAN;  // a NaN / Inf panel row poisons its rows and columns of C, like the fp64 update would
  } else if (amax >= 1e-290) {
    int e;
    const double f = frexp(amax, &e);
    const int E = e + (f > 0.996 ? 2 : 1);
    inv = ldexp(1.0, 56 - E);
    sc = ldexp(1.0, E - 8);
  }
}
// Balanced radix-256 digits of Y = rint(x 2^56) (|Y| <= 0.498 2^56), most significant first: s[0] .. s[6],
// Y = sum_t s[t] 256^(6 - t), every s[t] in [-128, 127].
__device__ __forceinline__ void oz_digits(double y, uint32_t (&d)[OZ_S]) {
  long long Y = __double2ll_rn(y);
#pragma unroll
  for (int t = OZ_S - 1; t >= 0; --t) {
    const int b = (int)(signed char)(Y & 0xff);
    d[t] = (uint32_t)b & 0xffu;
    Y = (Y - b) >> 8;
  }
}

struct OzParams {
  double* C; int ldc; int64_t sC;
  const int8_t* S; int RB, KB;  // slice workspace (layout below), its 128-row blocks and 32-byte k blocks
  const int8_t* SB; int RBb;    // the Q operand's slices (the same workspace for the symmetric updates)
  const double* scaleA; const double* scaleB; int64_t sScale, sScaleB;  // per slice row; batch strides
  // sampling mode: max / first arg-max over the 128 rows of mean + P Q^T, per column
  const double* mean; int64_t sMean; double* tile_max; int32_t* tile_arg; int out_ld;
  // covariance mode: C = K(x_i, x_j) - P Q^T with the Gram term evaluated in the epilogue
  const double* X; int64_t sX; int dim, kind; const double* theta; int64_t sTheta;
  int rowA0, rowB0;      // slice rows of the region's first row (multiple of 128) / first column (of 64)
  int R, ntc, ntiles;    // tri: row tiles of the square region; rect: column tiles (64 wide); tiles per problem
  int ktiles;            // k / 32
  int kb0;               // first k block of the product inside the workspace (0: the whole k range)
  int tri, lower_skip, row0, col0;
  const int32_t* info;
  int total;             // ntiles * batch
};

// Work item w -> (problem, tile, k steps); false: nothing to do (parked problem, tile above the diagonal).  Every
// role evaluates this identically, so skipped items never touch a pipeline.
// MODE 0: C -= P Q^T.  MODE 2: C = K(x_i, x_j) - P Q^T (posterior covariance, same tiles).  MODE 1 (sampling): tiles of mean + L Z with L lower triangular -- k stops at the row tile's
// diagonal block, so the tiles differ in length: items are numbered by DESCENDING length over all problems and
// dealt to the CTAs in snake order (round r: CTA c takes item r G + c, or r G + G - 1 - c when r is odd).
template <int MODE>
__device__ __forceinline__ bool oz_tile(const OzParams& p, int w, int& b, int& ti, int& tj, int& ktiles) {
  if (MODE == 1) {
    const int per_row = p.ntc * (p.total / p.ntiles);  // column tiles x problems
    ti = p.R - 1 - w / per_row;
    const int rem = w % per_row;
    b = rem / p.ntc;
    tj = rem - b * p.ntc;
    ktiles = (ti + 1) * (OZ_BM / OZ_BK);
    return !(p.info && p.info[b] != 0);
  }
  ktiles = p.ktiles;
  b = w / p.ntiles;
  const int t = w - b * p.ntiles;
  if (p.info && p.info[b] != 0) return false;
  if (p.tri) {
    tri_index_2to1_banded<8>(t, p.R, ti, tj);
  } else {
    ti = t / p.ntc;
    tj = t - ti * p.ntc;
    if (p.lower_skip && p.row0 + ti * OZ_BM + OZ_BM - 1 < p.col0 + tj * OZ_BN) return false;
  }
  return true;
}
template <int MODE>
__device__ __forceinline__ int oz_item(int round) {
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  return round * G + ((MODE == 1 && (round & 1)) ? G - 1 - c : c);
}

template <int MODE>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_update_kernel(OzParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the shared array: rounding the pointer through an integer loses the address
  // space, and every epilogue access to shared memory became a generic LD / ST (ncu: the Gram epilogue's dot products
  // waited on them)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* epi = reinterpret_cast<double*>(smem + OZ_STAGES * OZ_STAGE_BYTES);
  double* red_v = reinterpret_cast<double*>(smem + OZ_STAGES * OZ_STAGE_BYTES + OZ_EPI_BYTES + OZ_BAR_BYTES);  // [2][8][32]
  int* red_r = reinterpret_cast<int*>(red_v + 2 * OZ_EPI_WARPS * 32);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OZ_STAGES * OZ_STAGE_BYTES + OZ_EPI_BYTES);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* tmem_full = empty + OZ_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, OZ_EPI_WARPS);
    mbar_init_fence();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(OZ_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {  // ---- TMA producer
    int it = 0;
    for (int round = 0;; ++round) {
      const int w = oz_item<MODE>(round);
      if (round * (int)gridDim.x >= p.total) break;
      if (w >= p.total) continue;
      int b, ti, tj, ktiles;
      if (!oz_tile<MODE>(p, w, b, ti, tj, ktiles)) continue;
      // block (k block kt, row block rb) of the workspace = [slice][128 rows][32 bytes], already swizzled: the A tile
      // is ONE contiguous 28 KiB copy, the B tile 7 copies of 2 KiB (64 of the 128 rows of every slice)
      const int rbA = p.rowA0 / OZ_BM + ti, rB = p.rowB0 + tj * OZ_BN;
      const int8_t* srcA = p.S + (((size_t)b * p.KB + p.kb0) * p.RB + rbA) * OZ_A_STAGE;
      const int8_t* srcB = p.SB + (((size_t)b * p.KB + p.kb0) * p.RBb + rB / OZ_BM) * OZ_A_STAGE + (size_t)(rB % OZ_BM) * OZ_BK;
      const size_t kstep = (size_t)p.RB * OZ_A_STAGE, kstepB = (size_t)p.RBb * OZ_A_STAGE;
      for (int kt = 0; kt < ktiles; ++kt, ++it) {
        const int s = it % OZ_STAGES;
        if (it >= OZ_STAGES) mbar_wait(empty + s, (uint32_t)(((it / OZ_STAGES) - 1) & 1));
        mbar_expect_tx(full + s, OZ_STAGE_BYTES);
        uint8_t* dst = smem + s * OZ_STAGE_BYTES;
        oz_bulk(dst, srcA + kt * kstep, OZ_A_STAGE, full + s);
#pragma unroll
        for (int u = 0; u < OZ_S; ++u)
          oz_bulk(dst + OZ_A_STAGE + u * OZ_B_SLICE, srcB + kt * kstepB + (size_t)u * OZ_A_SLICE, OZ_B_SLICE, full + s);
      }
    }
  } else if (threadIdx.x == 32) {  // ---- MMA issuer
    int it = 0, done = 0;
    for (int round = 0;; ++round) {
      const int w = oz_item<MODE>(round);
      if (round * (int)gridDim.x >= p.total) break;
      if (w >= p.total) continue;
      int b, ti, tj, ktiles;
      if (!oz_tile<MODE>(p, w, b, ti, tj, ktiles)) continue;
      if (done > 0) {  // the epilogue has drained the previous tile's accumulators
        mbar_wait(tmem_empty, (uint32_t)((done - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int kt = 0; kt < ktiles; ++kt, ++it) {
        const int s = it % OZ_STAGES;
        mbar_wait(full + s, (uint32_t)((it / OZ_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da0 = oz_desc(smem_u32(smem + s * OZ_STAGE_BYTES));
        const uint64_t db0 = oz_desc(smem_u32(smem + s * OZ_STAGE_BYTES + OZ_A_STAGE));
        // Slice t of P meets slices u = 0 .. 6 - t of Q, whose results are the accumulators t .. 6 -- adjacent in
        // tensor memory (64 columns each), as the Q slices are adjacent in shared memory (64 rows each).  One MMA
        // of N = 64 L therefore serves L pairs (N <= 256: two MMAs when L > 4) and reads the P slice ONCE: with one
        // N = 64 MMA per pair the kernel was bound by the tensor core's shared-memory reads (ncu: 48 wavefronts per
        // MMA of 32 cycles).  10 MMAs per k step instead of 28.
#pragma unroll
        for (int t = 0; t < OZ_S; ++t) {
          const int L = OZ_S - t, L0 = L > 4 ? 4 : L;
          const uint64_t da = da0 + (uint64_t)((t * OZ_A_SLICE) >> 4);
          // the first k step's t = 0 MMAs cover all 7 accumulators and overwrite them
          oz_mma(tmem + (uint32_t)(t * OZ_BN), da, db0, oz_idesc(L0 * OZ_BN), (uint32_t)((kt | t) != 0));
          if (L > 4)
            oz_mma(tmem + (uint32_t)((t + 4) * OZ_BN), da, db0 + (uint64_t)((4 * OZ_B_SLICE) >> 4),
                   oz_idesc((L - 4) * OZ_BN), (uint32_t)((kt | t) != 0));
        }
        oz_commit(empty + s);
      }
      oz_commit(tmem_full);
      ++done;
    }
  } else if (warp >= 2) {  // ---- epilogue: TMEM -> fp64 recombination -> C
    // Warp = (TMEM lane quarter q, column half): thread = one row of the tile x 32 columns.  Phase (a) reads the 7
    // accumulators of those columns and recombines them into 32 doubles in registers, then hands tensor memory back
    // to the MMA thread; phase (b), the coalesced read-modify-write of C through a shared-memory transposition,
    // runs beside the NEXT tile's main loop.
    const int q = warp & 3, half = (warp - 2) >> 2;
    double* ebuf = epi + (warp - 2) * 32 * OZ_EPI_LD;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
    const int rr = lane >> 4, cc = lane & 15;
    int done = 0;
    for (int round = 0;; ++round) {
      const int w = oz_item<MODE>(round);
      if (round * (int)gridDim.x >= p.total) break;
      if (w >= p.total) continue;
      int b, ti, tj, ktiles;
      if (!oz_tile<MODE>(p, w, b, ti, tj, ktiles)) continue;
      const double sr = p.scaleA[(size_t)b * p.sScale + p.rowA0 + ti * OZ_BM + q * 32 + lane];
      const double* scB = p.scaleB + (size_t)b * p.sScaleB + p.rowB0 + tj * OZ_BN + half * 32;
      double* Ct = MODE != 1 ? p.C + (size_t)b * p.sC + (size_t)(ti * OZ_BM + q * 32) * p.ldc + (size_t)tj * OZ_BN + half * 32
                             : nullptr;
      double acc[32];
      mbar_wait(tmem_full, (uint32_t)(done & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t v[OZ_S][8];
#pragma unroll
        for (int g = 0; g < OZ_S; ++g) oz_ld8(tq + (uint32_t)(g * OZ_BN + c8 * 8), v[g]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double x = oz_i2d(v[OZ_S - 1][j]);
#pragma unroll
          for (int g = OZ_S - 2; g >= 0; --g) x = fma(x, 0.00390625, oz_i2d(v[g][j]));
          acc[c8 * 8 + j] = x * sr;
        }
      }
      // every accumulator column of this warp has been read: hand TMEM back to the MMA thread
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      if (MODE == 2) {
        // Gram term in the epilogue (K(Xs, Xs) never exists in HBM): the tile's points, scaled by 1 / bandwidth, and
        // their squared norms are staged once per tile by the eight epilogue warps; Dragonfly's expanded-form
        // distance in numpy's operation order (common.cuh: sqnorm_np / dist2_np), like every Gram kernel here
        const int D = p.dim;
        const double* th = p.theta + (size_t)b * p.sTheta;
        double* xa = red_v;                          // [D][128]
        double* xb = xa + OZ_GRAM_D * OZ_BM;         // [D][64]
        double* na = xb + OZ_GRAM_D * OZ_BN;         // [128] then [64]
        asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory");  // the previous tile's points are dead
        {
          const int e = threadIdx.x - 64;  // 0 .. 255
          if (e < OZ_BM + OZ_BN) {
            const bool isa = e < OZ_BM;
            const int r = isa ? e : e - OZ_BM;
            const double* x = p.X + (size_t)b * p.sX + (size_t)((isa ? ti * OZ_BM : tj * OZ_BN) + r) * D;
            double* dst = isa ? xa + r : xb + r;
            const int ld = isa ? OZ_BM : OZ_BN;
            double xv[OZ_GRAM_D], bw[OZ_GRAM_D];  // all loads first: they are independent
#pragma unroll
            for (int d = 0; d < OZ_GRAM_D; ++d) {
              xv[d] = d < D ? x[d] : 0.0;
              bw[d] = d < D ? th[TH_BW + d] : 1.0;
            }
            double nrm = 0.0;
#pragma unroll
            for (int d = 0; d < OZ_GRAM_D; ++d)
              if (d < D) {
                const double v = xv[d] / bw[d];
                dst[d * ld] = v;
                nrm = sqnorm_np(nrm, v);
              }
            na[e] = nrm;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory");
        const double kscale = th[TH_SCALE];
        const double* nb = na + OZ_BM;
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ebuf[lane * OZ_EPI_LD + j] = acc[c16 * 16 + j];
          __syncwarp();
          const int col = half * 32 + c16 * 16 + cc;   // column within the tile
          const double sc = scB[c16 * 16 + cc];
          double xc[OZ_GRAM_D];
#pragma unroll
          for (int d = 0; d < OZ_GRAM_D; ++d) xc[d] = d < D ? xb[d * OZ_BN + col] : 0.0;
          const double ncol = nb[col];
          double* Cc = Ct + c16 * 16 + cc;
          // dimension-major: 16 independent FMAs between two dependent ones (a dependent fp64 operation costs ~35
          // clocks here, and the eight epilogue warps are all the latency hiding there is)
          double kv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) kv[i] = 0.0;
#ifndef OCBO_OZ_NODOT   // diagnostic variant: what do the shared-memory reads of the row points cost?
#pragma unroll
          for (int d = 0; d < OZ_GRAM_D; ++d)
            if (d < D) {
#pragma unroll
              for (int i = 0; i < 16; ++i) kv[i] = fma(xa[d * OZ_BM + q * 32 + i * 2 + rr], xc[d], kv[i]);
            }
#endif
#pragma unroll
          for (int i = 0; i < 16; ++i) kv[i] = dist2_np(na[q * 32 + i * 2 + rr], ncol, kv[i]);
#ifdef OCBO_OZ_NOEXP  // diagnostic variant (tests/tools/oz_gram_variants.py): what does the exp chain cost?
#pragma unroll
          for (int i = 0; i < 16; ++i) kv[i] = kscale - 0.5 * kv[i];
#else
          if (p.kind == OCBO_KERNEL_SE) {  // 16 exp chains in lockstep (common.cuh: exp_core), same bits as kern_eval
            kern_eval_se<8>(kscale, kv);       // 8 per call: 16 would spill at the kernel's 168 registers
            kern_eval_se<8>(kscale, kv + 8);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) kv[i] = kern_eval(p.kind, kscale, kv[i]);
          }
#endif
#pragma unroll
          for (int i = 0; i < 16; ++i) Cc[(size_t)(i * 2 + rr) * p.ldc] = fma(-ebuf[(i * 2 + rr) * OZ_EPI_LD + cc], sc, kv[i]);
          __syncwarp();
        }
      } else if (MODE == 0) {
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ebuf[lane * OZ_EPI_LD + j] = acc[c16 * 16 + j];
          __syncwarp();
          const double sc = scB[c16 * 16 + cc];
          double* Cc = Ct + c16 * 16 + cc;
          double cv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) cv[i] = Cc[(size_t)(i * 2 + rr) * p.ldc];
#pragma unroll
          for (int i = 0; i < 16; ++i) Cc[(size_t)(i * 2 + rr) * p.ldc] = fma(-ebuf[(i * 2 + rr) * OZ_EPI_LD + cc], sc, cv[i]);
          __syncwarp();
        }
      } else {
        // F = mean + scale_i scale_j (...) is never written: per column, max / FIRST arg-max (np.argmax's rule,
        // arg_better) over the warp's 32 rows, then over the four lane quarters through shared memory
        const double* mu = p.mean + (size_t)b * p.sMean + (size_t)ti * OZ_BM + q * 32;
        double* rv = red_v + ((done & 1) * OZ_EPI_WARPS + (warp - 2)) * 32;
        int* ra = red_r + ((done & 1) * OZ_EPI_WARPS + (warp - 2)) * 32;
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ebuf[lane * OZ_EPI_LD + j] = acc[c16 * 16 + j];
          __syncwarp();
          const double sc = scB[c16 * 16 + cc];
          double bv = 0.0;
          int br = -1;
#pragma unroll
          for (int i = 0; i < 16; ++i) {  // rows ascending: a later equal value never wins
            const int r = i * 2 + rr;
            const double v = fma(ebuf[r * OZ_EPI_LD + cc], sc, mu[r]);
            if (br < 0 || arg_better(v, q * 32 + r, bv, br)) {
              bv = v;
              br = q * 32 + r;
            }
          }
          const double ov = __shfl_xor_sync(0xffffffffu, bv, 16);
          const int orr = __shfl_xor_sync(0xffffffffu, br, 16);
          if (arg_better(ov, orr, bv, br)) {
            bv = ov;
            br = orr;
          }
          if (rr == 0) {
            rv[c16 * 16 + cc] = bv;
            ra[c16 * 16 + cc] = br;
          }
          __syncwarp();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory");  // the eight epilogue warps
        if (q == 2) {  // warps 2 and 6 (lane quarter 2): one per column half, lane = column
          double bv = 0.0;
          int br = -1;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {  // lane quarter qq is epilogue warp (qq + 2) % 4 + 4 * half
            const int ew = ((qq + 2) & 3) + 4 * half;
            const double v = red_v[((done & 1) * OZ_EPI_WARPS + ew) * 32 + lane];
            const int r = red_r[((done & 1) * OZ_EPI_WARPS + ew) * 32 + lane];
            if (br < 0 || arg_better(v, r, bv, br)) {
              bv = v;
              br = r;
            }
          }
          const size_t o = ((size_t)b * p.R + ti) * p.out_ld + (size_t)tj * OZ_BN + half * 32 + lane;
          p.tile_max[o] = bv;
          p.tile_arg[o] = br;
        }
      }
      ++done;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OZ_TMEM_COLS) : "memory");
}

// One warp per row, 8 rows per CTA: row maximum -> exponent, then the 7 balanced radix-256 digits of every element,
// written in the layout the update kernel copies verbatim into shared memory:
//   S[problem][k block of 32][row block of 128][slice][128 rows][32 bytes], 16-byte chunk c of row r at c ^ ((r >> 2) & 1)
//   (the 32-byte swizzle of the UMMA operand descriptor);  scale[problem][row] = 2^(E - 8) with |P / 2^E| <= 0.498.
// A lane reads 4 consecutive doubles with one 256-bit load (a warp instruction covers 1 KiB of the row: 8 full
// lines) and writes their digits as one 32-bit word per slice; 4 lanes fill a 16-byte chunk, 8 lanes a 32-byte row of
// a slice image.  (The first version -- 16 consecutive k per lane, 128-bit loads -- spent 32 LSU wavefronts per load
// instruction: ncu had it LSU-bound at 73 % with 2 TB/s of traffic.)
__device__ __forceinline__ void oz_ld4(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__global__ void __launch_bounds__(256)
ozaki_split_kernel(const double* __restrict__ P, int ldp, int64_t sP, int rows, int kw_full, int8_t* __restrict__ S,
                   double* __restrict__ scale, const int32_t* __restrict__ info, int tri) {
  const int b = blockIdx.y;
  if (info && info[b] != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int KB = kw_full / OZ_BK, RB = rows / OZ_BM;
  // tri: P is lower triangular (a Cholesky factor) -- only the k blocks up to the row's diagonal block exist
  const int kw = tri ? min(kw_full, (row / OZ_BM + 1) * OZ_BM) : kw_full;
  const double* src = P + (size_t)b * sP + (size_t)row * ldp + lane * 4;
  double amax = 0.0;
  bool bad = false;
#pragma unroll 4
  for (int k0 = 0; k0 + lane * 4 < kw; k0 += 128) {
    double v[4];
    oz_ld4(src + k0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double a = fabs(v[j]);
      bad |= !(a <= 1.7976931348623157e308);
      amax = fmax(amax, a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  bad = __any_sync(0xffffffffu, bad);
  double inv, sc;
  oz_exponent(amax, bad, inv, sc);
  if (lane == 0) scale[(size_t)b * rows + row] = sc;
  // lane -> (k block within the 128-k step, chunk, word of the chunk)
  const int kbl = lane >> 3, c = (lane >> 2) & 1, wq = lane & 3;
  int8_t* dst = S + (size_t)b * KB * RB * OZ_A_STAGE + (size_t)(row / OZ_BM) * OZ_A_STAGE + (size_t)(row % OZ_BM) * OZ_BK +
                ((c ^ ((row >> 2) & 1)) << 4) + (wq << 2) + (size_t)kbl * RB * OZ_A_STAGE;
  // two 128-k steps per iteration: both loads are issued before the digit chains start (the loop was bound by the
  // latency of one 256-bit load per iteration: ncu long-scoreboard stalls, 2.5 TB/s)
  for (int k0 = 0; k0 + lane * 4 < kw; k0 += 256) {
    double v[2][4];
    const bool two = k0 + 128 + lane * 4 < kw;
    oz_ld4(src + k0, v[0]);
    if (two) oz_ld4(src + k0 + 128, v[1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      uint32_t wd[OZ_S];
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) wd[t] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t dg[OZ_S];
        oz_digits(v[h][j] * inv, dg);
#pragma unroll
        for (int t = 0; t < OZ_S; ++t) wd[t] |= dg[t] << (8 * j);
      }
      int8_t* d = dst + (size_t)((k0 + h * 128) / OZ_BK) * RB * OZ_A_STAGE;
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) *reinterpret_cast<uint32_t*>(d + (size_t)t * OZ_A_SLICE) = wd[t];
    }
  }
}

// ---- one workspace for a whole Cholesky factor (ocbo_potrf_ws, "global" mode) --------------------------------------
// Row scales fixed BEFORE the factorisation from its diagonal: |L_ik| <= sqrt(A_ii) for every k, so
// E_i = exponent of sqrt(A_ii) bounds the whole row i of L.  With one scale per row for all panels the slices of
// different panels are compatible: every element of L is split ONCE, and the middle updates, the trailing updates
// and the sampling tail (mean + L Z, k = m) all read the same images.  The absolute truncation, 2^-57 2^E_i per
// element, is that of rounding A's own entries to fp64; a row whose elements exceed the bound (A not positive
// definite: pivot i is about to fail) gets a NaN scale, which poisons row / column i of the trailing matrix only --
// LAPACK's info (first failing pivot) is unchanged.
__global__ void __launch_bounds__(256)
ozaki_diag_scale_kernel(const double* __restrict__ A, int lda, int64_t sA, int n, double* __restrict__ scale,
                        const int32_t* __restrict__ info) {
  const int b = blockIdx.y;
  if (info && info[b] != 0) return;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double d = A[(size_t)b * sA + (size_t)i * lda + i];
  double inv, sc;
  // (a hair above sqrt(A_ii): the computed L_ik may exceed the exact bound by rounding)
  oz_exponent(sqrt(d) * 1.0000001, !(d > 0.0) || !(d <= 1.7976931348623157e308), inv, sc);
  scale[(size_t)b * n + i] = sc;
}

// Split of the block columns [kb0 * 32, kb0 * 32 + kw) of L, rows row0 .. row0 + rows - 1 (P points at that
// corner), with the given row scales, into the factor-wide layout S[problem][KB][RB][slice][128][32]; a row's
// k range ends at its diagonal block (tri).  Same lane mapping as ozaki_split_kernel.
__global__ void __launch_bounds__(256)
ozaki_split_fixed_kernel(const double* __restrict__ P, int ldp, int64_t sP, int rows, int kw_full, int row0, int kb0,
                         int n, int8_t* __restrict__ S, double* __restrict__ scale, const int32_t* __restrict__ info) {
  const int b = blockIdx.y;
  if (info && info[b] != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lrow = blockIdx.x * 8 + warp;
  if (lrow >= rows) return;
  const int row = row0 + lrow;
  const int KB = n / OZ_BK, RB = n / OZ_BM;
  const int kw = min(kw_full, (row / OZ_BM + 1) * OZ_BM - kb0 * OZ_BK);
  const double* src = P + (size_t)b * sP + (size_t)lrow * ldp + lane * 4;
  const double sc = scale[(size_t)b * n + row];
  const double inv = 281474976710656.0 / sc;  // 2^48 / 2^(E - 8) = 2^(56 - E), exact (NaN scale: NaN)
  const int kbl = lane >> 3, c = (lane >> 2) & 1, wq = lane & 3;
  int8_t* dst = S + ((size_t)b * KB + kb0) * RB * OZ_A_STAGE + (size_t)(row / OZ_BM) * OZ_A_STAGE +
                (size_t)(row % OZ_BM) * OZ_BK + ((c ^ ((row >> 2) & 1)) << 4) + (wq << 2) + (size_t)kbl * RB * OZ_A_STAGE;
  bool over = false;
  for (int k0 = 0; k0 + lane * 4 < kw; k0 += 256) {
    double v[2][4];
    const bool two = k0 + 128 + lane * 4 < kw;
    oz_ld4(src + k0, v[0]);
    if (two) oz_ld4(src + k0 + 128, v[1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      uint32_t wd[OZ_S];
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) wd[t] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t dg[OZ_S];
        const double y = v[h][j] * inv;
        over |= !(fabs(y) <= 35886843558709250.0);  // 0.49803 * 2^56 (the digits reach 127/255): out of range, or NaN
        oz_digits(y, dg);
#pragma unroll
        for (int t = 0; t < OZ_S; ++t) wd[t] |= dg[t] << (8 * j);
      }
      int8_t* d = dst + (size_t)((k0 + h * 128) / OZ_BK) * RB * OZ_A_STAGE;
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) *reinterpret_cast<uint32_t*>(d + (size_t)t * OZ_A_SLICE) = wd[t];
    }
  }
  over = __any_sync(0xffffffffu, over);
  if (over && lane == 0) scale[(size_t)b * n + row] = CUDART_NAN;
}

// The same split for an operand stored k-major the other way round (V: k rows x operand rows, row-major, as the
// posterior covariance's V^T V needs it): CTA = 32 operand rows (lane = row, loads of one k are 256 contiguous
// bytes), warp w handles the k blocks w, w + 8, ...; a thread converts 32 k values of its row and writes the
// full 32-byte row of every slice (the swizzle only orders the two 16-byte halves): 1 KiB per warp and slice.
__global__ void __launch_bounds__(256)
ozaki_split_t_kernel(const double* __restrict__ V, int ldv, int64_t sV, int rows, int kw, int8_t* __restrict__ S,
                     double* __restrict__ scale, const int32_t* __restrict__ info) {
  // grid = (operand rows / 32, k chunks, problems): with few operand rows (Z^T: 256) the k range is cut into
  // chunks so that the GPU is filled; every CTA then finds the row maxima over the WHOLE k range itself (the
  // operand is small and L2 resident) and converts only its chunk
  const int b = blockIdx.z;
  if (info && info[b] != 0) return;
  __shared__ double s_max[8][32];
  __shared__ unsigned s_bad[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 32 + lane;
  const int KB = kw / OZ_BK, RB = (rows + OZ_BM - 1) / OZ_BM;
  const double* src = V + (size_t)b * sV + row;
  double amax = 0.0;
  bool bad = false;
  for (int kb = warp; kb < KB; kb += 8) {
#pragma unroll 8
    for (int j = 0; j < OZ_BK; ++j) {
      const double a = fabs(src[(size_t)(kb * OZ_BK + j) * ldv]);
      bad |= !(a <= 1.7976931348623157e308);
      amax = fmax(amax, a);
    }
  }
  s_max[warp][lane] = amax;
  const unsigned badmask = __ballot_sync(0xffffffffu, bad);
  if (lane == 0) s_bad[warp] = badmask;
  __syncthreads();
  unsigned badrows = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    amax = fmax(amax, s_max[w][lane]);
    badrows |= s_bad[w];
  }
  double inv, sc;
  oz_exponent(amax, (badrows >> lane) & 1u, inv, sc);
  if (warp == 0 && blockIdx.y == 0) scale[(size_t)b * rows + row] = sc;
  const int kb_per = (KB + (int)gridDim.y - 1) / (int)gridDim.y;
  const int kb0 = (int)blockIdx.y * kb_per, kb1 = min(KB, kb0 + kb_per);
  int8_t* dst = S + (size_t)b * KB * RB * OZ_A_STAGE + (size_t)(row / OZ_BM) * OZ_A_STAGE + (size_t)(row % OZ_BM) * OZ_BK;
  const int flip = (row >> 2) & 1;
  for (int kb = kb0 + warp; kb < kb1; kb += 8) {
    uint32_t wd[OZ_S][8];
#pragma unroll
    for (int t = 0; t < OZ_S; ++t)
#pragma unroll
      for (int q = 0; q < 8; ++q) wd[t][q] = 0u;
#pragma unroll
    for (int j = 0; j < OZ_BK; ++j) {
      uint32_t dg[OZ_S];
      oz_digits(src[(size_t)(kb * OZ_BK + j) * ldv] * inv, dg);
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) wd[t][j >> 2] |= dg[t] << (8 * (j & 3));
    }
    int8_t* d = dst + (size_t)kb * RB * OZ_A_STAGE;
#pragma unroll
    for (int t = 0; t < OZ_S; ++t) {
      uint4* o = reinterpret_cast<uint4*>(d + (size_t)t * OZ_A_SLICE);
      o[flip] = make_uint4(wd[t][0], wd[t][1], wd[t][2], wd[t][3]);
      o[flip ^ 1] = make_uint4(wd[t][4], wd[t][5], wd[t][6], wd[t][7]);
    }
  }
}

// CTAs of the persistent update kernel (one per SM, it owns the SM's shared and tensor memory).  OCBO_OZAKI_SMS
// leaves some SMs to the latency-bound chain kernels of the other streams (diagonal blocks, panel solves).
int oz_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    const char* e = getenv("OCBO_OZAKI_SMS");
    if (e && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
  }
  return n;
}
}  // namespace

int ozaki_slice_pairs() { return OZ_S * (OZ_S + 1) / 2; }

static size_t oz_slices_bytes(int rows, int kw, int batch) {  // rows rounded up to whole 128-row blocks
  const size_t rb = (size_t)(rows + OZ_BM - 1) / OZ_BM;
  return (((size_t)batch * OZ_S * rb * OZ_BM * kw) + 255) & ~(size_t)255;
}
size_t ozaki_ws_bytes(int rows, int kw, int batch) {
  // slices, then the row scales (kept 256-byte aligned)
  return oz_slices_bytes(rows, kw, batch) + ((((size_t)batch * rows * sizeof(double)) + 255) & ~(size_t)255);
}

cudaError_t launch_ozaki_split(const double* P, int ldp, int64_t sP, int rows, int kw, void* ws, const int32_t* info,
                               int batch, cudaStream_t st, int transposed, int tri) {
  if (rows <= 0 || kw <= 0 || batch <= 0) return cudaSuccess;
  if ((kw & 31) || ((uintptr_t)ws & 255)) return cudaErrorInvalidValue;
  // row-major panels are read with 256-bit loads
  if (!transposed && ((rows & 127) || (ldp & 3) || (sP & 3) || ((uintptr_t)P & 31))) return cudaErrorInvalidValue;
  if (transposed && ((rows & 31) || (ldp & 1) || (sP & 1))) return cudaErrorInvalidValue;
  int8_t* S = reinterpret_cast<int8_t*>(ws);
  double* scale = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(ws) + oz_slices_bytes(rows, kw, batch));
  if (transposed) {  // P: kw rows x `rows` columns
    const int ctas = (rows / 32) * batch;
    int ks = ctas >= 148 ? 1 : (296 + ctas - 1) / ctas;
    if (ks > 16) ks = 16;
    if (ks > kw / OZ_BK / 8) ks = kw / OZ_BK / 8 > 0 ? kw / OZ_BK / 8 : 1;
    ozaki_split_t_kernel<<<dim3(rows / 32, ks, batch), 256, 0, st>>>(P, ldp, sP, rows, kw, S, scale, info);
  }
  else
    ozaki_split_kernel<<<dim3(rows / 8, batch), 256, 0, st>>>(P, ldp, sP, rows, kw, S, scale, info, tri);
  return counted(cudaGetLastError());
}

template <int MODE>
static cudaError_t oz_launch(const OzParams& p, cudaStream_t st) {
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(ozaki_update_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  const int grid = p.total < oz_sm_count() ? p.total : oz_sm_count();
  ozaki_update_kernel<MODE><<<grid, OZ_THREADS, OZ_SMEM, st>>>(p);
  return counted(cudaGetLastError());
}

// C (region) -= P Q^T from a split workspace: rows of the region = slice rows rowA0 .., columns = slice rows rowB0 ..
// tri: lower-triangular 128 x 64 tiles of a square region of R row tiles (banded order); otherwise R x ntc tiles,
// lower_skip drops tiles entirely above the diagonal (row0 / col0: global coordinates of the region origin).
cudaError_t launch_ozaki_update(const void* ws, int rows, int kw, int rowA0, int rowB0, double* C, int ldc, int64_t sC,
                                int R, int ntc, int tri, int lower_skip, int row0, int col0, const int32_t* info,
                                int batch, cudaStream_t st, const OzGram* gram) {
  if (R <= 0 || batch <= 0 || (!tri && ntc <= 0)) return cudaSuccess;
  if (kw > 16384) return cudaErrorInvalidValue;  // int32 accumulators
  const int8_t* S = reinterpret_cast<const int8_t*>(ws);
  const double* scale = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(ws) + oz_slices_bytes(rows, kw, batch));
  if ((rowA0 % OZ_BM) || (rowB0 % OZ_BN) || (kw % OZ_BK) || (rows % OZ_BM)) return cudaErrorInvalidValue;
  OzParams p{};
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.S = S; p.RB = rows / OZ_BM; p.KB = kw / OZ_BK;
  p.SB = S; p.RBb = p.RB;
  p.scaleA = scale; p.scaleB = scale; p.sScale = rows; p.sScaleB = rows;
  p.rowA0 = rowA0; p.rowB0 = rowB0;
  p.R = R; p.ntc = ntc; p.ntiles = tri ? R * (R + 1) : R * ntc;
  p.ktiles = kw / OZ_BK;
  p.tri = tri; p.lower_skip = lower_skip; p.row0 = row0; p.col0 = col0;
  p.info = info;
  p.total = p.ntiles * batch;
  if (gram) {  // C = K(X, X) - P Q^T on the same tiles (rows / columns of the region = points rowA0 .. / rowB0 ..)
    if (gram->dim < 1 || gram->dim > OZ_GRAM_D || rowA0 != 0 || rowB0 != 0) return cudaErrorInvalidValue;
    p.X = gram->X; p.sX = gram->sX; p.dim = gram->dim; p.kind = gram->kind;
    p.theta = gram->theta; p.sTheta = gram->sTheta;
    return oz_launch<2>(p, st);
  }
  return oz_launch<0>(p, st);
}
int ozaki_gram_max_dim() { return OZ_GRAM_D; }

// ---- factor-wide workspace (see ozaki_diag_scale_kernel): layout of ozaki_ws_bytes(n, n, batch) ------------------
cudaError_t launch_ozaki_diag_scale(const double* A, int lda, int64_t sA, int n, void* ws, const int32_t* info, int batch,
                                    cudaStream_t st) {
  double* scale = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(ws) + oz_slices_bytes(n, n, batch));
  ozaki_diag_scale_kernel<<<dim3((n + 255) / 256, batch), 256, 0, st>>>(A, lda, sA, n, scale, info);
  return counted(cudaGetLastError());
}

// P = &L[row0][kb0 * 32]: rows x kw block of the factor (whole block columns, row0 and kw multiples of 128)
cudaError_t launch_ozaki_split_fixed(const double* P, int ldp, int64_t sP, int rows, int kw, int row0, int kb0, int n,
                                     void* ws, const int32_t* info, int batch, cudaStream_t st) {
  if (rows <= 0 || kw <= 0 || batch <= 0) return cudaSuccess;
  if ((kw & 127) || (rows & 127) || (row0 & 127) || (n & 127) || (ldp & 3) || (sP & 3) || ((uintptr_t)P & 31) ||
      ((uintptr_t)ws & 255))
    return cudaErrorInvalidValue;
  int8_t* S = reinterpret_cast<int8_t*>(ws);
  double* scale = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(ws) + oz_slices_bytes(n, n, batch));
  ozaki_split_fixed_kernel<<<dim3(rows / 8, batch), 256, 0, st>>>(P, ldp, sP, rows, kw, row0, kb0, n, S, scale, info);
  return counted(cudaGetLastError());
}

// C (region) -= L[rowA0.., k range] L[rowB0.., k range]^T from the factor-wide workspace; k range = nkb blocks of 32
// from kb0; rowA0 / rowB0 are rows of the factor.
cudaError_t launch_ozaki_update_range(const void* ws, int n, int kb0, int nkb, int rowA0, int rowB0, double* C, int ldc,
                                      int64_t sC, int R, int ntc, int tri, int lower_skip, int row0, int col0,
                                      const int32_t* info, int batch, cudaStream_t st) {
  if (R <= 0 || batch <= 0 || nkb <= 0 || (!tri && ntc <= 0)) return cudaSuccess;
  if ((rowA0 % OZ_BM) || (rowB0 % OZ_BN) || (n % OZ_BM) || nkb * OZ_BK > 16384) return cudaErrorInvalidValue;
  OzParams p{};
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.S = reinterpret_cast<const int8_t*>(ws); p.RB = n / OZ_BM; p.KB = n / OZ_BK;
  p.SB = p.S; p.RBb = p.RB;
  p.scaleA = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(ws) + oz_slices_bytes(n, n, batch));
  p.scaleB = p.scaleA; p.sScale = n; p.sScaleB = n;
  p.rowA0 = rowA0; p.rowB0 = rowB0;
  p.R = R; p.ntc = ntc; p.ntiles = tri ? R * (R + 1) : R * ntc;
  p.ktiles = nkb; p.kb0 = kb0;
  p.tri = tri; p.lower_skip = lower_skip; p.row0 = row0; p.col0 = col0;
  p.info = info;
  p.total = p.ntiles * batch;
  return oz_launch<0>(p, st);
}

// Fused sampling tail on the int8 tensor cores: per (row tile of 128, column) max / first arg-max of mean + L Z.
// ws holds the slices of L (m rows, k = m, lower triangular: ozaki_ws_bytes(m, m, batch)) followed by those of
// Z^T (S rows, k = m: ozaki_ws_bytes(S, m, batch)).
size_t ozaki_sample_ws_bytes(int m, int S, int batch) { return ozaki_ws_bytes(m, m, batch) + ozaki_ws_bytes(S, m, batch); }

cudaError_t launch_ozaki_sample(const double* mean, int64_t sMean, const double* L, int m, int ldl, int64_t sL,
                                const double* Z, int S, int ldz, int64_t sZ, double* tile_max, int32_t* tile_arg,
                                const int32_t* info, int batch, void* ws, cudaStream_t st, int l_in_ws) {
  if ((m % OZ_BM) || (S % OZ_BN) || m > 16384) return cudaErrorInvalidValue;
  uint8_t* wsL = reinterpret_cast<uint8_t*>(ws);
  uint8_t* wsZ = wsL + ozaki_ws_bytes(m, m, batch);
  cudaError_t e = cudaSuccess;
  // l_in_ws: ocbo_potrf_ws left the slices of this very factor (and its row scales) at the head of the workspace
  if (!l_in_ws) e = launch_ozaki_split(L, ldl, sL, m, m, wsL, info, batch, st, 0, 1);
  if (e != cudaSuccess) return e;
  e = launch_ozaki_split(Z, ldz, sZ, S, m, wsZ, info, batch, st, 1, 0);
  if (e != cudaSuccess) return e;
  OzParams p{};
  p.S = reinterpret_cast<const int8_t*>(wsL); p.RB = m / OZ_BM; p.KB = m / OZ_BK;
  p.SB = reinterpret_cast<const int8_t*>(wsZ); p.RBb = (S + OZ_BM - 1) / OZ_BM;
  p.scaleA = reinterpret_cast<const double*>(wsL + oz_slices_bytes(m, m, batch)); p.sScale = m;
  p.scaleB = reinterpret_cast<const double*>(wsZ + oz_slices_bytes(S, m, batch)); p.sScaleB = S;
  p.R = m / OZ_BM; p.ntc = S / OZ_BN; p.ntiles = p.R * p.ntc;
  p.ktiles = m / OZ_BK;
  p.info = info;
  p.total = p.ntiles * batch;
  p.mean = mean; p.sMean = sMean; p.tile_max = tile_max; p.tile_arg = tile_arg; p.out_ld = S;
  return oz_launch<1>(p, st);
}

}  // namespace ocbo
