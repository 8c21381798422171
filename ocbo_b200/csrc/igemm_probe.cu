// tcgen05 probe: C[m x n] (int32) = A[m x k] (int8, k contiguous) * B[n x k]^T (int8, k contiguous)
// on the 5th-generation tensor cores -- tcgen05.mma kind::i8 issued by ONE thread per CTA, operands
// staged in shared memory by TMA (128-byte swizzle), the 128 x 128 int32 accumulator in TMEM, read
// back with tcgen05.ld.  FP64 has no tcgen05 kind; this kernel is the building block the Ozaki-split
// rank-k update would stand on (tests/tools/ozaki_probe.py) and is exported in
// the header's OCBO_BENCH section only.
//
// Roles (128 threads): thread 0 = TMA producer, thread 32 = MMA issuer, all four warps = epilogue
// (warp w owns TMEM lanes 32 w .. 32 w + 31 = rows of the tile).  Every wait is bounded (tma.cuh).
#include "common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace ocbo {

namespace {
constexpr int IBM = 128, IBN = 128;    // UMMA M x N
constexpr int IBK = 128;               // k per stage = one 128-byte swizzle row of int8
constexpr int IUK = 32;                // k per tcgen05.mma of 8-bit operands
constexpr int ISTAGES = 4;
constexpr int ITILE_A = IBM * IBK, ITILE_B = IBN * IBK;
constexpr int ISTAGE_BYTES = ITILE_A + ITILE_B;
constexpr int ISMEM = ISTAGES * ISTAGE_BYTES + 256 + 1024;  // + barriers + slack for the 1024-byte alignment
constexpr uint32_t TMEM_COLS = 128;

// instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D = S32, A = B = S8,
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(IBN >> 3) << 17) |
                           ((uint32_t)(IBM >> 4) << 24);

// K-major operand tile, 128-byte swizzle: 8-row groups 1024 bytes apart (SBO), LBO unused (1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(128, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int32_t* C,
             int64_t ldc, int ktiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ISTAGES * ISTAGE_BYTES);
  uint64_t* empty = full + ISTAGES;
  uint64_t* done = empty + ISTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ISTAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(done, 1);
    mbar_init_fence();
  }
  if (warp == 1) {  // one warp allocates the accumulator's TMEM columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int m0 = blockIdx.y * IBM, n0 = blockIdx.x * IBN;

  if (threadIdx.x == 0) {  // ---- TMA producer
    for (int kt = 0; kt < ktiles; ++kt) {
      const int s = kt % ISTAGES;
      if (kt >= ISTAGES) mbar_wait(empty + s, (uint32_t)(((kt / ISTAGES) - 1) & 1));
      mbar_expect_tx(full + s, ISTAGE_BYTES);
      tma_load_2d(smem + s * ISTAGE_BYTES, &mapA, full + s, kt * IBK, m0);
      tma_load_2d(smem + s * ISTAGE_BYTES + ITILE_A, &mapB, full + s, kt * IBK, n0);
    }
  } else if (threadIdx.x == 32) {  // ---- MMA issuer
    for (int kt = 0; kt < ktiles; ++kt) {
      const int s = kt % ISTAGES;
      mbar_wait(full + s, (uint32_t)((kt / ISTAGES) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(smem + s * ISTAGE_BYTES), b0 = a0 + ITILE_A;
#pragma unroll
      for (int j = 0; j < IBK / IUK; ++j)
        umma_i8(tmem, smem_desc(a0 + j * IUK), smem_desc(b0 + j * IUK), (kt | j) != 0);
      umma_commit(empty + s);  // the stage is free once these MMAs have read it
    }
    umma_commit(done);         // the accumulator is complete
  }
  // ---- epilogue: TMEM -> registers -> global
  __syncwarp();
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();
  int32_t* Crow = C + (int64_t)(m0 + warp * 32 + lane) * ldc + n0;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < IBN; c0 += 32) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr + (uint32_t)c0)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 32; q += 4)
      *reinterpret_cast<int4*>(Crow + c0 + q) = make_int4((int)v[q], (int)v[q + 1], (int)v[q + 2], (int)v[q + 3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

bool make_map_i8(CUtensorMap* map, const int8_t* base, uint64_t k, uint64_t rows, int64_t ld) {
  TmaEncodeFn enc = tma_encoder();
  if (!enc || ((uintptr_t)base & 15) || (ld & 15)) return false;
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)IBK, (cuuint32_t)IBM};
  cuuint32_t es[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

cudaError_t launch_probe_igemm(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C, int64_t ldc,
                               int m, int n, int k, cudaStream_t st) {
  CUtensorMap mA, mB;
  if (!make_map_i8(&mA, A, k, m, lda) || !make_map_i8(&mB, B, k, n, ldb)) return cudaErrorInvalidValue;
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ISMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  igemm_kernel<<<dim3(n / IBN, m / IBM), 128, ISMEM, st>>>(mA, mB, C, ldc, k / IBK);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
