// TMA (cp.async.bulk.tensor) operand loads for the DMMA tile core: one elected thread per CTA
// moves whole operand tiles global -> shared with the tensor-memory accelerator and the
// consumers wait on an mbarrier transaction count, instead of 256 threads issuing 6 cp.async
// (LDGSTS) each per k-tile.  FP64 has no tcgen05 kind, so the consumer stays mma.sync DMMA; the
// data movement is Blackwell's.
//
// Bank-conflict-free fragment reads need padded shared-memory rows (leading dimension == 4 mod
// 16 doubles, gemm_core.cuh) and a TMA box is written densely, so the box itself carries the
// padding: its inner extent is the tile's plus 4 doubles (k-major tile: 20 x R instead of
// 16 x R), i.e. every row over-fetches 32 bytes of the NEXT k-tile (an L2 hit; zero-filled past
// the tensor's edge) that land exactly in the pad columns.  No swizzle, no change to the
// fragment addressing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocbo {

// ---- host: tensor maps ------------------------------------------------------
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

inline TmaEncodeFn tma_encoder() {
  static TmaEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TmaEncodeFn)p;
  }
  return fn;
}

// rank-3 fp64 tensor {inner, rows, batch}: row stride ld elements, batch stride sb elements
// (sb == 0: one shared matrix -> batch extent 1, the kernel then uses coordinate 0).
inline bool tma_make_map(CUtensorMap* map, const double* base, uint64_t inner, uint64_t rows,
                         int64_t ld, int64_t sb, int batch, uint32_t box_inner, uint32_t box_rows) {
  TmaEncodeFn enc = tma_encoder();
  if (!enc) return false;
  if (((uintptr_t)base & 15) || (ld & 1) || (sb & 1)) return false;
  cuuint64_t dims[3] = {inner, rows, (cuuint64_t)(sb ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(sb ? sb : (int64_t)ld * (int64_t)rows) * 8};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device: mbarrier + bulk tensor copy --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin until the phase with the given parity has completed.  Bounded: a transaction count that
// never arrives (a wrong box / byte count) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

}  // namespace ocbo
