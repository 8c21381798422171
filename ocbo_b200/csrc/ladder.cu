// Device-side walk of Dragonfly's stable_cholesky jitter ladder
// for a whole batch, with no host round trip per rung.
//
// Protocol on info[b] while a ladder is being walked:
//     0   attempt pending / last attempt succeeded
//    >0   last attempt failed at that column (LAPACK dpotrf convention)
//    -1   parked: an earlier attempt succeeded, later rungs must not touch b
// One rung p is three steps on the stream:
//   ladder_restore  (problems with info > 0)  A <- A0 + 10^p max(diag A0) I
//   ladder_advance  info > 0 -> rung[b] = p, info = 0;   info == 0 -> -1
//   potrf           (runs where info == 0)
// and a final ladder_advance(finish) turns the parked -1 back into 0, leaving > 0
// only on problems that failed every rung tried so far.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

namespace {
constexpr int RT = 32;  // restore tile edge

__global__ void __launch_bounds__(256)
ladder_restore_kernel(double* A, const double* A0, int n, int lda, int64_t sA, int lda0, int64_t sA0,
                      const int32_t* nv, const int32_t* info, double factor) {
  const int b = blockIdx.z;
  if (info[b] <= 0) return;
  const int tj = blockIdx.x, ti = blockIdx.y;
  // everything on or below the diagonal of 128-blocks (what potrf reads and overwrites)
  if ((tj * RT) / TILE > (ti * RT) / TILE) return;
  const double* S = A0 + (size_t)b * sA0;
  double* Dm = A + (size_t)b * sA;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  double jit = 0.0;
  if (ti == tj) {
    // max of the valid diagonal of the ORIGINAL matrix (redundantly per diagonal tile)
    __shared__ double red[8];
    const int lim = nv ? nv[b] : n;
    double m = -CUDART_INF;
    for (int i = threadIdx.x; i < lim; i += 256) m = fmax(m, S[(size_t)i * lda0 + i]);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (tx == 0) red[ty] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    jit = factor * m;
  }
  const int lim = nv ? nv[b] : n;
#pragma unroll
  for (int k = 0; k < RT; k += 8) {
    const int r = ti * RT + ty + k, c = tj * RT + tx;
    double v = S[(size_t)r * lda0 + c];
    if (r == c && r < lim) v += jit;
    Dm[(size_t)r * lda + c] = v;
  }
}

__global__ void ladder_advance_kernel(int32_t* info, int32_t* rung, int p, int finish, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int s = info[b];
  if (finish) {
    if (s < 0) info[b] = 0;
    return;
  }
  if (s > 0) {
    rung[b] = p;
    info[b] = 0;
  } else if (s == 0) {
    info[b] = -1;
  }
}
// Speculative first rung, run CONCURRENTLY with the plain attempt on a copy (A1 = A0 + jitter):
// problems whose plain attempt failed (info > 0) while the speculative one succeeded
// (info1 == 0) adopt the speculative factor.  ladder_adopt copies, ladder_adopt_mark updates the
// status afterwards (two kernels: many CTAs of a problem read info while it must not change).
__global__ void __launch_bounds__(256)
ladder_adopt_kernel(double* A, const double* A1, int n, int lda, int64_t sA, double* invd,
                    const double* invd1, int64_t sInv, const int32_t* info, const int32_t* info1) {
  const int b = blockIdx.z;
  if (!(info[b] > 0 && info1[b] == 0)) return;
  const int tj = blockIdx.x, ti = blockIdx.y;
  if ((tj * RT) / TILE > (ti * RT) / TILE) return;
  const double* S = A1 + (size_t)b * sA;
  double* Dm = A + (size_t)b * sA;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < RT; k += 8) {
    const int r = ti * RT + ty + k, c = tj * RT + tx;
    Dm[(size_t)r * lda + c] = S[(size_t)r * lda + c];
  }
  if (ti == tj) {  // the inverted diagonal blocks: RT rows of the 128 x 128 block ti*RT/128
    const int blk = (ti * RT) / TILE, r0 = (ti * RT) % TILE;
    const double* Is = invd1 + (size_t)b * sInv + (size_t)blk * TILE * TILE + (size_t)r0 * TILE;
    double* Id = invd + (size_t)b * sInv + (size_t)blk * TILE * TILE + (size_t)r0 * TILE;
    for (int e = threadIdx.x; e < RT * TILE; e += 256) Id[e] = Is[e];
  }
}

__global__ void ladder_adopt_mark_kernel(int32_t* info, const int32_t* info1, int32_t* rung, int p,
                                         int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  if (info[b] > 0 && info1[b] == 0) {
    info[b] = 0;
    rung[b] = p;
  }
}
}  // namespace

cudaError_t launch_ladder_adopt(double* A, const double* A1, int n, int lda, int64_t sA, double* invd,
                                const double* invd1, int64_t sInv, int32_t* info,
                                const int32_t* info1, int32_t* rung, int p, int batch,
                                cudaStream_t st) {
  ladder_adopt_kernel<<<dim3(n / RT, n / RT, batch), 256, 0, st>>>(A, A1, n, lda, sA, invd, invd1,
                                                                    sInv, info, info1);
  cudaError_t e = counted(cudaGetLastError());
  if (e != cudaSuccess) return e;
  ladder_adopt_mark_kernel<<<(batch + 127) / 128, 128, 0, st>>>(info, info1, rung, p, batch);
  return counted(cudaGetLastError());
}

cudaError_t launch_ladder_restore(double* A, const double* A0, int n, int lda, int64_t sA, int lda0,
                                  int64_t sA0, const int32_t* nv, const int32_t* info,
                                  double factor, int batch, cudaStream_t st) {
  ladder_restore_kernel<<<dim3(n / RT, n / RT, batch), 256, 0, st>>>(A, A0, n, lda, sA, lda0, sA0,
                                                                      nv, info, factor);
  return counted(cudaGetLastError());
}

cudaError_t launch_ladder_advance(int32_t* info, int32_t* rung, int p, int finish, int batch,
                                  cudaStream_t st) {
  ladder_advance_kernel<<<(batch + 127) / 128, 128, 0, st>>>(info, rung, p, finish, batch);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
