// Device normals (Philox4x32-10 + Box-Muller), small-S sampling, the
// segmented max/argmax reduction and the joint-MTS pick; FP64 rate probes.
#include "gemm_core.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace ocbo {

// one thread per Philox call = two consecutive elements of the row-major m x S matrix
__global__ void philox_normal_kernel(double* Z, int m, int S, int ldz, int64_t sZ, uint64_t seed,
                                     const int32_t* trial_ids) {
  const int b = blockIdx.y;
  const uint64_t total = (uint64_t)m * (uint64_t)S;
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * q >= total) return;
  const uint32_t trial = trial_ids ? (uint32_t)trial_ids[b] : (uint32_t)b;
  double z0, z1;
  philox_normal_pair(seed, trial, q, z0, z1);
  double* Zb = Z + (size_t)b * sZ;
  const uint64_t e0 = 2 * q, e1 = 2 * q + 1;
  Zb[(size_t)(e0 / S) * ldz + (e0 % S)] = z0;
  if (e1 < total) Zb[(size_t)(e1 / S) * ldz + (e1 % S)] = z1;
}

cudaError_t launch_philox_normal(double* Z, int m, int S, int ldz, int64_t sZ, uint64_t seed,
                                 const int32_t* trial_ids, int batch, cudaStream_t st) {
  const uint64_t pairs = ((uint64_t)m * S + 1) / 2;
  const unsigned blocks = (unsigned)((pairs + 255) / 256);
  philox_normal_kernel<<<dim3(blocks, batch), 256, 0, st>>>(Z, m, S, ldz, sZ, seed, trial_ids);
  return counted(cudaGetLastError());
}

// F[i][s] = mean_i + sum_{k<=i} L[i][k] Z[k][s] for S <= 8: one warp per row,
// L read once, coalesced (HBM-bound: m^2/2 * 8 bytes).
template <int SMAX>
__global__ void __launch_bounds__(256)
sample_rowdot_kernel(const double* mean, int64_t sMean, const double* L, int m, int ldl, int64_t sL,
                     const double* Z, int S, int ldz, int64_t sZ, double* F, int ldf, int64_t sF,
                     const int32_t* info) {
  const int b = blockIdx.y;
  if (info && info[b] != 0) return;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const double* Lr = L + (size_t)b * sL + (size_t)i * ldl;
  const double* Zb = Z + (size_t)b * sZ;
  double acc[SMAX];
#pragma unroll
  for (int s = 0; s < SMAX; ++s) acc[s] = 0.0;
  for (int k = lane; k <= i; k += 32) {
    const double l = Lr[k];
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
      if (s < S) acc[s] = fma(l, Zb[(size_t)k * ldz + s], acc[s]);
  }
#pragma unroll
  for (int s = 0; s < SMAX; ++s) {
    if (s < S) {
      const double v = warp_sum(acc[s]);
      if (lane == 0) F[(size_t)b * sF + (size_t)i * ldf + s] = mean[(size_t)b * sMean + i] + v;
    }
  }
}

cudaError_t launch_sample_rowdot(const double* mean, int64_t sMean, const double* L, int m, int ldl,
                                 int64_t sL, const double* Z, int S, int ldz, int64_t sZ,
                                 double* F, int ldf, int64_t sF, const int32_t* info, int batch,
                                 cudaStream_t st) {
  dim3 grid((m + 7) / 8, batch);
  if (S == 1)
    sample_rowdot_kernel<1><<<grid, 256, 0, st>>>(mean, sMean, L, m, ldl, sL, Z, S, ldz, sZ, F, ldf, sF, info);
  else
    sample_rowdot_kernel<8><<<grid, 256, 0, st>>>(mean, sMean, L, m, ldl, sL, Z, S, ldz, sZ, F, ldf, sF, info);
  return counted(cudaGetLastError());
}

// (np.max / np.argmax semantics incl. NaN: arg_better in common.cuh)
// Per (problem, segment): max and FIRST argmax over rows [seg[j], seg[j+1]) of
// every sample column.  256 threads = CS columns x (256/CS) row lanes.
__global__ void __launch_bounds__(256)
segment_argmax_kernel(const double* F, int S, int ldf, int64_t sF, const int32_t* seg_ptr, int nseg,
                      double* out_max, int32_t* out_arg, int CS) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const int j = blockIdx.x, b = blockIdx.z;
  const int32_t* seg = seg_ptr + (size_t)b * (nseg + 1);
  const int r0 = seg[j], r1 = seg[j + 1];
  const double* Fb = F + (size_t)b * sF;
  const int tx = threadIdx.x % CS, ty = threadIdx.x / CS, RL = 256 / CS;
  // one CTA per (segment, chunk of CS <= 32 sample columns, problem): for the sweep (64
  // segments x 256 draws) that is 512 CTAs instead of 64, rows walked 4 at a time
  {
    const int s0 = blockIdx.y * CS;
    const int s = s0 + tx;
    double best = -CUDART_INF;
    int arg = -1;
    if (s < S) {
      int i = r0 + ty;
      for (; i + 3 * RL < r1; i += 4 * RL) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = Fb[(size_t)(i + u * RL) * ldf + s];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (arg < 0 || arg_better(v[u], i + u * RL, best, arg)) {
            best = v[u];
            arg = i + u * RL;
          }
      }
      for (; i < r1; i += RL) {
        const double v = Fb[(size_t)i * ldf + s];
        if (arg < 0 || arg_better(v, i, best, arg)) {
          best = v;
          arg = i;
        }
      }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = arg;
    __syncthreads();
    for (int h = RL >> 1; h > 0; h >>= 1) {
      if (ty < h) {
        const double ov = sv[threadIdx.x + h * CS];
        const int oi = si[threadIdx.x + h * CS];
        const double mv = sv[threadIdx.x];
        const int mi = si[threadIdx.x];
        if (oi >= 0 && (mi < 0 || arg_better(ov, oi, mv, mi))) {
          sv[threadIdx.x] = ov;
          si[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    if (ty == 0 && s < S) {
      const size_t o = ((size_t)b * nseg + j) * S + s;
      out_max[o] = sv[tx];
      out_arg[o] = si[tx] < 0 ? -1 : si[tx] - r0;
    }
    __syncthreads();
  }
}

cudaError_t launch_segment_argmax(const double* F, int S, int ldf, int64_t sF,
                                  const int32_t* seg_ptr, int nseg, double* out_max,
                                  int32_t* out_arg, int batch, cudaStream_t st) {
  int CS = 1;
  while (CS * 2 <= S && CS < 32) CS *= 2;
  segment_argmax_kernel<<<dim3(nseg, (S + CS - 1) / CS, batch), 256, 0, st>>>(
      F, S, ldf, sF, seg_ptr, nseg, out_max, out_arg, CS);
  return counted(cudaGetLastError());
}

// joint_mts.py:43-57 on the segment reductions; one thread per (problem, draw)
__global__ void joint_pick_kernel(const double* seg_max, const int32_t* seg_arg, int nseg,
                                  int n_old_seg, int T, int S, int32_t* out_task,
                                  int32_t* out_action, double* out_gain, int batch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * S) return;
  const int b = idx / S, s = idx % S;
  const double* mx = seg_max + (size_t)b * nseg * S;
  const int32_t* ag = seg_arg + (size_t)b * nseg * S;
  // the gains are independent loads (8 in flight); the action index is fetched once, for the
  // winner only (it used to sit inside the compare: a dependent load per task)
  double best = -CUDART_INF;
  int bt = -1;
  for (int t0 = 0; t0 < T; t0 += 8) {
    double gain[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + u;
      if (t < T) {
        const double old = n_old_seg ? mx[(size_t)t * S + s] : 0.0;
        gain[u] = mx[(size_t)(n_old_seg + t) * S + s] - old;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (t0 + u < T && (bt < 0 || arg_better(gain[u], t0 + u, best, bt))) {
        best = gain[u];
        bt = t0 + u;
      }
  }
  out_task[idx] = bt;
  out_action[idx] = ag[(size_t)(n_old_seg + bt) * S + s];
  out_gain[idx] = best;
}

cudaError_t launch_joint_pick(const double* seg_max, const int32_t* seg_arg, int nseg,
                              int n_old_seg, int T, int S, int32_t* out_task, int32_t* out_action,
                              double* out_gain, int batch, cudaStream_t st) {
  const int total = batch * S;
  joint_pick_kernel<<<(total + 127) / 128, 128, 0, st>>>(seg_max, seg_arg, nseg, n_old_seg, T, S,
                                                         out_task, out_action, out_gain, batch);
  return counted(cudaGetLastError());
}

// ------------------------------ rate probes --------------------------------
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* sink, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, bb = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, bb);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double* sink, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-9, bb = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, bb);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) sink[0] = s;
}

// kern_eval_se (8 exp chains in lockstep, common.cuh) next to kern_eval (libdevice's exp, one call per element)
__global__ void __launch_bounds__(256) probe_kern_se_kernel(const double* __restrict__ d2, double scale, int64_t n,
                                                            double* __restrict__ lockstep, double* __restrict__ plain) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = base + i < n ? d2[base + i] : 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (base + i < n) plain[base + i] = kern_eval(OCBO_KERNEL_SE, scale, v[i]);
  kern_eval_se<8>(scale, v);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (base + i < n) lockstep[base + i] = v[i];
}
cudaError_t launch_probe_kern_se(const double* d2, double scale, int64_t n, double* lockstep, double* plain,
                                 cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  probe_kern_se_kernel<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(d2, scale, n, lockstep, plain);
  return counted(cudaGetLastError());
}

cudaError_t launch_probe_dmma(double* sink, int blocks, int iters, cudaStream_t st) {
  probe_dmma_kernel<<<blocks, 256, 0, st>>>(sink, iters);
  return counted(cudaGetLastError());
}
cudaError_t launch_probe_dfma(double* sink, int blocks, int iters, cudaStream_t st) {
  probe_dfma_kernel<<<blocks, 256, 0, st>>>(sink, iters);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
