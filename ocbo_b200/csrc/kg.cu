// Knowledge gradient of a set of lines, batched: the inner function of REVI
// (/root/reference/src/util/misc_util.py:180-212 inside the loops of
// src/strategies/corr_opt.py:103-119 and src/cstrats/postmax_cts.py:147-160).
//
//   kg[c][g] = E[ max_i (a_gi + b_cgi Z) ] - max_i a_gi ,  Z ~ N(0, 1)
//   a_gi = judge_means[g][i],  b_cgi = inter[c][g E + i] / sqrt(noise + cand_var[c])
//
// One CTA per (group g, candidate c), E <= 1024 lines in shared memory:
//   1. bitonic sort of the lines by (slope b, intercept a) -- Python's sort of (b, a) tuples;
//   2. a -= max a;
//   3. upper envelope with a stack, ONE thread, the reference's loop verbatim (IEEE divisions,
//      pops while the new breakpoint lies left of the last one);
//   4. segment terms a (Phi(z_hi) - Phi(z_lo)) + b (phi(z_lo) - phi(z_hi)) in parallel, added up
//      by one thread in segment order (the reference's running sum; no FMA contraction).
// The reference spends ~1 ms of Python per (c, g) pair and 10^4 pairs per decision.
#include "common.cuh"
#include "kernels.h"

namespace ocbo {

namespace {
struct KgArgs {
  const double* means; int ldm;
  const double* inter; int64_t ldi;
  const double* cand_var; double noise;
  double* out;
  int G, E, NP;
};

__device__ __forceinline__ double kg_cdf(double z) {  // cephes ndtr's branch structure
  const double x = z * 0.70710678118654752440;
  const double ax = fabs(x);
  if (ax < 0.70710678118654752440) return __dadd_rn(0.5, __dmul_rn(0.5, erf(x)));
  const double y = __dmul_rn(0.5, erfc(ax));
  return x > 0.0 ? __dadd_rn(1.0, -y) : y;
}
__device__ __forceinline__ double kg_pdf(double z) {
  return exp(-__dmul_rn(z, z) / 2.0) / 2.50662827463100050242;
}
__device__ __forceinline__ bool kg_less(double b1, double a1, double b2, double a2) {
  return b1 < b2 || (b1 == b2 && a1 < a2);
}

__global__ void __launch_bounds__(256) kg_kernel(KgArgs p) {
  const int NT = blockDim.x;  // NP / 2 threads (32 .. 256): no idle warps at the sort's barriers
  extern __shared__ double sm[];
  const int NP = p.NP, E = p.E;
  double* b = sm;             // NP
  double* a = b + NP;         // NP
  double* zs = a + NP;        // NP + 1
  double* term = zs + NP + 1; // NP
  int* stack = reinterpret_cast<int*>(term + NP);  // NP
  __shared__ double red[8];
  __shared__ int nseg_s;
  const int g = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  const double denom = sqrt(p.noise + p.cand_var[c]);
  const double* mrow = p.means + (size_t)g * p.ldm;
  const double* irow = p.inter + (size_t)c * p.ldi + (size_t)g * E;
  for (int i = tid; i < NP; i += NT) {
    b[i] = (i < E) ? irow[i] / denom : CUDART_INF;
    a[i] = (i < E) ? mrow[i] : CUDART_INF;
  }
  __syncthreads();
  // 1. bitonic sort, ascending by (b, a)
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < NP; t += NT) {
        const int partner = t ^ j;
        if (partner > t) {
          const bool up = (t & k) == 0;
          const double b1 = b[t], a1 = a[t], b2 = b[partner], a2 = a[partner];
          const bool swap = up ? kg_less(b2, a2, b1, a1) : kg_less(b1, a1, b2, a2);
          if (swap) {
            b[t] = b2; a[t] = a2;
            b[partner] = b1; a[partner] = a1;
          }
        }
      }
      __syncthreads();
    }
  }
  // 2. a -= max(a)
  double m = -CUDART_INF;
  for (int i = tid; i < E; i += NT) m = fmax(m, a[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (NT >> 5); ++w) m = fmax(m, red[w]);
  for (int i = tid; i < E; i += NT) a[i] = a[i] - m;
  __syncthreads();
  // 3. upper envelope (one thread; the stack never underflows: zs[0] = -inf stops the pops)
  if (tid == 0) {
    int top = 1;
    stack[0] = 0;
    stack[1] = 1;
    zs[0] = -CUDART_INF;
    zs[1] = (a[0] - a[1]) / (b[1] - b[0]);
    for (int i = 2; i < E; ++i) {
      int j = stack[top];
      double z = (a[i] - a[j]) / (b[j] - b[i]);
      while (z < zs[top]) {
        --top;
        j = stack[top];
        z = (a[i] - a[j]) / (b[j] - b[i]);
      }
      ++top;
      stack[top] = i;
      zs[top] = z;
    }
    zs[top + 1] = CUDART_INF;
    nseg_s = top + 1;
  }
  __syncthreads();
  // 4. segment terms in parallel, running sum in order
  const int nseg = nseg_s;
  for (int s = tid; s < nseg; s += NT) {
    const int idx = stack[s];
    const double cdf_dif = __dadd_rn(kg_cdf(zs[s + 1]), -kg_cdf(zs[s]));
    const double pdf_dif = __dadd_rn(kg_pdf(zs[s]), -kg_pdf(zs[s + 1]));
    term[s] = __dadd_rn(__dmul_rn(a[idx], cdf_dif), __dmul_rn(b[idx], pdf_dif));
  }
  __syncthreads();
  if (tid == 0) {
    double r = 0.0;
    for (int s = 0; s < nseg; ++s) r = __dadd_rn(r, term[s]);
    p.out[(size_t)c * p.G + g] = r;
  }
}
}  // namespace

cudaError_t launch_kg(const double* means, int ldm, const double* inter, int64_t ldi,
                      const double* cand_var, double noise, int C, int G, int E, double* out,
                      cudaStream_t st) {
  int NP = 2;
  while (NP < E) NP <<= 1;
  KgArgs p{means, ldm, inter, ldi, cand_var, noise, out, G, E, NP};
  const int smem = (4 * NP + 1) * 8 + NP * 4;
  int threads = NP / 2;
  threads = threads < 32 ? 32 : (threads > 256 ? 256 : threads);
  kg_kernel<<<dim3(G, C), threads, smem, st>>>(p);
  return counted(cudaGetLastError());
}

}  // namespace ocbo
