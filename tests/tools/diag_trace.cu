// Phase-level clock64() trace of the blocked diagonal-block kernel (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DOCBO_BLK_TRACE \
//      tests/tools/diag_trace.cu -o /tmp/diag_trace && /tmp/diag_trace
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ocbo_b200/csrc/potrf_diag_blk.cu"
namespace ocbo { unsigned long long g_launch_count = 0; }

int main() {
  const int n = 128;
  std::vector<double> G(n * (n + 8)), A(n * n);
  srand(1);
  for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = (i == j) ? 4.0 : 0.0;  // strongly diagonal: the factor is PD too (OCBO_BLK_REPS)
      for (int k = 0; k < n + 8; ++k) s += 0.01 * G[i * (n + 8) + k] * G[j * (n + 8) + k] / n;
      A[i * n + j] = s;
    }
  double *dA, *dInv; int* dInfo;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dInv, n * n * 8); cudaMalloc(&dInfo, 4);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaMemset(dInfo, 0, 4);
    ocbo::launch_potrf_diag_blk(dA, n, 0, 0, dInv, n * n, dInfo, 1, 0, nullptr);
    cudaDeviceSynchronize();
  }
  long long t[68];
  cudaMemcpyFromSymbol(t, ocbo::g_blk_trace, sizeof(t));
  printf("load %lld clk, store %lld clk, total %lld clk\n", t[0] - t[64], t[66] - t[65], t[66] - t[64]);
  for (int J = 0; J < 8; ++J)
    printf("J=%d: gather+bar %lld factor+solve %lld bar %lld stores %lld first-load %lld loop %lld bar %lld | step %lld\n",
           J, t[8 * J + 1] - t[8 * J], t[8 * J + 2] - t[8 * J + 1], t[8 * J + 5] - t[8 * J + 2],
           t[8 * J + 3] - t[8 * J + 5], J < 7 ? t[8 * J + 4] - t[8 * J + 3] : 0,
           J < 7 ? t[8 * J + 6] - t[8 * J + 4] : 0, t[8 * J + 7] - t[8 * J + 6], t[8 * J + 7] - t[8 * J]);
  int info; cudaMemcpy(&info, dInfo, 4, cudaMemcpyDeviceToHost);
  printf("info %d\n", info);
  return 0;
}
