// Stand-alone harness of the blocked diagonal-block kernel (profit_b200/csrc/potrf_diagr.cuh): correctness against a
// host Cholesky, device time, and per-phase clock64 stamps of the three role programs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DGPB_DIAG_CLOCKS -o tools/diagr_bench tools/diagr_bench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../profit_b200/csrc/potrf_diagr.cuh"

using namespace gpb;

#ifndef DR_BENCH_FULL
#define DR_BENCH_FULL false   // the panel-chain variant; -DDR_BENCH_FULL=true times the one with the full inverse
#endif

int main() {
  const int n = 128;
  std::vector<double> A(n * n), L(n * n, 0.0);
  srand(1);
  std::vector<double> M(n * n);
  for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += M[i * n + k] * M[j * n + k];
      A[i * n + j] = s + (i == j ? 8.0 : 0.0);
    }
  // host reference
  std::vector<double> R = A;
  for (int j = 0; j < n; ++j) {
    double d = R[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    L[j * n + j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = R[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / L[j * n + j];
    }
  }
  double *dA, *dB, *dW, *dl;
  int* dinfo;
  cudaMalloc(&dA, n * n * 8);
  cudaMalloc(&dB, n * n * 8);
  cudaMalloc(&dW, n * n * 8);
  cudaMalloc(&dl, 8);
  cudaMalloc(&dinfo, 4);
  cudaMemset(dinfo, 0, 4);
  cudaMemset(dW, 0, n * n * 8);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(potrf_diagr_kernel<DR_BENCH_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, DR_SMEM);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 20; ++r) {
    cudaMemcpy(dB, dA, n * n * 8, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    potrf_diagr_kernel<DR_BENCH_FULL><<<1, DR_THREADS, DR_SMEM>>>(dB, n, 0, dW, dl, dinfo, n);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(err));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<double> Lg(n * n), Wg(n * n);
  cudaMemcpy(Lg.data(), dB, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(Wg.data(), dW, n * n * 8, cudaMemcpyDeviceToHost);
  double el = 0, ew = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (j <= i) el = fmax(el, fabs(Lg[i * n + j] - L[i * n + j]));
      if (DR_BENCH_FULL || (i / 32 == j / 32)) {   // the chain variant only produces the 32x32 diagonal blocks of L^-1
        double s = 0;
        const int k0 = DR_BENCH_FULL ? 0 : (i / 32) * 32, k1 = DR_BENCH_FULL ? n : k0 + 32;
        for (int k = k0; k < k1; ++k) s += Wg[i * n + k] * L[k * n + j];
        ew = fmax(ew, fabs(s - (i == j ? 1.0 : 0.0)));
      }
    }
  printf("best %.2f us  L err %.2e  |W L - I| %.2e\n", best * 1e3, el, ew);
#ifdef GPB_DIAG_CLOCKS
  long long clk[3][64];
  cudaMemcpyFromSymbol(clk, g_dr_clk, sizeof(clk));
  long long stp[4][36];
  cudaMemcpyFromSymbol(stp, g_dr_step, sizeof(stp));
  const long long t0 = clk[0][0];
  auto rel = [&](int role, int slot) { return (double)(clk[role][slot] - t0); };
  printf("cycles since kernel start (last run); worker S1(0) passed at %.0f\n", rel(0, 1));
  for (int p = 0; p < 4; ++p) {
    printf(" P%d sweep: enter %lld loop %lld..%lld (%lld) exit %lld | helper enter %.0f done %.0f | worker: S1 %.0f inv(P-1) done %.0f S2 %.0f update done %.0f\n",
           p, stp[p][33] - t0, stp[p][0] - t0, stp[p][32] - t0, stp[p][32] - stp[p][0], stp[p][34] - t0, rel(2, 4 * p),
           rel(2, 4 * p + 1), rel(0, 8 * p + 1), p ? rel(0, 8 * p + 2) : 0.0, rel(0, 8 * p + 3), p < 3 ? rel(0, 8 * p + 4) : 0.0);
    printf("    cycles per column:");
    for (int j = 0; j < 32; ++j) printf(" %lld", stp[p][j + 1] - stp[p][j]);
    printf("\n");
  }
  printf(" tail: worker inverse<3> done %.0f  final barrier %.0f  output done %.0f\n", rel(0, 41), rel(0, 42), rel(0, 43));
#endif
  return 0;
}
