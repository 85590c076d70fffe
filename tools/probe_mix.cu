// Do DMMA.8x8x4 and DFMA share an execution pipe on B200?  Half the warps of every CTA run a DMMA loop, the other
// half a DFMA loop; compare against each running alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) mix(double* out, int iters, int mode, double seed) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = (mode == 0) || (mode == 2 && warp < 4);
  const bool do_fma = (mode == 1) || (mode == 2 && warp >= 4);
  double s = 0;
  if (do_mma) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = seed * i; c[i][1] = seed; }
    double a = seed + threadIdx.x, b = seed - threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  }
  if (do_fma) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = seed * i;
    double a = 1.0000001 + threadIdx.x * 1e-9, b = seed * 1e-9;
    for (int it = 0; it < iters * 8; it++) {   // 8x more iterations: one DMMA = 8 DFMA worth of FMAs per thread
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 2 * 256 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  const char* names[3] = {"DMMA only (8 warps)", "DFMA only (8 warps)", "4 warps DMMA + 4 warps DFMA"};
  for (int mode = 0; mode < 3; mode++) {
    mix<<<148, 256>>>(out, iters, mode, 1e-3); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<<<148, 256>>>(out, iters, mode, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mma_fl = (mode == 0 ? 8 : mode == 2 ? 4 : 0) * 148.0 * iters * 16 * 512.0;
    double fma_fl = (mode == 1 ? 8 : mode == 2 ? 4 : 0) * 148.0 * 32 * (double)iters * 8 * 16 * 2.0;
    printf("%-32s %.3f ms  total %.2f TFLOP/s (DMMA part %.2f, DFMA part %.2f)\n", names[mode], ms, (mma_fl + fma_fl) / ms * 1e-9, mma_fl / ms * 1e-9, fma_fl / ms * 1e-9);
  }
  return 0;
}
