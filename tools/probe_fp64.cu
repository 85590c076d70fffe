// FP64 pipe probe for B200: register-resident DMMA.8x8x4 and DFMA issue-rate ceilings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = seed * i; c[i][1] = seed; }
  double a = seed + threadIdx.x, b = seed - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = seed * i;
  double a = seed + threadIdx.x * 1e-9, b = seed * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  const int iters = 20000;
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = 148 * cps;  // cps CTAs of 8 warps per SM
    float ms = time_ms([&] { dmma_loop<16><<<grid, 256>>>(out, iters, 1e-3); }, 5);
    double flops = (double)grid * 8 * iters * 16 * 512.0;
    printf("DMMA.8x8x4 16acc  %d CTAs/SM x 8 warps: %.3f ms  %.2f TFLOP/s\n", cps, ms, flops / ms * 1e-9);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = 148 * cps;
    float ms = time_ms([&] { dmma_loop<32><<<grid, 256>>>(out, iters, 1e-3); }, 5);
    double flops = (double)grid * 8 * iters * 32 * 512.0;
    printf("DMMA.8x8x4 32acc  %d CTAs/SM x 8 warps: %.3f ms  %.2f TFLOP/s\n", cps, ms, flops / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { dmma_loop<4><<<148, 128>>>(out, iters, 1e-3); }, 5);
    double flops = 148.0 * 4 * iters * 4 * 512.0;
    printf("DMMA.8x8x4 4acc 4 warps/SM (latency probe): %.3f ms  %.2f TFLOP/s -> %.1f ns per dependent-4 round\n", ms, flops / ms * 1e-9, ms * 1e6 / iters);
  }
  {
    float ms = time_ms([&] { dmma_loop<1><<<148, 32>>>(out, iters, 1e-3); }, 5);
    printf("DMMA dependent chain latency: %.1f ns per mma\n", ms * 1e6 / iters);
  }
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = 148 * cps;
    float ms = time_ms([&] { dfma_loop<16><<<grid, 256>>>(out, iters, 1.0000001); }, 5);
    double flops = (double)grid * 256 * iters * 16 * 2.0;
    printf("DFMA 16acc  %d CTAs/SM x 8 warps: %.3f ms  %.2f TFLOP/s\n", cps, ms, flops / ms * 1e-9);
  }
  return 0;
}
