// Dependent-issue latencies of the FP64 path on B200 (single warp, one SM): what bounds a serial elimination chain.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp3(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0); r = fma(r, e, r); e = fma(-d, r, 1.0); r = fma(r, e, r); return r; }
template <int MODE> __global__ void chain(double* out, int iters, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) x = fma(x, b, a);                         // DFMA chain
    if (MODE == 1) x = x * b;                                // DMUL chain
    if (MODE == 2) x = 1.0 / x + a;                          // IEEE division (+DADD)
    if (MODE == 3) x = __drcp_rn(x) + a;                     // drcp (+DADD)
    if (MODE == 4) x = fast_rcp3(x) + a;                     // MUFU + 2 Newton (+DADD)
    if (MODE == 5) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + a; }  // MUFU.RCP64H (+DADD)
    if (MODE == 6) x = sqrt(x) + a;
    if (MODE == 7) { __syncthreads(); x += 1.0; }            // barrier (+DADD)
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / iters; }
}
__global__ void lds_chain(double* out, int iters) {
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x + 1) & 255; __syncthreads();
  int p = threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = idx[p];
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = p; out[1] = (double)(t1 - t0) / iters; }
}
int main() {
  double* out; cudaMalloc(&out, 16); double h[2];
  const char* names[8] = {"DFMA dependent", "DMUL dependent", "1.0/x + DADD", "__drcp_rn + DADD", "MUFU+2NR + DADD", "MUFU.RCP64H + DADD", "sqrt + DADD", "__syncthreads(256 thr) + DADD"};
#define RUN(M, T) chain<M><<<1, T>>>(out, 2000, 1.3, 0.999); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("%-32s %7.1f cycles/iter (%d threads)\n", names[M], h[1], T);
  RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 256) RUN(0, 256) RUN(3, 256)
  lds_chain<<<1, 32>>>(out, 2000); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("%-32s %7.1f cycles/iter\n", "LDS dependent", h[1]);
  return 0;
}
