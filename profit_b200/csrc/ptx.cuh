// Thin inline-PTX wrappers used by the sm_100a kernels: mbarrier, bulk-tensor TMA, FP64 DMMA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make freshly initialised mbarriers visible to the async (TMA) proxy.
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 2-D tiled TMA load: box lands at smem_dst, completion counted in bytes on `bar`.
// x = innermost (column / k) element coordinate, y = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

// Pull a 2-D tile into L2 only (used to warm the C tile before the epilogue reads it).
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(x), "r"(y)
               : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col), all FP64.  SASS: DMMA.8x8x4 (the only native FP64 MMA on sm_100a).
// Fragment ownership (g = lane/4, t = lane%4):  a = A[g][t],  b = B[t][g],  {c0,c1} = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Programmatic dependent launch (sm_90+): let the next kernel of the stream start its prologue early, and wait for
// the previous kernel's results before touching them.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef GPB_TRACE
// Debug build only (tools/trace_potrf.py): every CTA leaves {smid, start, end, grid, tag} in a global ring so that
// the occupancy timeline of a factorisation can be reconstructed.  Not compiled into the product library.
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ unsigned int g_trace_cap = 0;
__device__ unsigned int g_trace_count = 0;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_emit(unsigned long long t0, unsigned int grid, unsigned int tag) {
  if (g_trace_buf == nullptr) return;
  const unsigned int i = atomicAdd(&g_trace_count, 1u);
  if (i >= g_trace_cap) return;
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  g_trace_buf[4 * i + 0] = t0;
  g_trace_buf[4 * i + 1] = trace_now();
  g_trace_buf[4 * i + 2] = ((unsigned long long)grid << 32) | smid;
  g_trace_buf[4 * i + 3] = tag;
}
#endif

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

}  // namespace gpb
