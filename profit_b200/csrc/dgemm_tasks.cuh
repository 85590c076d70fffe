// Task-list driven FP64 tensor-core GEMM for sm_100a:  C_tile (op)= alpha * A_panel * B_panel^T.
//
// Every dense O(N^3) stage of the GP path (Cholesky panel solve and trailing update, the recursive
// triangular inverse, K^-1 = U U^T, the predictive-variance product L^-1 K*) is expressed on the host as
// a list of independent tile tasks; one CTA executes one task.  Both operands are row-major with the
// contraction index contiguous ("NT" form), so a 16-wide k-slab of either operand is a box of 128-byte
// rows that one TMA bulk-tensor copy drops into shared memory with the 128-byte swizzle.
//
//   producer warp : one elected lane walks the k-slabs through a STAGES-deep mbarrier ring (full/empty).
//   consumer warps: conflict-free LDS.128 fragment loads + DMMA.8x8x4 (mma.sync.m8n8k4.f64).
//
// Swizzle/fragment mapping.  In a slab, element (r, c) (c = 0..15 doubles) sits at byte
//   r*128 + (((c>>1) ^ (r&7)) << 4) + ((c&1) << 3).
// Lane (g = lane/4, t = lane%4) owns row g of an 8-row block and reads the two 16-byte chunks 2t and 2t+1,
// i.e. k = 4t..4t+3.  MMA number j (0..3) of the slab feeds element j of every lane, so its four k-slots are
// the actual k = {j, 4+j, 8+j, 12+j}; A and B use the same permutation, hence the contraction is exact.
// Within a quarter-warp (8 lanes: two g values x four t) the chunks {2t ^ g} are the eight distinct
// 16-byte columns of the 128-byte row, so every LDS.128 is bank-conflict free.
#pragma once
#include "ptx.cuh"

namespace gpb {

struct GemmTask {
  int a_row, a_k;   // origin of the A panel (row, first k column) in the matrix behind tensor map A
  int b_row, b_k;   // origin of the B panel in the matrix behind tensor map B
  int nk;           // number of 16-wide k-slabs
  int c_row, c_col; // origin of the output tile
  int aux;          // epilogue-specific (row index of the partial buffer for EPI_COLSQ)
};

enum { EPI_STORE = 0, EPI_COLSQ = 1 };

struct GemmEpilogue {
  double* C;        // EPI_STORE: output (row-major, ldc);  EPI_COLSQ: partial sums [aux][ldc]
  long long ldc;
  double* Ct;       // optional transposed copy: Ct[col*ldct + row] (may be nullptr)
  long long ldct;
  double alpha;
  double beta;      // C = alpha*acc + beta*C
};

template <int BM, int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + alignment slack
};

template <int BM, int BN, int WM, int WN, int STAGES, int EPI, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32 + 32, MINB)
dgemm_tasks_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmTask* __restrict__ tasks, const GemmEpilogue ep) {
  constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, NCW = WARPS_M * WARPS_N;
  constexpr int MI = WM / 8, NI = WN / 8;
  using SM = GemmSmem<BM, BN, STAGES>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const GemmTask task = tasks[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    mbar_init_fence();
  }
  __syncthreads();
  // Small (single-wave) launches sit on the Cholesky's panel chain: let the next kernel's CTAs come up while this
  // one runs.  Everything above is independent of the previous kernel; everything below may read its output.
  if (gridDim.x <= 296) pdl_launch_dependents();
  pdl_wait();
#ifdef GPB_TRACE
  const unsigned long long trace_t0 = trace_now();
#endif

  if (warp == NCW) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int kt = 0; kt < task.nk; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], SM::STAGE_BYTES);
        uint8_t* st = smem + s * SM::STAGE_BYTES;
        tma_load_2d(st, &tmA, task.a_k + kt * 16, task.a_row, &full[s]);
        tma_load_2d(st + SM::A_BYTES, &tmB, task.b_k + kt * 16, task.b_row, &full[s]);
      }
    }
    return;
  }

  // ---------------- DMMA consumers ----------------
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
  const uint32_t ch0 = static_cast<uint32_t>(((2 * t) ^ g) << 4);
  const uint32_t ch1 = static_cast<uint32_t>(((2 * t + 1) ^ g) << 4);
  const uint32_t a_off = static_cast<uint32_t>((wm0 + g) * 128);
  const uint32_t b_off = static_cast<uint32_t>(SM::A_BYTES + (wn0 + g) * 128);
  const uint32_t smem_base = smem_u32(smem);

  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  if constexpr (EPI == EPI_STORE) {
    // C = alpha*acc + beta*C reads the old tile in the epilogue: pull it into L2 now, behind the main loop.
    if (ep.beta != 0.0 && t < WN / 16) {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const double* line = ep.C + (long long)(task.c_row + wm0 + mi * 8 + g) * ep.ldc + task.c_col + wn0 + 16 * t;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
      }
    }
  }

  for (int kt = 0; kt < task.nk; ++kt) {
    const int s = kt % STAGES;
    const uint32_t ph = (kt / STAGES) & 1;
    mbar_wait(&full[s], ph);
    // Deferred release: the slab of the PREVIOUS step is handed back only here, behind the spin loop above.
    // Its LDS results were consumed by DMMAs that sit before the loop back-edge, so every shared-memory read of
    // that slab has returned before the producer may overwrite it (an arrive placed right after the loads can
    // overtake them in the memory pipeline when two CTAs share an SM).
    if (kt > 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[(kt - 1) % STAGES]);
    }
    const uint32_t st = smem_base + s * SM::STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t ch = h ? ch1 : ch0;
      double2 af[MI], bf[NI];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) af[mi] = lds128(st + a_off + mi * 1024 + ch);
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) bf[ni] = lds128(st + b_off + ni * 1024 + ch);
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi].x, bf[ni].x);
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi].y, bf[ni].y);
    }
  }

  // ---------------- epilogue ----------------
  if constexpr (EPI == EPI_STORE) {
    const double alpha = ep.alpha, beta = ep.beta;
    const bool has_beta = (beta != 0.0);
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const long long row = task.c_row + wm0 + mi * 8 + g;
      double* crow = ep.C + row * ep.ldc + task.c_col + wn0 + 2 * t;
      double2 old[NI];
      if (has_beta) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) old[ni] = *reinterpret_cast<const double2*>(crow + ni * 8);
      }
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        double2 o;
        o.x = alpha * acc[mi][ni][0];
        o.y = alpha * acc[mi][ni][1];
        if (has_beta) {
          o.x = fma(beta, old[ni].x, o.x);
          o.y = fma(beta, old[ni].y, o.y);
        }
        *reinterpret_cast<double2*>(crow + ni * 8) = o;
        if (ep.Ct != nullptr) {
          const long long col = task.c_col + wn0 + ni * 8 + 2 * t;
          ep.Ct[col * ep.ldct + row] = o.x;
          ep.Ct[(col + 1) * ep.ldct + row] = o.y;
        }
      }
    }
  } else {
    // Column sums of squares of the tile, reduced in a fixed order: rows of a lane, lanes (g) by shuffle,
    // warps along M through shared memory.  One partial row per task -> deterministic final reduction.
    double cs[NI][2];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        s0 = fma(acc[mi][ni][0], acc[mi][ni][0], s0);
        s1 = fma(acc[mi][ni][1], acc[mi][ni][1], s1);
      }
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
      }
      cs[ni][0] = s0;
      cs[ni][1] = s1;
    }
    // all k-slabs are consumed: stage 0 is free to reuse as the cross-warp scratch [WARPS_M][BN]
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    double* red = reinterpret_cast<double*>(smem);
    if (g == 0) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        red[(warp / WARPS_N) * BN + wn0 + ni * 8 + 2 * t] = cs[ni][0];
        red[(warp / WARPS_N) * BN + wn0 + ni * 8 + 2 * t + 1] = cs[ni][1];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    for (int c = threadIdx.x; c < BN; c += NCW * 32) {
      double s = 0.0;
#pragma unroll
      for (int wm = 0; wm < WARPS_M; ++wm) s += red[wm * BN + c];
      ep.C[static_cast<long long>(task.aux) * ep.ldc + task.c_col + c] = s;
    }
  }
#ifdef GPB_TRACE
  // tag: nk | (BN / 32) << 12 | BM << 16 | (beta != 0) << 28 | (Ct != nullptr) << 29 | EPI << 30; consumer warp 0 reports for the CTA
  if (threadIdx.x == 0)
    trace_emit(trace_t0, gridDim.x, (unsigned)task.nk | ((BN / 32) << 12) | (BM << 16) | ((ep.beta != 0.0) << 28) | ((ep.Ct != nullptr) << 29) | (EPI << 30));
#endif
}

}  // namespace gpb
