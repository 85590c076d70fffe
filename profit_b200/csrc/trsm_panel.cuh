// Panel solve of the blocked Cholesky as a blocked triangular solve, and the completion of the inverted diagonal
// blocks off the panel chain (round 2).  Reference: the panel rows of np.linalg.cholesky -> LAPACK dpotrf's dtrsm,
// gp_functions.py:143.
//
// The diagonal-block kernel on the chain (potrf_diagr_kernel<false>) only leaves L_kk and the four 32x32 diagonal
// blocks W_pp = L_pp^-1 of its inverse.  With them a 32-row stripe A of the panel below is solved block column by
// block column,
//     X_p = (A_p - sum_{q<p} X_q L_pq^T) W_pp^T ,   p = 0..3,
// which costs 10 block products of 32x32x32 on the FP64 tensor pipe instead of the 16 of a product with the full
// 128x128 inverse -- and keeps the 128x128 inverse off the critical path altogether.  trsm_panel_kernel executes the
// same tile-task lists as the GEMM formulation (one task = one 32-row stripe, CFG_32x128), so the scheduling plans and
// their host replays (tests/test_plans.py) are unchanged: the plan still says  L_ik = A_ik W_kk^T.
//
// complete_inverse_kernel fills the blocks of W_kk below the block diagonal,
//     W_ij = -W_ii sum_{m=j}^{i-1} L_im W_mj ,   i > j,
// one CTA per diagonal block, batched over the blocks a phase of the triangular inverse is about to read.
#pragma once
#include "dgemm_tasks.cuh"
#include "potrf_diagr.cuh"

namespace gpb {

constexpr int TP_S = 36;                                   // shared-memory row stride (doubles), as in potrf_diagr.cuh
constexpr int TP_BLK = 32 * TP_S;                          // one 32x32 block
constexpr int TP_THREADS = 256;
// 13 blocks = 117 KB: fits next to one resident tile-GEMM CTA (97 KB), so a panel solve does not have to wait for an
// SM to drain completely while a trailing update saturates the GPU
constexpr int TP_SMEM = (6 + 4 + 3) * TP_BLK * (int)sizeof(double);       // L_pq (6), W_pp (4), R_p / X_p (3, shared)
constexpr int CI_SMEM = 3 * TP_BLK * (int)sizeof(double);                 // T^T of the current block row (static)

__host__ __device__ constexpr int tp_lblk(int p, int q) { return p * (p - 1) / 2 + q; }   // p > q, 6 blocks

// global 32x32 block (row-major, ld) -> shared block (stride TP_S), 256 threads
__device__ __forceinline__ void tp_load_block(const double* __restrict__ src, long long ld, double* dst, int tid) {
#pragma unroll
  for (int e = tid; e < 32 * 16; e += TP_THREADS) {
    const int r = e >> 4, c = (e & 15) * 2;
    *reinterpret_cast<double2*>(dst + r * TP_S + c) = *reinterpret_cast<const double2*>(src + (long long)r * ld + c);
  }
}

// One CTA per task: the 32-row stripe at (task.a_row, task.a_k .. +128) of `A` is replaced by its solution against the
// diagonal block at (task.a_k, task.a_k) of the same matrix; Winv = the inverted diagonal block (128x128, ld 128;
// only its 32x32 diagonal blocks are read).
__global__ void __launch_bounds__(TP_THREADS, 1)
trsm_panel_kernel(double* __restrict__ A, long long lda, const double* __restrict__ Winv_all,
                  const GemmTask* __restrict__ tasks) {
  extern __shared__ double tp_sm[];
  double* Lk = tp_sm;                    // 6 blocks L_pq, p > q
  double* Wd = Lk + 6 * TP_BLK;          // 4 blocks W_pp
  double* Xb = Wd + 4 * TP_BLK;          // slot p < 3: right-hand side R_p, then the solved block column X_p
  const GemmTask task = tasks[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, tr = warp >> 1, tcb = 2 * (warp & 1);
  if (gridDim.x <= 592) pdl_launch_dependents();
  pdl_wait();
#ifdef GPB_TRACE
  const unsigned long long trace_t0 = trace_now();
#endif
  const int kb = task.a_k;                                             // first row / column of the diagonal block
  const double* Lkk = A + (long long)kb * lda + kb;
  const double* Wkk = Winv_all + (long long)(kb >> 7) * 128 * 128;
  double* stripe = A + (long long)task.a_row * lda + kb;

  // the stripe as DMMA accumulators: acc[p] = block column p (32x32), this warp's 8x16 strip of it
  double acc[4][2][2];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int tc = 0; tc < 2; ++tc) {
      const double2 v = *reinterpret_cast<const double2*>(stripe + (long long)(8 * tr + g) * lda + 32 * p + 8 * (tcb + tc) + 2 * t);
      acc[p][tc][0] = v.x;
      acc[p][tc][1] = v.y;
    }
#pragma unroll
  for (int p = 1; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < p; ++q) tp_load_block(Lkk + (long long)(32 * p) * lda + 32 * q, lda, Lk + tp_lblk(p, q) * TP_BLK, tid);
#pragma unroll
  for (int p = 0; p < 4; ++p) tp_load_block(Wkk + (32 * p) * 128 + 32 * p, 128, Wd + p * TP_BLK, tid);

#pragma unroll
  for (int p = 0; p < 4; ++p) {
    // right-hand side of block column p -> shared memory (the last one can use slot 0: X_0 is not needed any more)
    double* slot = Xb + (p < 3 ? p : 0) * TP_BLK;
#pragma unroll
    for (int tc = 0; tc < 2; ++tc)
      *reinterpret_cast<double2*>(slot + (8 * tr + g) * TP_S + 8 * (tcb + tc) + 2 * t) = make_double2(acc[p][tc][0], acc[p][tc][1]);
    __syncthreads();   // also covers the loads of L_pq / W_pp on the first pass
    double ar[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ar[ks] = slot[(8 * tr + g) * TP_S + 4 * ks + t];
    // X_p = R W_pp^T: W_pp is lower triangular, so output column n only sums k <= n
    double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const double* Wp = Wd + p * TP_BLK;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int tc = 0; tc < 2; ++tc)
        if (4 * ks <= 8 * (tcb + tc) + 7) dmma884(x[tc][0], x[tc][1], ar[ks], Wp[(8 * (tcb + tc) + g) * TP_S + 4 * ks + t]);
#pragma unroll
    for (int tc = 0; tc < 2; ++tc)
      *reinterpret_cast<double2*>(stripe + (long long)(8 * tr + g) * lda + 32 * p + 8 * (tcb + tc) + 2 * t) = make_double2(x[tc][0], x[tc][1]);
    if (p < 3) {
      __syncthreads();   // every warp has its R fragments: the slot may take X_p
#pragma unroll
      for (int tc = 0; tc < 2; ++tc)
        *reinterpret_cast<double2*>(slot + (8 * tr + g) * TP_S + 8 * (tcb + tc) + 2 * t) = make_double2(x[tc][0], x[tc][1]);
      __syncthreads();
      // R_p' -= X_p L_p'p^T for the block columns still to come
      double ax[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) ax[ks] = -slot[(8 * tr + g) * TP_S + 4 * ks + t];
#pragma unroll
      for (int pp = p + 1; pp < 4; ++pp) {
        const double* Lb = Lk + tp_lblk(pp, p) * TP_BLK;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
          for (int tc = 0; tc < 2; ++tc) dmma884(acc[pp][tc][0], acc[pp][tc][1], ax[ks], Lb[(8 * (tcb + tc) + g) * TP_S + 4 * ks + t]);
      }
    }
  }
#ifdef GPB_TRACE
  if (tid == 0) trace_emit(trace_t0, gridDim.x, 8u | (32u << 16));   // reported like a 32-row panel-solve GEMM tile
#endif
}

// One CTA per diagonal block kblk0 + blockIdx.x: L_kk (lower, in A) and the 32x32 diagonal blocks of Winv are given,
// the blocks of Winv below the block diagonal are produced (block rows 1..3, left to right dependencies only on
// earlier block rows).  The blocks above the block diagonal are not touched (zero-initialised by the library).
__global__ void __launch_bounds__(TP_THREADS)
complete_inverse_kernel(const double* __restrict__ A, long long lda, int kblk0, double* Winv_all) {
  // Operands come straight from global memory (the 64 KB of L_kk and W_kk stay in L1/L2; every fragment is used once
  // per warp), so the CTA only needs 27 KB of shared memory and fits next to the tile-GEMM CTAs of a saturated GPU --
  // the early phases of the triangular inverse run it on a low-priority stream during the Cholesky.
  __shared__ double Tt[3 * TP_BLK];      // transposed partial sums T_ij^T of the current block row
  const int kblk = kblk0 + blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, tr = warp >> 1, tcb = 2 * (warp & 1);
  const double* Lkk = A + (long long)kblk * 128 * lda + (long long)kblk * 128;
  double* Wkk = Winv_all + (long long)kblk * 128 * 128;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    // T_ij = sum_{m=j}^{i-1} L_im W_mj for j < i, kept as C fragments, then stored transposed
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double tacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int m = j; m < i; ++m) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const double a = Lkk[(long long)(32 * i + 8 * tr + g) * lda + 32 * m + 4 * ks + t];
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)   // B[k][n] = W_mj[k][n]
            dmma884(tacc[tc][0], tacc[tc][1], a, Wkk[(32 * m + 4 * ks + t) * 128 + 32 * j + 8 * (tcb + tc) + g]);
        }
      }
#pragma unroll
      for (int tc = 0; tc < 2; ++tc)
#pragma unroll
        for (int e = 0; e < 2; ++e) Tt[j * TP_BLK + (8 * (tcb + tc) + 2 * t + e) * TP_S + 8 * tr + g] = tacc[tc][e];
    }
    __syncthreads();
    // W_ij = -W_ii T_ij (W_ii lower triangular: k <= row tile)
    double aw[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) aw[ks] = -Wkk[(32 * i + 8 * tr + g) * 128 + 32 * i + 4 * ks + t];
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double w[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        if (ks < 2 * (tr + 1)) {
#pragma unroll
          for (int tc = 0; tc < 2; ++tc) dmma884(w[tc][0], w[tc][1], aw[ks], Tt[j * TP_BLK + (8 * (tcb + tc) + g) * TP_S + 4 * ks + t]);
        }
#pragma unroll
      for (int tc = 0; tc < 2; ++tc)
        *reinterpret_cast<double2*>(Wkk + (32 * i + 8 * tr + g) * 128 + 32 * j + 8 * (tcb + tc) + 2 * t) = make_double2(w[tc][0], w[tc][1]);
    }
    __syncthreads();   // block row i of W is visible to the whole CTA (global memory), T^T may be overwritten
  }
}

}  // namespace gpb
