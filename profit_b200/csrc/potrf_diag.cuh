// Diagonal-block kernel of the blocked Cholesky: one CTA factorises a 128x128 FP64 block and inverts the
// triangular factor in the same sweep.  This is the only serial stage of the factorisation (reference:
// np.linalg.cholesky -> LAPACK dpotrf, gp_functions.py:143); everything else is tile GEMMs on the DMMA pipe
// (dgemm_tasks.cuh), including the panel solve, which multiplies by the inverse produced here.
//
// Register-resident right-looking elimination.  256 threads; thread (ty = tid/32, tx = tid%32) owns the elements
// (i = 8a + ty, k = 32b + tx), a = 0..15, b = 0..3, of the lower triangle of A and of R (R starts as I) --
// 40 + 40 doubles in registers.  Column step j costs one barrier:
//     owners of column j append c = A[j:, j] (unscaled: c_ij = L_ij sqrt(d_j), c_jj = d_j) to a packed
//     column store in shared memory, the warp owning row j appends r = R[j, :j+1] to a packed row store,
//     A[i, k] -= (c_i / d_j) c_k      for i >= k > j          (Cholesky trailing update)
//     R[i, m] -= (c_i / d_j) r_m      for i > j >= m          (forward substitution of the identity)
// and at the end  L[i, k] = c_ik / sqrt(d_k),  W[i, m] = r_im / sqrt(d_i)  with W = L^{-1}, read back from the
// packed stores.  Because a column/row is saved the moment it becomes final, the updates run unmasked over
// whole 8x32 register sub-blocks (finished entries just collect garbage), and the set of live sub-blocks only
// depends on j/8 -- the step body is instantiated 16 times (eight columns each), so it is straight-line DFMA
// code without per-element predicates or dispatch.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ptx.cuh"

namespace gpb {

constexpr int NB = 128;            // block size of the tile algorithms
constexpr int DIAG_RB = 8;                       // rows per register sub-block; the CTA has 32 * DIAG_RB threads
constexpr int DIAG_NA = NB / DIAG_RB;           // register row blocks per thread
constexpr int DIAG_THREADS = 32 * DIAG_RB;
constexpr int DIAG_TRI = NB * (NB + 1) / 2;                       // 8256 packed entries
constexpr int DIAG_SMEM = (2 * DIAG_TRI + 2 * NB) * (int)sizeof(double);

__host__ __device__ constexpr bool diag_owned(int a, int b) { return 32 * b <= DIAG_RB * a + DIAG_RB - 1; }
__device__ __forceinline__ int diag_colbase(int j) { return NB * j - (j * (j - 1)) / 2; }   // column j, rows j..127
__device__ __forceinline__ int diag_rowbase(int j) { return (j * (j + 1)) / 2; }            // row j, cols 0..j

// Eight consecutive column steps j = 8*JA .. 8*JA+7.  Everything that depends on j/8 (which register sub-blocks are
// live, which are masked) is a compile-time property of the instantiation, so the step body is straight-line code.
template <int JA, bool FACTOR>
__device__ __forceinline__ void diag_block8(double (&av)[DIAG_NA][4], double (&rv)[DIAG_NA][4], double* ccomp, double* rcomp,
                                            double* dvec, int ty, int tx, int tid, int k0, int n_valid, int* info) {
  constexpr int JB = JA * DIAG_RB / 32;
#pragma unroll 1
  for (int jj = 0; jj < DIAG_RB; ++jj) {
    const int j = DIAG_RB * JA + jj;
    double* col = ccomp + diag_colbase(j) - j;   // col[i] = c_ij for i >= j (reads below j stay inside the store)
    double* row = rcomp + diag_rowbase(j);       // row[m] = r_jm for m <= j
    // ---- publish the now-final column j of A and row j of R
    if (tx == (j & 31)) {
      if (ty >= jj) col[DIAG_RB * JA + ty] = av[JA][JB];
#pragma unroll
      for (int a = JA + 1; a < DIAG_NA; ++a) col[DIAG_RB * a + ty] = av[a][JB];
    }
    if (ty == jj) {
#pragma unroll
      for (int b = 0; b < JB; ++b) row[32 * b + tx] = rv[JA][b];
      if (32 * JB + tx <= j) row[32 * JB + tx] = rv[JA][JB];
    }
    __syncthreads();
    const double d = col[j];
    if (tid == 0) {
      dvec[j] = d;
      if (!(d > 0.0) && (k0 + j) < n_valid) atomicCAS(info, 0, k0 + j + 1);
    }
    const double invd = __drcp_rn(d);
    // ---- rank-1 updates, unmasked over the live sub-blocks
    double ck[4], rk[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      ck[b] = (FACTOR && b >= JB) ? col[32 * b + tx] : 0.0;
      rk[b] = (b < JB) ? row[32 * b + tx] : 0.0;
    }
    rk[JB] = (32 * JB + tx <= j) ? row[32 * JB + tx] : 0.0;
#pragma unroll
    for (int a0 = JA; a0 < DIAG_NA; a0 += 4) {
      double ci[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (a0 + u < DIAG_NA) ci[u] = col[DIAG_RB * (a0 + u) + ty] * invd;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (a0 + u < DIAG_NA) {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (diag_owned(a0 + u, b)) {
              if (FACTOR && b >= JB) av[a0 + u][b] = fma(-ci[u], ck[b], av[a0 + u][b]);
              if (b <= JB) rv[a0 + u][b] = fma(-ci[u], rk[b], rv[a0 + u][b]);
            }
        }
    }
  }
}

// In:  A (row-major, lda) holds the lower triangle of an SPD block at rows/cols [k0, k0+128).
// Out: A block   <- L (lower, strict upper zeroed)            [FACTOR]
//      Winv      <- L^{-1} (128x128 row-major, strict upper zeroed),
//      logdet[0] <- sum_j log L_jj over rows < n_valid (fixed summation order),
//      info      <- first non-positive pivot (global 1-based index), LAPACK style; untouched otherwise.
// With FACTOR == false the block already holds L and only Winv (and logdet) are produced.
// blockIdx.x selects consecutive diagonal blocks (grid > 1 only when they are independent: FACTOR == false).
template <bool FACTOR>
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, long long lda, int k0_first, double* __restrict__ Winv_first,
                  double* __restrict__ logdet_first, int* __restrict__ info, int n_valid) {
  const int k0 = k0_first + blockIdx.x * NB;
  double* Winv = Winv_first + (long long)blockIdx.x * NB * NB;
  double* logdet = logdet_first ? logdet_first + blockIdx.x : nullptr;
  extern __shared__ double diag_sm[];
  double* ccomp = diag_sm;               // packed final columns of A (unscaled)
  double* rcomp = diag_sm + DIAG_TRI;    // packed final rows of R (unscaled)
  double* dvec = rcomp + DIAG_TRI;       // pivots d_j
  double* rsv = dvec + NB;               // 1/sqrt(d_j)  (FACTOR)  or  1/L_jj
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  double* Ablk = A + (long long)k0 * lda + k0;
  pdl_launch_dependents();   // the panel solve that follows may start its prologue now
  pdl_wait();                // ... and this block reads what the previous update wrote

  double av[DIAG_NA][4], rv[DIAG_NA][4];
#pragma unroll
  for (int a = 0; a < DIAG_NA; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      av[a][b] = rv[a][b] = 0.0;
      if (diag_owned(a, b)) {
        const int i = DIAG_RB * a + ty, k = 32 * b + tx;
        if (k <= i) av[a][b] = Ablk[(long long)i * lda + k];
        if (k == i) rv[a][b] = 1.0;
      }
    }

#define GPB_DIAG_BLK(JA) \
  if constexpr (JA < DIAG_NA) diag_block8<JA, FACTOR>(av, rv, ccomp, rcomp, dvec, ty, tx, tid, k0, n_valid, info);
  GPB_DIAG_BLK(0) GPB_DIAG_BLK(1) GPB_DIAG_BLK(2) GPB_DIAG_BLK(3) GPB_DIAG_BLK(4) GPB_DIAG_BLK(5) GPB_DIAG_BLK(6)
  GPB_DIAG_BLK(7) GPB_DIAG_BLK(8) GPB_DIAG_BLK(9) GPB_DIAG_BLK(10) GPB_DIAG_BLK(11) GPB_DIAG_BLK(12)
  GPB_DIAG_BLK(13) GPB_DIAG_BLK(14) GPB_DIAG_BLK(15)
#undef GPB_DIAG_BLK
  __syncthreads();
  if (tid < NB) rsv[tid] = FACTOR ? 1.0 / sqrt(dvec[tid]) : 1.0 / dvec[tid];
  if (logdet != nullptr && tid >= 128 && tid < 160) {  // warp 4 (exists for every DIAG_RB >= 8)
    const int lane = tid - 128;
    double s = 0.0;
    for (int q = 0; q < 4; ++q) {
      const int j = lane + 32 * q;
      if (k0 + j < n_valid) s += FACTOR ? 0.5 * log(dvec[j]) : log(dvec[j]);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) logdet[0] = s;
  }
  __syncthreads();

  for (int a = 0; a < DIAG_NA; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = DIAG_RB * a + ty, k = 32 * b + tx;
      double l = 0.0, w = 0.0;
      if (k < i) l = ccomp[diag_colbase(k) + i - k] * rsv[k];
      else if (k == i) l = sqrt(dvec[i]);
      if (k <= i) w = rcomp[diag_rowbase(i) + k] * rsv[i];
      if (FACTOR) Ablk[(long long)i * lda + k] = l;
      Winv[i * NB + k] = w;
    }
}

}  // namespace gpb
