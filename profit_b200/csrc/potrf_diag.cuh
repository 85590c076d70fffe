// Diagonal-block kernels of the blocked Cholesky: one CTA factorises a 128x128 FP64 block in shared
// memory, inverts the triangular factor, and emits both.  This is the only serial stage of the
// factorisation (reference: np.linalg.cholesky -> LAPACK dpotrf, gp_functions.py:143); everything
// else is tile GEMMs on the DMMA pipe (dgemm_tasks.cuh).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpb {

constexpr int NB = 128;            // block size of the tile algorithms
constexpr int DIAG_LDS = NB + 1;   // padded leading dimension in shared memory (column walks conflict-free)
constexpr int DIAG_THREADS = 1024;
constexpr int DIAG_SMEM = (NB * DIAG_LDS + 2 * NB) * (int)sizeof(double);

// In:  A (row-major, lda) holds the lower triangle of an SPD block at rows/cols [k0, k0+128).
// Out: A block   <- L (lower, strict upper zeroed),
//      Winv      <- L^{-1} (128x128 row-major, strict upper zeroed),
//      logdet[0] <- sum_j log L_jj  (fixed summation order),
//      info      <- first non-positive pivot (global 1-based index), LAPACK style; untouched otherwise.
// With factor == false the block already holds L and only Winv (and logdet) are produced.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, long long lda, int k0_first, double* __restrict__ Winv_first,
                  double* __restrict__ logdet_first, int* __restrict__ info, int n_valid, bool factor) {
  // blockIdx.x selects consecutive diagonal blocks (grid > 1 only when the blocks are independent, i.e. factor == false)
  const int k0 = k0_first + blockIdx.x * NB;
  double* Winv = Winv_first + (long long)blockIdx.x * NB * NB;
  double* logdet = logdet_first ? logdet_first + blockIdx.x : nullptr;
  extern __shared__ double sm[];
  double* S = sm;                       // [128][129]
  double* dvec = sm + NB * DIAG_LDS;    // pivots d_j
  double* rdiag = dvec + NB;            // 1 / L_jj
  const int tid = threadIdx.x;
  double* Ablk = A + (long long)k0 * lda + k0;

  for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
    const int i = e >> 7, j = e & 127;
    S[i * DIAG_LDS + j] = (j <= i) ? Ablk[(long long)i * lda + j] : 0.0;
  }

  if (factor) {
    // Right-looking, square-root free in the loop: column j stays unscaled (c_ij = L_ij * sqrt(d_j)) and the
    // rank-1 update uses c_i c_k / d_j, so one barrier per column suffices.
    const int ty = tid >> 5, tx = tid & 31;
    for (int j = 0; j < NB; ++j) {
      __syncthreads();
      const double d = S[j * DIAG_LDS + j];
      if (tid == 0) {
        dvec[j] = d;
        if (!(d > 0.0) && (k0 + j) < n_valid) atomicCAS(info, 0, k0 + j + 1);
      }
      const double invd = 1.0 / d;
      for (int i = j + 1 + ty; i < NB; i += 32) {
        const double ci = S[i * DIAG_LDS + j] * invd;
        for (int k = j + 1 + tx; k <= i; k += 32)
          S[i * DIAG_LDS + k] = fma(-ci, S[k * DIAG_LDS + j], S[i * DIAG_LDS + k]);
      }
    }
    __syncthreads();
    // scale columns: L_ij = c_ij / sqrt(d_j), L_jj = sqrt(d_j)
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
      const int i = e >> 7, j = e & 127;
      if (j < i) S[i * DIAG_LDS + j] = S[i * DIAG_LDS + j] / sqrt(dvec[j]);
    }
    __syncthreads();
    if (tid < NB) S[tid * DIAG_LDS + tid] = sqrt(dvec[tid]);
  }
  __syncthreads();
  if (tid < NB) rdiag[tid] = 1.0 / S[tid * DIAG_LDS + tid];
  __syncthreads();

  if (logdet != nullptr && tid < 32) {
    double s = 0.0;
    for (int q = 0; q < 4; ++q) {
      const int j = tid + 32 * q;
      s += (k0 + j < n_valid) ? log(S[j * DIAG_LDS + j]) : 0.0;
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (tid == 0) logdet[0] = s;
  }

  // W = L^{-1}, one warp per column j (4 columns per warp), forward substitution with the residual held in
  // registers (lane owns rows lane+32q).  Column j of W is parked in row j of the (free) strict upper
  // triangle of S; its diagonal entry is rdiag[j].
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < NB; j += 32) {
      double r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = (lane + 32 * q == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        for (int ii = 0; ii < 32; ++ii) {
          const int i = 32 * q + ii;
          if (i < j) continue;  // warp-uniform
          const double wi = __shfl_sync(0xffffffffu, r[q], ii) * rdiag[i];
          if (lane == ii) r[q] = wi;
#pragma unroll
          for (int q2 = q; q2 < 4; ++q2) {
            const int k = 32 * q2 + lane;
            if (k > i) r[q2] = fma(-S[k * DIAG_LDS + i], wi, r[q2]);
          }
        }
      }
      // r[q] now holds W[32q+lane][j] for rows >= j.  Nobody reads S[j][k>j] as L, so park it there.
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = 32 * q + lane;
        if (k > j) S[j * DIAG_LDS + k] = r[q];
      }
    }
  }
  __syncthreads();

  for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
    const int i = e >> 7, j = e & 127;
    if (factor) Ablk[(long long)i * lda + j] = (j <= i) ? S[i * DIAG_LDS + j] : 0.0;
    double w = 0.0;
    if (j == i) w = rdiag[i];
    else if (j < i) w = S[j * DIAG_LDS + i];   // W[i][j] parked at S[j][i]
    Winv[i * NB + j] = w;
  }
}

}  // namespace gpb
