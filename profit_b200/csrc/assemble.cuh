// Covariance-kernel evaluation kernels (HBM-bound): K(X,X) assembly into the factorisation buffer, the
// rectangular K(Xa,Xb) [+ dK/dtheta planes] behind the `kernel(X, Y, ...)` plug-in API, the cross-kernel
// K(X*,X) with the predictive mean fused in, and the gradient contraction sum_ij A_ij dK_ij/dtheta_p
// that never materialises dK.
//
// Per-element arithmetic follows the reference: inputs are first divided by the length scale
// (python_kernels.py:35-36), differences are taken per dimension (:37-39), K = sf^2 exp(-d2/2) + sn^2 I
// (:41), dK/dl_d = (dx_d^2 / l_d) K_pure (:49-53), dK/dsf = 2/sf K_pure (:54), dK/dsn = 2 sn I (:55).
// Matern-5/2 has no reference implementation on this path; formulas in oracle/gp_oracle.py::matern52.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ptx.cuh"

namespace gpb {

enum { KIND_RBF = 0, KIND_MATERN52 = 1 };
constexpr int MAX_D = 32;

struct KernelHyp {
  int kind;
  int D;            // input dimension
  int n_ell;        // 1 (isotropic) or D (ARD)
  double ell[MAX_D];
  double sf, sn;
};

__device__ __forceinline__ double ell_of(const KernelHyp& h, int d) { return h.n_ell == 1 ? h.ell[0] : h.ell[d]; }

// k_pure and the radial derivative factor g with dK/dl_d = g * dx_d^2 / l_d  (dx scaled by 1/l).
template <int KIND>
__device__ __forceinline__ void kern_eval(double d2, double sf2, double& kp, double& gfac) {
  if (KIND == KIND_RBF) {
    kp = sf2 * exp(-0.5 * d2);
    gfac = kp;
  } else {
    const double s5 = 2.23606797749978969640917366873128;
    const double r = sqrt(d2);
    const double e = exp(-s5 * r);
    kp = sf2 * (1.0 + s5 * r + (5.0 / 3.0) * d2) * e;
    gfac = sf2 * (5.0 / 3.0) * (1.0 + s5 * r) * e;
  }
}

// Xs[i][d] = X[i][d] / l_d (row-major, DP columns, zero padded) and XsT[d][i] (DP x ldt).
__global__ void prescale_kernel(const double* __restrict__ X, int n, KernelHyp h, int DP, double* __restrict__ Xs,
                                double* __restrict__ XsT, long long ldt, int n_pad) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n_pad * DP) return;
  const int i = (int)(e / DP), d = (int)(e % DP);
  double v = 0.0;
  if (i < n && d < h.D) v = X[(long long)i * h.D + d] / ell_of(h, d);
  if (Xs) Xs[(long long)i * DP + d] = v;
  if (XsT) XsT[(long long)d * ldt + i] = v;
}

constexpr int AS_TM = 64, AS_TN = 128, AS_THREADS = 256;

// Radial-factor stash: when a gradient follows, the assembly leaves k_pure and the derivative factor g of every pair of the
// lower triangle behind, so that the gradient contraction needs neither exp nor sqrt again.  Layout: the lower 128x128
// blocks (I >= J) packed one after the other, block (I, J) at index I (I + 1) / 2 + J, row-major inside the block -- half
// the footprint of a square array, and a tile row of either kernel is one contiguous span.
__device__ __forceinline__ long long stash_offset(int r, int c) {
  const long long I = r >> 7, J = c >> 7;
  return ((I * (I + 1) / 2 + J) << 14) + ((r & 127) << 7) + (c & 127);
}

// Lower-triangular tiles of K(X,X) + sn^2 I into L (row-major, ldl).  Rows/cols >= n are padding: identity.
// G (and Kp for Matern-5/2, where g != k_pure) are the optional radial-factor stashes; nullptr = none.
template <int KIND, int DP>
__global__ void __launch_bounds__(AS_THREADS)
assemble_sym_kernel(const double* __restrict__ Xs, const double* __restrict__ XsT, long long ldt, int n, int n_pad,
                    double sf2, double sn2, double* __restrict__ L, long long ldl, double* __restrict__ G,
                    double* __restrict__ Kp) {
  const int row0 = blockIdx.y * AS_TM, col0 = blockIdx.x * AS_TN;
  if (col0 > row0 + AS_TM - 1) return;  // tile strictly above the diagonal
  __shared__ double xi[AS_TM][DP];
  for (int e = threadIdx.x; e < AS_TM * DP; e += AS_THREADS)
    xi[e / DP][e % DP] = Xs[(long long)(row0 + e / DP) * DP + e % DP];
  const int cp = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int c = col0 + 2 * cp;
  double xj0[DP], xj1[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const double2 v = *reinterpret_cast<const double2*>(XsT + (long long)d * ldt + c);
    xj0[d] = v.x;
    xj1[d] = v.y;
  }
  __syncthreads();
  for (int rr = rg; rr < AS_TM; rr += 4) {
    const int r = row0 + rr;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const double x = xi[rr][d];
      const double u0 = x - xj0[d], u1 = x - xj1[d];
      a0 = fma(u0, u0, a0);
      a1 = fma(u1, u1, a1);
    }
    double k0, k1, g0, g1;
    kern_eval<KIND>(a0, sf2, k0, g0);
    kern_eval<KIND>(a1, sf2, k1, g1);
    if (G != nullptr) {
      const long long so = stash_offset(r, c);
      *reinterpret_cast<double2*>(G + so) = make_double2(g0, g1);
      if (KIND != KIND_RBF) *reinterpret_cast<double2*>(Kp + so) = make_double2(k0, k1);
    }
    if (r == c) k0 += sn2;
    if (r == c + 1) k1 += sn2;
    if (r >= n || c >= n) k0 = (r == c) ? 1.0 : 0.0;
    if (r >= n || c + 1 >= n) k1 = (r == c + 1) ? 1.0 : 0.0;
    *reinterpret_cast<double2*>(L + (long long)r * ldl + c) = make_double2(k0, k1);
  }
}

// Rows [n_pad, n_pad+128) of the factorisation buffer carry y^T (one output per row) so that the panel solves
// of the Cholesky also perform the forward substitution z = L^{-1} y.
__global__ void fill_aug_kernel(const double* __restrict__ y, int n, int m, int n_pad, double* __restrict__ L,
                                long long ldl) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)128 * n_pad) return;
  const int r = (int)(e / n_pad), j = (int)(e % n_pad);
  double v = 0.0;
  if (r < m && j < n) v = y[(long long)j * m + r];
  L[(long long)(n_pad + r) * ldl + j] = v;
}

// Rectangular K(Xa, Xb) (+ sn^2 on the main diagonal, as np.eye(n, m) does) and optionally the
// (na, nb, P) derivative tensor, P ordered [l..., sf, sn] like python_kernels.py:42-56.  HBM-bound when the
// derivative planes are requested: 8 (1 + P) bytes written per element against ~100 FP64 operations.
//
// CTA tile: KM_ROWS rows x 64 columns, 256 threads = 64 columns x 4 row lanes.  A thread keeps its column's
// scaled coordinates in registers; the row coordinates are broadcast from shared memory.  Per pass (4 rows) the
// P derivative values of every element are written to a shared-memory stage laid out exactly like the global
// tensor -- a row of the tile is one contiguous span of 64 P doubles -- and shipped with one bulk TMA store per
// row (cp.async.bulk.global.shared::cta, SASS UBLKCP); two stages alternate so the stores of pass p overlap the
// arithmetic of pass p+1.  Ragged or 16-byte-misaligned rows fall back to a cooperative coalesced copy.
constexpr int KM_ROWS = 32, KM_COLS = 64, KM_THREADS = 256;

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

template <int KIND, int DP>
__global__ void __launch_bounds__(KM_THREADS, (DP <= 12) ? 3 : 2)
kernel_matrix_kernel(const double* __restrict__ XsA, int na, const double* __restrict__ XsTB, long long ldt, int nb,
                     KernelHyp h, double* __restrict__ K, double* __restrict__ dK, int nbuf) {
  extern __shared__ __align__(16) double km_sm[];
  const int P = h.n_ell + 2;
  double* xi = km_sm;                                  // [KM_ROWS][DP]
  double* rl = xi + KM_ROWS * DP;                      // [DP]  1 / l_d  (0 for padding dimensions)
  double* stage = rl + DP + (DP & 1);                  // [nbuf][4][64 * P], 16-byte aligned
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int row0 = blockIdx.y * KM_ROWS, col0 = blockIdx.x * KM_COLS;
  const int j = col0 + tx;
  const double sf2 = h.sf * h.sf, sn2 = h.sn * h.sn;
  for (int e = threadIdx.x; e < KM_ROWS * DP; e += KM_THREADS) {
    const int r = row0 + e / DP;
    xi[e] = (r < na) ? XsA[(long long)r * DP + e % DP] : 0.0;
  }
  if (threadIdx.x < DP) rl[threadIdx.x] = (threadIdx.x < h.D) ? 1.0 / ell_of(h, threadIdx.x) : 0.0;
  double xj[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) xj[d] = (j < nb) ? XsTB[(long long)d * ldt + j] : 0.0;
  const int ncols = min(KM_COLS, nb - col0);
  // bulk stores need 16-byte aligned global addresses: every row start is (i*nb + col0)*P*8 bytes into dK
  const bool bulk = (dK != nullptr) && ncols == KM_COLS && ((((long long)nb * P) & 1) == 0) &&
                    ((reinterpret_cast<unsigned long long>(dK) & 15ull) == 0);
  const bool ard_even = (h.n_ell > 1) && ((P & 1) == 0);   // 16-byte staged stores are aligned
  const double two_over_sf = 2.0 / h.sf;
  __syncthreads();

  for (int pass = 0; pass < KM_ROWS / 4; ++pass) {
    const int rr = pass * 4 + ty, i = row0 + rr;
    const bool valid = (i < na && j < nb);
    double u2[DP];
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const double u = xi[rr * DP + d] - xj[d];
      u2[d] = u * u;
      d2 += u2[d];
    }
    double kp, gf;
    kern_eval<KIND>(d2, sf2, kp, gf);
    if (valid) K[(long long)i * nb + j] = kp + ((i == j) ? sn2 : 0.0);
    if (dK == nullptr) continue;

    double* st = stage + (size_t)(pass % nbuf) * 4 * KM_COLS * P;
    if (nbuf == 2 && pass >= 2) {  // two stages: wait for the stores issued two passes ago before overwriting
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
    }
    double* my = st + ((size_t)ty * KM_COLS + tx) * P;
    const double dsf = two_over_sf * kp, dsn = (i == j) ? 2.0 * h.sn : 0.0;
    if (h.n_ell == 1) {
      my[0] = d2 * rl[0] * gf;
      my[1] = dsf;
      my[2] = dsn;
    } else if (ard_even) {
      // D is even here: pairs (d, d+1) are 16-byte aligned in the stage
#pragma unroll
      for (int d = 0; d < DP; d += 2)
        if (d < h.D) *reinterpret_cast<double2*>(my + d) = make_double2(u2[d] * rl[d] * gf, u2[d + 1] * rl[d + 1] * gf);
      *reinterpret_cast<double2*>(my + P - 2) = make_double2(dsf, dsn);
    } else {
#pragma unroll
      for (int d = 0; d < DP; ++d)
        if (d < h.D) my[d] = u2[d] * rl[d] * gf;
      my[P - 2] = dsf;
      my[P - 1] = dsn;
    }
    if (bulk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int r = 0; r < 4; ++r) {
          const int ii = row0 + pass * 4 + r;
          if (ii < na)
            bulk_store(dK + ((long long)ii * nb + col0) * P, st + (size_t)r * KM_COLS * P,
                       (unsigned)(KM_COLS * P * sizeof(double)));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // three stages: the stage written in the NEXT pass was shipped two passes ago; waiting here, before this
        // thread reaches the next barrier, makes one barrier per pass sufficient
        if (nbuf == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    } else {
      __syncthreads();
      for (int r = 0; r < 4; ++r) {
        const int ii = row0 + pass * 4 + r;
        if (ii >= na) break;
        double* dst = dK + ((long long)ii * nb + col0) * P;
        const double* src = st + (size_t)r * KM_COLS * P;
        for (int e = threadIdx.x; e < ncols * P; e += KM_THREADS) dst[e] = src[e];
      }
      __syncthreads();
    }
  }
  if (bulk && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Gradient contraction over the lower triangle (off-diagonal entries weighted twice):
//   acc_p = sum_ij w_ij * dKfac_ij,p   with  w = n_out * Kinv_ij - sum_r alpha_r[i] alpha_r[j]
// (gp_functions.py:169-180 with the trace written as an elementwise contraction).  Partial sums per CTA go to
// part[cta][DP+2] (slots: l_0..l_{DP-1}, sf, sn) and are reduced in a fixed order by reduce_partials_kernel.
template <int KIND, int DP>
__global__ void __launch_bounds__(AS_THREADS, (DP <= 16) ? 2 : 1)
grad_contract_kernel(const double* __restrict__ Xs, const double* __restrict__ XsT, long long ldt, int n, double sf2,
                     const double* __restrict__ Kinv, long long ldk, const double* __restrict__ alpha,
                     long long lda, int n_out, double* __restrict__ part) {
  const int row0 = blockIdx.y * AS_TM, col0 = blockIdx.x * AS_TN;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  double acc[DP + 2];
#pragma unroll
  for (int p = 0; p < DP + 2; ++p) acc[p] = 0.0;
  __shared__ double xi[AS_TM][DP];
  __shared__ double ar[AS_TM];                       // alpha of the tile rows (single-output fast path)
  __shared__ double red[AS_THREADS / 32][DP + 2];
  const bool active = !(col0 > row0 + AS_TM - 1) && row0 < n && col0 < n;
  if (active) {
    for (int e = threadIdx.x; e < AS_TM * DP; e += AS_THREADS)
      xi[e / DP][e % DP] = Xs[(long long)(row0 + e / DP) * DP + e % DP];
    if (threadIdx.x < AS_TM) ar[threadIdx.x] = alpha[row0 + threadIdx.x];
    // thread owns ONE column (128 column lanes) and every second row of the tile: half the registers of a
    // two-column layout, two CTAs per SM
    const int cl = threadIdx.x & 127, rg = threadIdx.x >> 7;
    const int c = col0 + cl;
    double xj[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) xj[d] = XsT[(long long)d * ldt + c];
    const double ac = alpha[c];
    __syncthreads();
    if (c < n) {
      // Rows in batches of GB: the K^-1 entries of a batch are all requested before the first one is used, so the
      // FP64 work of one row covers the memory latency of the next ones.
      constexpr int GB = 8;
      for (int rb = rg; rb < AS_TM; rb += 2 * GB) {
        double kin[GB];
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          const int r = row0 + rb + 2 * q;
          kin[q] = (r < n && c <= r) ? Kinv[(long long)r * ldk + c] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          const int rr = rb + 2 * q, r = row0 + rr;
          if (r >= n || c > r) continue;
          double aa;
          if (n_out == 1) {
            aa = ar[rr] * ac;
          } else {
            aa = 0.0;
            for (int o = 0; o < n_out; ++o) aa = fma(alpha[(long long)o * lda + r], alpha[(long long)o * lda + c], aa);
          }
          double w = n_out * kin[q] - aa;
          w = (c == r) ? w : 2.0 * w;
          double u2[DP];
          double d2 = 0.0;
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            const double v = xi[rr][d] - xj[d];
            u2[d] = v * v;
            d2 += u2[d];
          }
          double k0, g0;
          kern_eval<KIND>(d2, sf2, k0, g0);
          const double wg = w * g0;
#pragma unroll
          for (int d = 0; d < DP; ++d) acc[d] = fma(wg, u2[d], acc[d]);
          acc[DP] = fma(w, k0, acc[DP]);
          if (c == r) acc[DP + 1] += w;
        }
      }
    }
  }
  // fixed-order block reduction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int p = 0; p < DP + 2; ++p) {
    double v = acc[p];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][p] = v;
  }
  __syncthreads();
  if (threadIdx.x < DP + 2) {
    double s = 0.0;
    for (int w = 0; w < AS_THREADS / 32; ++w) s += red[w][threadIdx.x];
    part[(long long)cta * (DP + 2) + threadIdx.x] = s;
  }
}

// The same contraction from the radial-factor stash of the assembly (see stash_offset): per pair two or three 8-byte
// operands and ~3 FP64 operations per input dimension -- no exp, no sqrt, no distance sum -- which makes the kernel
// HBM-bound (8 bytes of K^-1 + 8 (RBF) or 16 (Matern-5/2) bytes of stash per pair of the lower triangle).
// One CTA owns a TM x 128 tile.  Its operand rows -- 1 KB contiguous each in K^-1 and in the stash blocks -- are streamed
// through a GC_STAGES-deep shared-memory ring by bulk TMA copies (cp.async.bulk + mbarrier, issued by one warp, one copy per lane): the bytes
// in flight sit in shared memory instead of registers, so two CTAs per SM keep ~100 KB outstanding.  The loop body is
// branch-free (invalid pairs get weight zero).
template <int DP>
struct GcTile {
  static constexpr int TM = (DP <= 16) ? 256 : 128;   // tall tile of large problems (shared memory: TM x DP doubles)
  static constexpr int TM_SMALL = 64;                 // below ~8k rows: more, smaller CTAs fill the GPU
};
constexpr int GC_ROWS = 8, GC_STAGES = 3;             // rows per stage, stages in the ring

template <int KIND, int DP, int TM>
struct GcSmem {
  static constexpr int NA = (KIND == KIND_RBF) ? 2 : 3;               // K^-1, g (, k_pure)
  static constexpr int STAGE_DOUBLES = NA * GC_ROWS * AS_TN;
  static constexpr int DOUBLES = GC_STAGES * STAGE_DOUBLES + TM * DP + TM + (AS_THREADS / 32) * (DP + 2);
  static constexpr int BYTES = DOUBLES * 8 + GC_STAGES * 8 + 16;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int KIND, int DP, int TM>
__global__ void __launch_bounds__(AS_THREADS, (DP <= 16) ? 2 : 1)
grad_contract_stash_kernel(const double* __restrict__ Xs, const double* __restrict__ XsT, long long ldt, int n, int n_pad,
                           const double* __restrict__ Kinv, long long ldk, const double* __restrict__ alpha, long long lda,
                           int n_out, const double* __restrict__ G, const double* __restrict__ Kp,
                           double* __restrict__ part) {
  using SM = GcSmem<KIND, DP, TM>;
  extern __shared__ __align__(16) double gc_sm[];
  double* stage = gc_sm;                                         // [GC_STAGES][NA][GC_ROWS][128]
  double* xi = stage + GC_STAGES * SM::STAGE_DOUBLES;            // [TM][DP]
  double* ar = xi + TM * DP;                                     // [TM]
  double* red = ar + TM;                                         // [warps][DP + 2]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + (AS_THREADS / 32) * (DP + 2));

  const int row0 = blockIdx.y * TM, col0 = blockIdx.x * AS_TN;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  double acc[DP + 2];
#pragma unroll
  for (int p = 0; p < DP + 2; ++p) acc[p] = 0.0;
  const bool active = !(col0 > row0 + TM - 1) && row0 < n && col0 < n;
  if (active) {
    // rows above the diagonal block of this column tile carry no pairs: start at the block's first row
    const int first = (col0 > row0) ? (col0 - row0) : 0;
    const int step0 = first / GC_ROWS, nsteps = TM / GC_ROWS;
    // Warp 0 issues the copies of a step into ring slot (step - step0) % GC_STAGES, one copy per lane so that they leave
    // together: lanes 0..7 the GC_ROWS rows of K^-1 (1 KB each, row stride ldk), lane 8 the GC_ROWS x 128 stash rows of g
    // and lane 9 those of k_pure -- 8 KB contiguous each, because GC_ROWS rows of one 128-block follow one another in
    // the packed layout.
    auto issue = [&](int step) {
      const int slot = (step - step0) % GC_STAGES, lane = threadIdx.x;
      double* dst = stage + slot * SM::STAGE_DOUBLES;
      const int r = row0 + step * GC_ROWS;
      if (lane == 0) mbar_arrive_expect_tx(&full[slot], SM::NA * GC_ROWS * AS_TN * 8);
      __syncwarp();
      if (lane < GC_ROWS)
        bulk_load(dst + lane * AS_TN, Kinv + (long long)(r + lane) * ldk + col0, AS_TN * 8, &full[slot]);
      else if (lane == GC_ROWS)
        bulk_load(dst + GC_ROWS * AS_TN, G + stash_offset(r, col0), GC_ROWS * AS_TN * 8, &full[slot]);
      else if (lane == GC_ROWS + 1 && KIND != KIND_RBF)
        bulk_load(dst + 2 * GC_ROWS * AS_TN, Kp + stash_offset(r, col0), GC_ROWS * AS_TN * 8, &full[slot]);
    };
    if (threadIdx.x < 32) {
      if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GC_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
      }
      __syncwarp();
      for (int s = 0; s < GC_STAGES - 1 && step0 + s < nsteps; ++s) issue(step0 + s);
    }
    // n_pad is a multiple of 128, not of TM: the last tile row may reach past the padded arrays
    for (int e = threadIdx.x; e < TM * DP; e += AS_THREADS)
      xi[e] = (row0 + e / DP < n_pad) ? Xs[(long long)row0 * DP + e] : 0.0;
    for (int e = threadIdx.x; e < TM; e += AS_THREADS) ar[e] = (row0 + e < n_pad) ? alpha[row0 + e] : 0.0;
    const int cl = threadIdx.x & 127, rg = threadIdx.x >> 7;
    const int c = col0 + cl;
    double xj[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) xj[d] = XsT[(long long)d * ldt + c];
    const double ac = alpha[c];
    __syncthreads();   // xi / ar filled, barriers initialised
    for (int step = step0; step < nsteps; ++step) {
      const int k = step - step0, slot = k % GC_STAGES;
      // the slot refilled now was read in the previous step; every thread has passed that step's closing barrier
      // (per-warp release barriers instead of the CTA barrier measured slower: 1.28 vs 1.10 ms at N = 16384)
      if (threadIdx.x < 32 && step + GC_STAGES - 1 < nsteps) issue(step + GC_STAGES - 1);
      mbar_wait(&full[slot], (k / GC_STAGES) & 1);
      const double* sk = stage + slot * SM::STAGE_DOUBLES;
#pragma unroll
      for (int q = 0; q < GC_ROWS / 2; ++q) {
        const int lr = rg + 2 * q, rr = step * GC_ROWS + lr, r = row0 + rr;
        const bool ok = (r < n && c <= r && c < n);
        const double kin = sk[lr * AS_TN + cl];
        const double gq = sk[(GC_ROWS + lr) * AS_TN + cl];
        const double kq = (KIND == KIND_RBF) ? gq : sk[(2 * GC_ROWS + lr) * AS_TN + cl];
        double aa;
        if (n_out == 1) {
          aa = ar[rr] * ac;
        } else {
          aa = 0.0;
          if (ok)
            for (int o = 0; o < n_out; ++o) aa = fma(alpha[(long long)o * lda + r], alpha[(long long)o * lda + c], aa);
        }
        double w = ok ? (n_out * kin - aa) : 0.0;
        const double wd = (c == r) ? w : 0.0;
        w = (c == r) ? w : 2.0 * w;
        const double wg = ok ? w * gq : 0.0;
        const double wk = ok ? w * kq : 0.0;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const double v = xi[rr * DP + d] - xj[d];
          acc[d] = fma(wg, v * v, acc[d]);
        }
        acc[DP] += wk;
        acc[DP + 1] += wd;
      }
      __syncthreads();
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int p = 0; p < DP + 2; ++p) {
    double v = acc[p];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp * (DP + 2) + p] = v;
  }
  __syncthreads();
  if (threadIdx.x < DP + 2) {
    double s = 0.0;
    for (int w = 0; w < AS_THREADS / 32; ++w) s += red[w * (DP + 2) + threadIdx.x];
    part[(long long)cta * (DP + 2) + threadIdx.x] = s;
  }
}

// out[p] = sum_c part[c][p] in a fixed (tree) order: deterministic across runs and launches.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ part, int ncta, int width,
                                                             double* __restrict__ out) {
  __shared__ double sh[256];
  const int p = blockIdx.x;
  double s = 0.0;
  for (int c = threadIdx.x; c < ncta; c += 256) s += part[(long long)c * width + p];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[p] = sh[0];
}

// Cross kernel for prediction: KsT[c][i] = k(x*_c, x_i) (no noise; zero in padded columns) for a chunk of
// candidates, with the predictive mean fused:  mean[c][o] = sum_i KsT[c][i] * alpha_o[i]
// (custom_surrogate.py:109,113).  One CTA owns CR candidate rows and sweeps all training points.
constexpr int CR_ROWS = 32;
template <int KIND, int DP>
__global__ void __launch_bounds__(AS_THREADS)
cross_mean_kernel(const double* __restrict__ Xc, int mc, int D, KernelHyp h, const double* __restrict__ XsT,
                  long long ldt, int n, int n_pad, const double* __restrict__ alpha, long long lda, int n_out,
                  double* __restrict__ KsT, long long ldk, double* __restrict__ mean) {
  __shared__ double xc[CR_ROWS][DP];
  __shared__ double red[AS_THREADS / 32][CR_ROWS / 4];
  const int c0 = blockIdx.x * CR_ROWS;
  for (int e = threadIdx.x; e < CR_ROWS * DP; e += AS_THREADS) {
    const int r = e / DP, d = e % DP;
    double v = 0.0;
    if (c0 + r < mc && d < D) v = Xc[(long long)(c0 + r) * D + d] / ell_of(h, d);
    xc[r][d] = v;
  }
  __syncthreads();
  const int cp = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const double sf2 = h.sf * h.sf;
  for (int o = 0; o < max(n_out, 1); ++o) {
    double msum[CR_ROWS / 4];
#pragma unroll
    for (int q = 0; q < CR_ROWS / 4; ++q) msum[q] = 0.0;
    for (int col0 = 0; col0 < n_pad; col0 += AS_TN) {
      const int c = col0 + 2 * cp;
      double xj0[DP], xj1[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const double2 v = *reinterpret_cast<const double2*>(XsT + (long long)d * ldt + c);
        xj0[d] = v.x;
        xj1[d] = v.y;
      }
      const double al0 = (c < n) ? alpha[(long long)o * lda + c] : 0.0;
      const double al1 = (c + 1 < n) ? alpha[(long long)o * lda + c + 1] : 0.0;
#pragma unroll
      for (int q = 0; q < CR_ROWS / 4; ++q) {
        const int rr = rg + 4 * q;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const double x = xc[rr][d];
          const double u0 = x - xj0[d], u1 = x - xj1[d];
          a0 = fma(u0, u0, a0);
          a1 = fma(u1, u1, a1);
        }
        double k0, k1, gd;
        kern_eval<KIND>(a0, sf2, k0, gd);
        kern_eval<KIND>(a1, sf2, k1, gd);
        if (c >= n) k0 = 0.0;
        if (c + 1 >= n) k1 = 0.0;
        if (o == 0 && KsT != nullptr)
          *reinterpret_cast<double2*>(KsT + (long long)(c0 + rr) * ldk + c) = make_double2(k0, k1);
        msum[q] = fma(k0, al0, fma(k1, al1, msum[q]));
      }
    }
    // reduce msum over the 64 column-pair threads of each row group (2 warps per rg), fixed order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < CR_ROWS / 4; ++q) {
      double v = msum[q];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < CR_ROWS) {
      const int rr = threadIdx.x, rgi = rr & 3, q = rr >> 2;
      const double s = red[2 * rgi][q] + red[2 * rgi + 1][q];
      if (c0 + rr < mc && n_out > 0) mean[(long long)(c0 + rr) * n_out + o] = s;
    }
    __syncthreads();
  }
}

// var[c] = max(sf^2 - sum_t part[t][c], 1e-10) (+ sn^2)   (custom_surrogate.py:110-119)
// add_first: the noise term joins before the clamp, as gp_functions.predict_f does (gp_functions.py:484-488).
__global__ void finish_variance_kernel(const double* __restrict__ part, long long ldp, int ntile, int mc, double sf2,
                                       double add, bool add_first, double* __restrict__ var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= mc) return;
  double q = 0.0;
  for (int t = 0; t < ntile; ++t) q += part[(long long)t * ldp + c];
  double v = sf2 - q;
  if (add_first) v = fmax(v + add, 1e-10);
  else v = fmax(v, 1e-10) + add;
  var[c] = v;
}

// Linear-embedding kernel of Garnett (2014) -- python_kernels.LinearEmbedding (python_kernels.py:60-114):
//   K = sf^2 exp(-1/2 |R (x - y)|^2) + sn^2 I,  R (d, D) row-major.
// One thread per element.  dK (na, nb, d*D + 2), ordered [R_00 .. R_(d-1)(D-1), sf, sn] like R.ravel():
//   dK/dR_ab = -K_pure (R delta)_a delta_b   with delta = x - y   (the analytic derivative; the expression at
//   python_kernels.py:108-110 is marked "TODO: check" upstream, only has consistent shapes for d = 1 and is not the
//   derivative there -- see DESIGN.md section 6),   dK/dsf = 2/sf K_pure (:111),  dK/dsn = 2 sn I (:112).
// Not a hot kernel (the embedding kernel is not on the NLL / predict path): plain coalesced K stores, strided dK.
__global__ void __launch_bounds__(256) linear_embedding_kernel(const double* __restrict__ Xa, int na,
                                                               const double* __restrict__ Xb, int nb, int D,
                                                               const double* __restrict__ R, int d, double sf, double sn,
                                                               double* __restrict__ K, double* __restrict__ dK) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)na * nb) return;
  const int i = (int)(e / nb), j = (int)(e % nb);
  double delta[MAX_D], proj[MAX_D];
  for (int b = 0; b < D; ++b) delta[b] = Xa[(long long)i * D + b] - Xb[(long long)j * D + b];
  double d2 = 0.0;
  for (int a = 0; a < d; ++a) {
    // (X R^T)_ia - (Y R^T)_ja evaluated like the reference: project first, then subtract
    double pa = 0.0, pb = 0.0;
    for (int b = 0; b < D; ++b) {
      pa = fma(Xa[(long long)i * D + b], R[a * D + b], pa);
      pb = fma(Xb[(long long)j * D + b], R[a * D + b], pb);
    }
    proj[a] = pa - pb;
    d2 = fma(proj[a], proj[a], d2);
  }
  const double kp = sf * sf * exp(-0.5 * d2);
  const double eye = (i == j) ? 1.0 : 0.0;
  K[e] = kp + sn * sn * eye;
  if (dK != nullptr) {
    double* out = dK + e * (long long)(d * D + 2);
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < D; ++b) out[a * D + b] = -kp * proj[a] * delta[b];
    out[d * D] = 2.0 / sf * kp;
    out[d * D + 1] = 2.0 * sn * eye;
  }
}

}  // namespace gpb
