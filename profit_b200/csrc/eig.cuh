// Largest eigenpairs of a dense symmetric matrix on the device -- the twin of scipy.sparse.linalg.eigsh(Ky, neig) on the
// truncated-eigen-decomposition exits of the reference (gp_functions.py:282-284 in negative_log_likelihood, :351/:359 in
// invert).  HBM-bound vector kernels; nothing here is on the throughput path (it only runs where the reference leaves the
// Cholesky branch), so the design goal is determinism and accuracy, not speed:
//
//   * Lanczos with FULL re-orthogonalisation (classical Gram-Schmidt applied twice against every earlier basis vector),
//     basis stored row-wise (Vt: one basis vector per row, so every access is a contiguous row); when the residual
//     vanishes (the Krylov space is invariant: rank-deficient or repeated spectrum) the run continues from a fresh
//     pseudo-random vector orthogonal to the basis, with a zero in the tridiagonal matrix;
//   * the projected tridiagonal matrix T_m is diagonalised on the device by the implicit QL iteration with eigenvector
//     accumulation (one CTA: lane 0 generates the rotations of a sweep, all threads apply them to their rows of Z);
//   * Ritz vectors Q = V Z[:, selected] and the eig-branch NLL pieces (alpha = Q diag(1/w) Q^T y, gp_functions.py:283-287).
//
// All reductions run in a fixed order: results are bit-reproducible from run to run (the reference's ARPACK call starts
// from a random vector and stops at tol = 1e-3 min|l|; its values scatter by ~5e-7 relative, tests/golden/reference_nonspd.json).
#pragma once
#include <cuda_runtime.h>

namespace gpb {

constexpr int EIG_BLOCK = 256;
constexpr int EIG_QL_THREADS = 1024;

__device__ __forceinline__ double eig_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fixed-order block sum (blockDim.x a multiple of 32, <= 1024); the result is valid in every thread.
__device__ __forceinline__ double eig_block_sum(double v, double* red) {
  v = eig_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// w = A v, A symmetric n x n (row-major, leading dimension ld): one warp per row, coalesced row reads.
__global__ void __launch_bounds__(EIG_BLOCK) eig_symv_kernel(const double* __restrict__ A, long long ld, int n,
                                                             const double* __restrict__ v, double* __restrict__ w) {
  const int row = blockIdx.x * (EIG_BLOCK / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = A + (long long)row * ld;
  double s0 = 0.0, s1 = 0.0;
  int c = lane;
  for (; c + 32 < n; c += 64) {
    s0 = fma(a[c], v[c], s0);
    s1 = fma(a[c + 32], v[c + 32], s1);
  }
  if (c < n) s0 = fma(a[c], v[c], s0);
  const double s = eig_warp_sum(s0 + s1);
  if (lane == 0) w[row] = s;
}

// Deterministic start vector (splitmix64 of the index): uniform in (-1, 1).
__global__ void eig_start_kernel(double* __restrict__ w, int n, unsigned long long seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  w[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// c[i] = Vt[i] . w for i < cnt: one block per basis vector.
__global__ void __launch_bounds__(EIG_BLOCK) eig_dots_kernel(const double* __restrict__ Vt, long long ldv, int n,
                                                             const double* __restrict__ w, double* __restrict__ c) {
  __shared__ double red[32];
  const double* v = Vt + (long long)blockIdx.x * ldv;
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_BLOCK) s = fma(v[k], w[k], s);
  s = eig_block_sum(s, red);
  if (threadIdx.x == 0) c[blockIdx.x] = s;
}

// w -= sum_{i < cnt} c[i] Vt[i]; when acc != nullptr (second pass): acc[0] = c_first[cnt-1] + c[cnt-1] = the Lanczos alpha.
__global__ void __launch_bounds__(EIG_BLOCK) eig_update_kernel(const double* __restrict__ Vt, long long ldv, int n, int cnt,
                                                               const double* __restrict__ c, double* __restrict__ w,
                                                               const double* __restrict__ c_first, double* __restrict__ acc) {
  const int k = blockIdx.x * EIG_BLOCK + threadIdx.x;
  if (k == 0 && acc) acc[0] = c_first[cnt - 1] + c[cnt - 1];
  if (k >= n) return;
  double s = w[k];
  for (int i = 0; i < cnt; ++i) s = fma(-c[i], Vt[(long long)i * ldv + k], s);
  w[k] = s;
}

// beta = |w|; dst = w / beta (the next basis vector; nullptr = norm only); beta_out[0] = beta (may be nullptr).  The host
// reads the norms after every step, adds Gram-Schmidt passes while they still drop and replaces the vector when the
// Krylov space turned out to be invariant (see eig_extend).
__global__ void __launch_bounds__(1024) eig_next_kernel(const double* __restrict__ w, int n, double* __restrict__ dst,
                                                        double* __restrict__ beta_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s = fma(w[k], w[k], s);
  s = eig_block_sum(s, red);
  const double beta = sqrt(s);
  if (threadIdx.x == 0 && beta_out) beta_out[0] = beta;
  if (!dst) return;
  const double inv = (beta > 0.0) ? 1.0 / beta : 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = w[k] * inv;
}

// The whole Lanczos recurrence for a SMALL matrix (n <= EIG_SMALL_N) in one CTA: steps [m0, m1) of the same algorithm as
// the kernels above (matrix-vector product, Gram-Schmidt passes until the vector stops shrinking, invariant-subspace
// detection and restart), without a launch or a host round trip per step -- at N = 400 a step costs ~8 us instead of
// ~80 us.  One SM cannot stream a large matrix (8 n^2 bytes per step from L2), hence the size limit.
// Dynamic shared memory: n + m1 doubles (the work vector and the Gram-Schmidt coefficients).
// info[0] += restarts, info[1] <- step at which a non-finite coefficient appeared (0 = none).
constexpr int EIG_SMALL_N = 1024;

__device__ __forceinline__ double eig_hash_uniform(unsigned long long seed, int i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void __launch_bounds__(1024) eig_lanczos_small_kernel(const double* __restrict__ A, int n, double* Vt,
                                                                 double* __restrict__ alpha, double* __restrict__ beta,
                                                                 int m0, int m1, unsigned long long seed,
                                                                 int* __restrict__ info) {
  extern __shared__ double sm[];
  double* w = sm;
  double* c = sm + n;
  __shared__ double red[32];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  auto norm = [&]() {
    double s = 0.0;
    for (int k = tid; k < n; k += nt) s = fma(w[k], w[k], s);
    return sqrt(eig_block_sum(s, red));
  };
  // one classical Gram-Schmidt pass of w against the first cnt basis vectors; returns the coefficient along the last one
  auto gs_pass = [&](int cnt) {
    for (int i = warp; i < cnt; i += nw) {
      const double* v = Vt + (size_t)i * n;
      double s = 0.0;
      for (int k = lane; k < n; k += 32) s = fma(v[k], w[k], s);
      s = eig_warp_sum(s);
      if (lane == 0) c[i] = s;
    }
    __syncthreads();
    const double last = c[cnt - 1];
    for (int k = tid; k < n; k += nt) {
      double s = w[k];
      for (int i = 0; i < cnt; ++i) s = fma(-c[i], Vt[(size_t)i * n + k], s);
      w[k] = s;
    }
    __syncthreads();
    return last;
  };
  // orthogonalise w against the first cnt basis vectors, normalise into row cnt; n0 / n2 = norm before / after
  auto orthonormalise = [&](int cnt, double& coeff, double& n0, double& n2) {
    n0 = norm();
    coeff = gs_pass(cnt);
    const double n1 = norm();
    coeff += gs_pass(cnt);
    n2 = norm();
    double prev = n1;
    for (int extra = 0; extra < 4 && n2 > 0.0 && n2 < 0.5 * prev; ++extra) {
      prev = n2;
      gs_pass(cnt);
      n2 = norm();
    }
    const double inv = (n2 > 0.0) ? 1.0 / n2 : 0.0;
    double* dst = Vt + (size_t)cnt * n;
    for (int k = tid; k < n; k += nt) dst[k] = w[k] * inv;
    __syncthreads();
  };
  for (int j = m0; j < m1; ++j) {
    const double* vj = Vt + (size_t)j * n;
    for (int row = warp; row < n; row += nw) {
      const double* a = A + (size_t)row * n;
      double s = 0.0;
      for (int k = lane; k < n; k += 32) s = fma(a[k], vj[k], s);
      s = eig_warp_sum(s);
      if (lane == 0) w[row] = s;
    }
    __syncthreads();
    double aj, n0, n2;
    orthonormalise(j + 1, aj, n0, n2);
    if (tid == 0) {
      alpha[j] = aj;
      beta[j] = n2;
    }
    if (!(fabs(aj) < 1e300) || !(n2 < 1e300)) {   // NaN / Inf (uniform across the block)
      if (tid == 0) info[1] = j + 1;
      return;
    }
    if (j + 1 < n && n2 <= 64.0 * 2.220446049250313e-16 * n0) {
      if (tid == 0) {
        beta[j] = 0.0;
        info[0] += 1;
      }
      for (int k = tid; k < n; k += nt) w[k] = eig_hash_uniform(seed + 7919ull * (unsigned long long)(j + 1), k);
      __syncthreads();
      double q, q0, q2;
      orthonormalise(j + 1, q, q0, q2);
    }
  }
}

// Eigen-decomposition of the symmetric tridiagonal matrix (alpha[0..m), beta[0..m-1)) by the implicit QL iteration
// (the EISPACK tql2 recurrences).  d[0..m) <- eigenvalues (unsorted); Zt[i * ldz + r] <- component r of eigenvector i.
// One CTA: thread 0 generates the plane rotations of a sweep into shared memory, then every thread applies them to the
// rows of Z it owns (r = tid, tid + blockDim, ...; coalesced because eigenvector i is a contiguous row of Zt).
// Dynamic shared memory: 4 m doubles.  info[0] <- 1 if an eigenvalue needed more than 60 iterations.
__global__ void __launch_bounds__(EIG_QL_THREADS) eig_tql2_kernel(const double* __restrict__ alpha,
                                                                  const double* __restrict__ beta, int m,
                                                                  double* __restrict__ d_out, double* __restrict__ Zt,
                                                                  long long ldz, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* d = sm;
  double* e = sm + m;
  double* cs = sm + 2 * (size_t)m;
  double* sn = sm + 3 * (size_t)m;
  __shared__ double sh_h;
  __shared__ int sh_m, sh_again, sh_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < m; i += nt) {
    d[i] = alpha[i];
    e[i] = (i + 1 < m) ? beta[i] : 0.0;
  }
  for (int i = 0; i < m; ++i)
    for (int r = tid; r < m; r += nt) Zt[(long long)i * ldz + r] = (i == r) ? 1.0 : 0.0;
  if (tid == 0) sh_fail = 0;
  __syncthreads();
  const double eps = 2.220446049250313e-16;
  double f = 0.0, tst1 = 0.0;      // only meaningful in thread 0
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    while (true) {
      // ---- thread 0: find the small sub-diagonal element, form the shift
      if (tid == 0) {
        if (iter == 0) tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
        int mm = l;
        while (mm < m - 1 && fabs(e[mm]) > eps * tst1) ++mm;
        sh_m = mm;
        sh_again = (mm > l && (iter == 0 || fabs(e[l]) > eps * tst1)) ? 1 : 0;
        if (sh_again && iter >= 60) {
          sh_again = 0;
          sh_fail = 1;
        }
        if (sh_again) {
          const double g = d[l];
          double p = (d[l + 1] - g) / (2.0 * e[l]);
          double r = sqrt(p * p + 1.0);
          if (p < 0) r = -r;
          d[l] = e[l] / (p + r);
          d[l + 1] = e[l] * (p + r);
          sh_h = g - d[l];
        }
      }
      __syncthreads();
      if (!sh_again) break;
      const int mm = sh_m;
      const double h = sh_h;
      for (int i = l + 2 + tid; i < m; i += nt) d[i] -= h;
      __syncthreads();
      // ---- thread 0: one implicit QL sweep from mm down to l, rotations to shared memory
      if (tid == 0) {
        f += h;
        const double dl1 = d[l + 1];
        double p = d[mm], c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = mm - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          const double g = c * e[i], hh = c * p;
          const double r = sqrt(p * p + e[i] * e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = hh + s * (c * g + s * d[i]);
          cs[i] = c;
          sn[i] = s;
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      }
      __syncthreads();
      // ---- all threads: apply the rotations (columns i, i+1 of Z = rows of Zt) to the owned components
      for (int r = tid; r < m; r += nt) {
        double carry = Zt[(long long)mm * ldz + r];
        for (int i = mm - 1; i >= l; --i) {
          const double zi = Zt[(long long)i * ldz + r];
          Zt[(long long)(i + 1) * ldz + r] = sn[i] * zi + cs[i] * carry;
          carry = cs[i] * zi - sn[i] * carry;
        }
        Zt[(long long)l * ldz + r] = carry;
      }
      __syncthreads();
      ++iter;
    }
    if (tid == 0) {
      d[l] += f;
      e[l] = 0.0;
    }
    __syncthreads();
  }
  for (int i = tid; i < m; i += nt) d_out[i] = d[i];
  if (tid == 0) info[0] = sh_fail;
}

// Ritz vectors: Qt[i] = sum_j Zt[sel[i]][j] Vt[j]  (i < k, j < m): grid (ceil(n / EIG_BLOCK), k).
__global__ void __launch_bounds__(EIG_BLOCK) eig_ritz_kernel(const double* __restrict__ Vt, long long ldv, int n, int m,
                                                             const double* __restrict__ Zt, long long ldz,
                                                             const int* __restrict__ sel, double* __restrict__ Qt,
                                                             long long ldq) {
  const int col = blockIdx.x * EIG_BLOCK + threadIdx.x;
  if (col >= n) return;
  const double* z = Zt + (long long)sel[blockIdx.y] * ldz;
  double s = 0.0;
  for (int j = 0; j < m; ++j) s = fma(z[j], Vt[(long long)j * ldv + col], s);
  Qt[(long long)blockIdx.y * ldq + col] = s;
}

// Normalise every Ritz vector (they are orthonormal up to rounding; this removes the O(eps m) drift) and form
// t[i] = (Qt[i] . y) / max(w[i], floor)   -- the coefficients of alpha = Q diag(1/w) Q^T y (gp_functions.py:283-284).
// One block per Ritz vector.  qy[i] = Qt[i] . y.
__global__ void __launch_bounds__(EIG_BLOCK) eig_project_kernel(double* __restrict__ Qt, long long ldq, int n,
                                                                const double* __restrict__ y, const double* __restrict__ w,
                                                                double floor_w, double* __restrict__ qy,
                                                                double* __restrict__ t) {
  __shared__ double red[32];
  double* q = Qt + (long long)blockIdx.x * ldq;
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_BLOCK) s = fma(q[k], q[k], s);
  s = eig_block_sum(s, red);
  const double inv = 1.0 / sqrt(s);
  double dot = 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_BLOCK) {
    const double v = q[k] * inv;
    q[k] = v;
    if (y) dot = fma(v, y[k], dot);
  }
  if (!y) return;
  dot = eig_block_sum(dot, red);
  if (threadIdx.x == 0) {
    qy[blockIdx.x] = dot;
    t[blockIdx.x] = dot / fmax(w[blockIdx.x], floor_w);
  }
}

// alpha[c] = sum_i t[i] Qt[i][c]
__global__ void __launch_bounds__(EIG_BLOCK) eig_alpha_kernel(const double* __restrict__ Qt, long long ldq, int n, int k,
                                                              const double* __restrict__ t, double* __restrict__ alpha) {
  const int c = blockIdx.x * EIG_BLOCK + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int i = 0; i < k; ++i) s = fma(t[i], Qt[(long long)i * ldq + c], s);
  alpha[c] = s;
}

// nll = 1/2 (y^T alpha + sum log w + neig_term log 2 pi) with w clipped at floor_w (gp_functions.py:283-287);
// y^T alpha = sum_i qy[i] t[i].  One block, fixed order.
__global__ void __launch_bounds__(EIG_BLOCK) eig_nll_kernel(const double* __restrict__ w, const double* __restrict__ qy,
                                                            const double* __restrict__ t, int k, double floor_w,
                                                            double neig_term, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < k; i += EIG_BLOCK) s += qy[i] * t[i] + log(fmax(w[i], floor_w));
  s = eig_block_sum(s, red);
  if (threadIdx.x == 0) out[0] = 0.5 * (s + neig_term * 1.8378770664093453);   // log(2 pi)
}

// out[a][b] = sum_i Qt[i][a] Qt[i][b] / w[i]   (K^-1 = Q diag(1/w) Q^T, gp_functions.py:369); grid (ceil(n/EIG_BLOCK), rows),
// rows <= 65535: every block row strides over the matrix rows.
__global__ void __launch_bounds__(EIG_BLOCK) eig_outer_kernel(const double* __restrict__ Qt, long long ldq, int n, int k,
                                                              const double* __restrict__ w, double* __restrict__ out) {
  const int b = blockIdx.x * EIG_BLOCK + threadIdx.x;
  if (b >= n) return;
  for (int a = blockIdx.y; a < n; a += gridDim.y) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s = fma(Qt[(long long)i * ldq + a] / w[i], Qt[(long long)i * ldq + b], s);
    out[(long long)a * n + b] = s;
  }
}

}  // namespace gpb
