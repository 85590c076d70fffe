// Acquisition losses and the masked arg-max over a candidate set (profit/al/aquisition_functions.py), evaluated on
// the device from the predictive mean / variance that gpb_predict left there, so that an active-learning pick
// returns 16 bytes instead of 2·M doubles.  All of it is HBM-bound element-wise work plus tree reductions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpb {

constexpr int ACQ_THREADS = 256;
constexpr int ACQ_MAX_OUT = 16;   // output columns summed per candidate (axis=1 of the reference's loss arrays)
constexpr int ACQ_MAX_DIM = 32;

enum AcqKind {
  ACQ_VARIANCE = 0,          // SimpleExploration.calculate_loss                 (aquisition_functions.py:107-119)
  ACQ_DISTANCE_PENALTY = 1,  // ExplorationWithDistancePenalty.calculate_loss    (:146-158)
  ACQ_WEIGHTED = 2,          // WeightedExploration.calculate_loss               (:183-198)
  ACQ_POI = 3,               // ProbabilityOfImprovement.calculate_loss          (:212-220)
  ACQ_EI = 4,                // ExpectedImprovement.calculate_loss / sigma_part  (:253-276)
  ACQ_EI2 = 5                // ExpectedImprovement2.calculate_loss              (:303-315)
};

struct AcqArgs {
  int kind, n_out, d, n_visited;
  long long M;
  double p0, p1;                 // weight | c1 | xi, find_min
  double ymax[ACQ_MAX_OUT];      // per-output incumbent (max / nanmax of the observed outputs)
  double last[ACQ_MAX_DIM];      // last evaluated input point (distance penalty)
};

constexpr double ACQ_EPSILON = 1e-12;        // AcquisitionFunction.EPSILON (:38)
constexpr double ACQ_SIGMA_EPSILON = 1e-10;  // ExpectedImprovement.SIGMA_EPSILON (:240) and the clamp at :309

// np.maximum: NaN wins
__device__ __forceinline__ double np_maximum(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : fmax(a, b); }

// np.argmax order: the first NaN wins, otherwise the largest value, ties to the lowest index
__device__ __forceinline__ bool acq_better(double v1, long long i1, double v2, long long i2) {
  const bool n1 = isnan(v1), n2 = isnan(v2);
  if (n1 != n2) return n1;
  if (!n1 && v1 != v2) return v1 > v2;
  return i1 < i2;
}

__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
__device__ __forceinline__ double norm_pdf(double z) { return exp(-0.5 * z * z) * 0.39894228040143267794; }

// ---- column-wise min / max of an (M, ncol) row-major array; transform 1 takes sqrt first -------------------------
// part[(c * gridDim.x + b) * 2 + {0,1}]
__global__ void __launch_bounds__(ACQ_THREADS)
acq_minmax_partial_kernel(const double* __restrict__ v, long long M, int ncol, int transform, double* __restrict__ part) {
  __shared__ double smin[ACQ_THREADS], smax[ACQ_THREADS];
  const int c = blockIdx.y;
  double lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * ACQ_THREADS + threadIdx.x; i < M; i += (long long)gridDim.x * ACQ_THREADS) {
    double x = v[i * ncol + c];
    if (transform == 1) x = sqrt(x);
    lo = fmin(lo, x);
    hi = fmax(hi, x);
  }
  smin[threadIdx.x] = lo;
  smax[threadIdx.x] = hi;
  __syncthreads();
  for (int off = ACQ_THREADS / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + off]);
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + off]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[((long long)c * gridDim.x + blockIdx.x) * 2] = smin[0];
    part[((long long)c * gridDim.x + blockIdx.x) * 2 + 1] = smax[0];
  }
}

// stats[2c] = min, stats[2c+1] = max of column c
__global__ void __launch_bounds__(ACQ_THREADS)
acq_minmax_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ stats) {
  __shared__ double smin[ACQ_THREADS], smax[ACQ_THREADS];
  const int c = blockIdx.x;
  double lo = INFINITY, hi = -INFINITY;
  for (int b = threadIdx.x; b < nblk; b += ACQ_THREADS) {
    lo = fmin(lo, part[((long long)c * nblk + b) * 2]);
    hi = fmax(hi, part[((long long)c * nblk + b) * 2 + 1]);
  }
  smin[threadIdx.x] = lo;
  smax[threadIdx.x] = hi;
  __syncthreads();
  for (int off = ACQ_THREADS / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + off]);
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + off]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[2 * c] = smin[0];
    stats[2 * c + 1] = smax[0];
  }
}

// ---- the once-per-batch transforms of the predictive mean ---------------------------------------------------------
// pass 0: ACQ_WEIGHTED -> copy;  ACQ_EI -> max(+-mean - ymax_o, 0)          (:264-267)
// pass 1: (x - min_o) / max(max_o - min_o, EPSILON) - shift   (AcquisitionFunction.normalize, :76-84; shift = xi, :268)
__global__ void __launch_bounds__(ACQ_THREADS)
acq_prepare_kernel(int pass, AcqArgs a, const double* __restrict__ mean, const double* __restrict__ stats,
                   double* __restrict__ out) {
  const long long tot = a.M * a.n_out;
  for (long long e = (long long)blockIdx.x * ACQ_THREADS + threadIdx.x; e < tot; e += (long long)gridDim.x * ACQ_THREADS) {
    const int o = (int)(e % a.n_out);
    if (pass == 0) {
      double x = mean[e];
      if (a.kind == ACQ_EI) {
        if (a.p1 != 0.0) x = -x;
        x = np_maximum(x - a.ymax[o], 0.0);
      }
      out[e] = x;
    } else {
      const double lo = stats[2 * o], hi = stats[2 * o + 1];
      const double shift = (a.kind == ACQ_EI) ? a.p0 : 0.0;
      out[e] = (out[e] - lo) / fmax(hi - lo, ACQ_EPSILON) - shift;
    }
  }
}

// ---- loss per candidate + block-level masked arg-max --------------------------------------------------------------
// term : (M, n_out)  predictive mean (POI, EI2) | normalised mean (WEIGHTED) | improvement (EI); unused otherwise
// var  : (M,)        predictive (or marginal) variance
// extra: (M, n_out) standard-normal draws (POI) | (M, d) candidate coordinates (DISTANCE_PENALTY)
// vstats: {min, max} of var (WEIGHTED, DISTANCE_PENALTY) or of sqrt(var) (EI)
// The reference masks with `loss * mask` (mask 0 on visited rows, :66-67): a visited row competes with value 0.
__global__ void __launch_bounds__(ACQ_THREADS)
acq_loss_kernel(AcqArgs a, const double* __restrict__ term, const double* __restrict__ var,
                const double* __restrict__ extra, const double* __restrict__ vstats,
                const long long* __restrict__ visited, double* __restrict__ loss_out, double* __restrict__ pval,
                long long* __restrict__ pidx) {
  __shared__ double sv[ACQ_THREADS];
  __shared__ long long si[ACQ_THREADS];
  double vlo = 0.0, vhi = 0.0;
  if (a.kind == ACQ_DISTANCE_PENALTY || a.kind == ACQ_WEIGHTED || a.kind == ACQ_EI) {
    vlo = vstats[0];
    vhi = vstats[1];
  }
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (long long i = (long long)blockIdx.x * ACQ_THREADS + threadIdx.x; i < a.M; i += (long long)gridDim.x * ACQ_THREADS) {
    const double v = var[i];
    double loss = 0.0;
    switch (a.kind) {
      case ACQ_VARIANCE: loss = v; break;
      case ACQ_DISTANCE_PENALTY: {
        double r2 = 0.0;
        for (int k = 0; k < a.d; ++k) {
          const double u = extra[i * a.d + k] - a.last[k];
          r2 += u * u;
        }
        loss = (v + vhi * (1.0 - exp(-a.p0 * sqrt(r2)))) / 2.0;
      } break;
      case ACQ_WEIGHTED: {
        const double vn = (v - vlo) / fmax(vhi - vlo, ACQ_EPSILON);
        for (int o = 0; o < a.n_out; ++o) loss += a.p0 * term[i * a.n_out + o] + (1.0 - a.p0) * vn;
      } break;
      case ACQ_POI: {
        const double s = sqrt(v);
        for (int o = 0; o < a.n_out; ++o)
          loss += np_maximum(term[i * a.n_out + o] + s * extra[i * a.n_out + o] - a.ymax[o], 0.0);
      } break;
      case ACQ_EI: {
        const double sg = fmax((sqrt(v) - vlo) / fmax(vhi - vlo, ACQ_EPSILON), ACQ_SIGMA_EPSILON);
        for (int o = 0; o < a.n_out; ++o) {
          const double imp = term[i * a.n_out + o];
          const double z = imp / sg;
          loss += imp * norm_cdf(z) + sg * norm_pdf(z);
        }
      } break;
      case ACQ_EI2: {
        const double sg = fmax(sqrt(v), ACQ_SIGMA_EPSILON);
        for (int o = 0; o < a.n_out; ++o) {
          double mu = term[i * a.n_out + o];
          if (a.p1 != 0.0) mu = -mu;
          const double imp = mu - a.ymax[o] - a.p0;
          const double z = imp / sg;
          loss += imp * norm_cdf(z) + sg * norm_pdf(z);
        }
      } break;
      default: break;
    }
    if (loss_out) loss_out[i] = loss;
    double masked = loss;
    for (int k = 0; k < a.n_visited; ++k)
      if (visited[k] == i) masked = loss * 0.0;
    if (acq_better(masked, i, bv, bi)) {
      bv = masked;
      bi = i;
    }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int off = ACQ_THREADS / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off && acq_better(sv[threadIdx.x + off], si[threadIdx.x + off], sv[threadIdx.x], si[threadIdx.x])) {
      sv[threadIdx.x] = sv[threadIdx.x + off];
      si[threadIdx.x] = si[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pval[blockIdx.x] = sv[0];
    pidx[blockIdx.x] = si[0];
  }
}

// Winner over the block partials (also the second stage of gpb_argmax).
__global__ void __launch_bounds__(ACQ_THREADS)
acq_argmax_final_kernel(const double* __restrict__ pval, const long long* __restrict__ pidx, int nblk,
                        double* __restrict__ oval, long long* __restrict__ oidx) {
  __shared__ double sv[ACQ_THREADS];
  __shared__ long long si[ACQ_THREADS];
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (int b = threadIdx.x; b < nblk; b += ACQ_THREADS)
    if (acq_better(pval[b], pidx[b], bv, bi)) {
      bv = pval[b];
      bi = pidx[b];
    }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int off = ACQ_THREADS / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off && acq_better(sv[threadIdx.x + off], si[threadIdx.x + off], sv[threadIdx.x], si[threadIdx.x])) {
      sv[threadIdx.x] = sv[threadIdx.x + off];
      si[threadIdx.x] = si[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    oval[0] = sv[0];
    oidx[0] = si[0];
  }
}

// ---- appending one training point to a factorised model (the reference's TODO at custom_surrogate.py:132) ---------
// With l = L^-1 k (k = k(X, x_new) without noise):  d^2 = kss - |l|^2,  z_o[n] = (y_o - l.z_o) / d.
// scal[0] = 1/d, scal[1 + o] = z_o[n].  A non-positive d^2 sets *info = n + 1 and changes nothing.
__global__ void __launch_bounds__(1024)
append_pivot_kernel(const double* __restrict__ l, long long n, double kss, const double* __restrict__ ynew, int n_out,
                    double* __restrict__ zaug, long long ldz, double* __restrict__ logdet_blk, double* __restrict__ scal,
                    int* __restrict__ info) {
  __shared__ double sh[1024];
  __shared__ double dshare;
  double s = 0.0;
  for (long long k = threadIdx.x; k < n; k += 1024) s = fma(l[k], l[k], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double d2 = kss - sh[0];
    double d = 0.0;
    if (d2 > 0.0) {
      d = sqrt(d2);
      scal[0] = 1.0 / d;
      logdet_blk[0] += log(d);
    } else {
      *info = (int)n + 1;
      scal[0] = 0.0;
    }
    dshare = d;
  }
  __syncthreads();
  const double d = dshare;
  if (d == 0.0) return;
  for (int o = 0; o < n_out; ++o) {
    double t = 0.0;
    for (long long k = threadIdx.x; k < n; k += 1024) t = fma(l[k], zaug[o * ldz + k], t);
    __syncthreads();
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double zn = (ynew[o] - sh[0]) / d;
      scal[1 + o] = zn;
      zaug[o * ldz + n] = zn;
    }
  }
}

// New row of W = L^-1 (= column of U = L^-T):  w_j = -(U l)_j / d  (j < n),  w_n = 1/d;  alpha_o += w * z_o[n].
__global__ void append_update_kernel(const double* __restrict__ t, long long n, const double* __restrict__ scal, int n_out,
                                     double* __restrict__ W, double* __restrict__ U, long long ld,
                                     double* __restrict__ alpha, long long lda) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  const double rd = scal[0];
  const double wj = (j < n) ? -t[j] * rd : rd;
  W[n * ld + j] = wj;
  U[j * ld + n] = wj;
  for (int o = 0; o < n_out; ++o) {
    const double zn = scal[1 + o];
    if (j < n) alpha[o * lda + j] += wj * zn;
    else alpha[o * lda + j] = wj * zn;
  }
}

}  // namespace gpb
