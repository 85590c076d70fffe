// Blocked diagonal-block kernel of the Cholesky factorisation (round 2): one CTA factorises a 128x128 FP64 block as
// four 32-column panels and inverts the triangular factor behind the factorisation, on the FP64 tensor pipe.
// Replaces the 128-step register sweep of potrf_diag.cuh on the panel chain (reference: np.linalg.cholesky -> LAPACK
// dpotrf, gp_functions.py:143): same inputs and outputs (L_kk in place, W_kk = L_kk^-1, log-determinant, info).
//
// Formulation.  A = Z D Z^T with Z unit lower triangular and D = diag(d) (the pivots); L = Z D^1/2.  With Y = Z D,
//   trailing update after panel p:   A_ij -= Y_ip Z_jp^T                       (DMMA, accumulators stay in registers)
//   inverse  V = Z^-1 block by block: V_qq = Z_qq^-1,  V_qj = -V_qq T_qj,  T_ij += Z_iq V_qj  (i > q >= j)
//   outputs  L_ij = z_ij sqrt(d_j),  W_ij = V_ij / sqrt(d_i).
// No square root or division sits on the dependent chain; only one reciprocal per column does.
//
// Work split (384 threads = 3 warps per scheduler, 168 registers each):
//   warp 8      serial warp: sweeps the 32x32 diagonal sub-block, one lane per row, one column per step (pivot
//               broadcast by shuffle, column all-gather through shared memory; ~100 cycles per column, bound by
//               shuffle + reciprocal + two dependent DFMAs).
//   warps 0..7  workers: hold the trailing matrix and the partial sums T as DMMA accumulators (every worker owns an
//               8x16 strip of every 32x32 block) and run the rank-32 updates and the block products of the inverse.
//   warps 9..11 helpers (no accumulators, so 32-double row buffers fit): solve the panel rows below the diagonal
//               sub-block, one row per thread, replaying the published columns; warp 9 also inverts Z_qq.
// While the serial warp sweeps panel p+1 the workers and warp 9 do the inverse work of panel p, so only the inverse
// of the last panel is left after the factorisation ends.
//
// Shared memory (8-byte fragment loads are conflict-free with a row stride of 36 doubles):
//   Zs  320 x 36  the four panels of Z (panel p: rows 32p..127)          Pb  128 x 36  panel dump / Y rows / V staging
//   Vt  6 x 32 x 36  transposed V blocks (B operands of the T updates)    Cd  32 x 32   published columns of the sweep
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ptx.cuh"

namespace gpb {

constexpr int DR_S = 36;                         // shared-memory row stride in doubles
constexpr int DR_WORKERS = 8;
constexpr int DR_SERIAL = DR_WORKERS;            // warp index of the serial warp
constexpr int DR_HELPER0 = DR_WORKERS + 1;       // first helper warp
constexpr int DR_THREADS = 32 * (DR_WORKERS + 4);
constexpr int DR_ZS = 320 * DR_S, DR_PB = 128 * DR_S, DR_VT = 6 * 32 * DR_S, DR_CD = 32 * 32;
constexpr int DR_SMEM = (DR_ZS + DR_PB + DR_VT + DR_CD + 4 * 128) * (int)sizeof(double);

__host__ __device__ constexpr int dr_zrow0(int p) { return p == 0 ? 0 : (p == 1 ? 128 : (p == 2 ? 224 : 288)); }
__host__ __device__ constexpr int dr_ablk(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }   // bi >= bj, 10 blocks
__host__ __device__ constexpr int dr_tblk(int i, int j) { return i * (i - 1) / 2 + j; }         // i > j, 6 blocks

struct DrSmem {
  double *Zs, *Pb, *Vt, *Cd, *dvec, *invv, *sdv, *rsv;
};

extern __shared__ double dr_sm[];

__device__ __forceinline__ DrSmem dr_smem() {
  DrSmem sm;
  sm.Zs = dr_sm;
  sm.Pb = sm.Zs + DR_ZS;
  sm.Vt = sm.Pb + DR_PB;
  sm.Cd = sm.Vt + DR_VT;
  sm.dvec = sm.Cd + DR_CD;
  sm.invv = sm.dvec + 128;
  sm.sdv = sm.invv + 128;
  sm.rsv = sm.sdv + 128;
  return sm;
}

// rcp.approx.ftz.f64 has ~20 good bits; two Newton steps give full precision (the pivots are far from the
// subnormal range: K + sn^2 I with O(1) entries).
__device__ __forceinline__ double dr_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ void dr_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(DR_WORKERS * 32) : "memory"); }
__device__ __forceinline__ void dr_bar_all() { asm volatile("bar.sync 0, %0;" ::"n"(DR_THREADS) : "memory"); }
// workers + the inverting helper warp
__device__ __forceinline__ void dr_bar_inverse() { asm volatile("bar.sync 2, %0;" ::"n"(DR_WORKERS * 32 + 32) : "memory"); }

// ---- serial warp: factor the 32x32 diagonal sub-block of panel p (rows 0..31 of Pb), publish its columns.
__device__ __forceinline__ void dr_sweep(int p, int lane, int k0, int n_valid, int* info) {
  const DrSmem sm = dr_smem();
  double a[32];
  const double* prow = sm.Pb + lane * DR_S;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    const double2 v = *reinterpret_cast<const double2*>(prow + 2 * h);
    a[2 * h] = v.x;
    a[2 * h + 1] = v.y;
  }
  double* Cd = sm.Cd;
  Cd[lane] = a[0];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    // column j as published at the end of step j-1 (entries k > j are final)
    double c[32];
#pragma unroll
    for (int h = (j + 1) / 2; h < 16; ++h) {
      const double2 v = *reinterpret_cast<const double2*>(Cd + j * 32 + 2 * h);
      c[2 * h] = v.x;
      c[2 * h + 1] = v.y;
    }
    const double inv = dr_rcp(d);
    const double f = a[j] * inv;
    a[j] = f;   // z_ij for the lanes below the diagonal
    if (j < 31) {
      a[j + 1] = fma(-f, c[j + 1], a[j + 1]);
      Cd[(j + 1) * 32 + lane] = a[j + 1];
      __syncwarp();
    }
    if (lane == 0) {
      sm.dvec[32 * p + j] = d;
      sm.invv[32 * p + j] = inv;
      if (!(d > 0.0) && (k0 + 32 * p + j) < n_valid) atomicCAS(info, 0, k0 + 32 * p + j + 1);
    }
#pragma unroll
    for (int k = j + 2; k < 32; ++k) a[k] = fma(-f, c[k], a[k]);
  }
  // unit-lower Z_pp into rows 0..31 of Z panel p
  double* zrow = sm.Zs + (dr_zrow0(p) + lane) * DR_S;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    double2 v;
    v.x = (2 * h < lane) ? a[2 * h] : ((2 * h == lane) ? 1.0 : 0.0);
    v.y = (2 * h + 1 < lane) ? a[2 * h + 1] : ((2 * h + 1 == lane) ? 1.0 : 0.0);
    *reinterpret_cast<double2*>(zrow + 2 * h) = v;
  }
}

// ---- workers: one thread per panel row below the diagonal sub-block; replays the published columns.
// Pb row lr (>= 32): in  A_ip row (after the updates of the earlier panels);  out  Y row (in place), Z row (Zs).
__device__ __forceinline__ void dr_trsm_row(int p, int lr) {
  const DrSmem sm = dr_smem();
  double a[32];
  double* prow = sm.Pb + lr * DR_S;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    const double2 v = *reinterpret_cast<const double2*>(prow + 2 * h);
    a[2 * h] = v.x;
    a[2 * h + 1] = v.y;
  }
  const double* Cd = sm.Cd;
  const double* invp = sm.invv + 32 * p;
  double* zrow = sm.Zs + (dr_zrow0(p) + lr) * DR_S;
  // two columns per step; the Z entries leave for shared memory at once (the accumulators stay live in registers)
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const double2 iv = *reinterpret_cast<const double2*>(invp + j);
    const double f0 = a[j] * iv.x;
    a[j + 1] = fma(-f0, Cd[j * 32 + j + 1], a[j + 1]);
    const double f1 = a[j + 1] * iv.y;
    *reinterpret_cast<double2*>(zrow + j) = make_double2(f0, f1);
#pragma unroll
    for (int h = j / 2 + 1; h < 16; ++h) {
      const double2 c0 = *reinterpret_cast<const double2*>(Cd + j * 32 + 2 * h);
      const double2 c1 = *reinterpret_cast<const double2*>(Cd + (j + 1) * 32 + 2 * h);
      a[2 * h] = fma(-f1, c1.x, fma(-f0, c0.x, a[2 * h]));
      a[2 * h + 1] = fma(-f1, c1.y, fma(-f0, c0.y, a[2 * h + 1]));
    }
  }
#pragma unroll
  for (int h = 0; h < 16; ++h) *reinterpret_cast<double2*>(prow + 2 * h) = make_double2(a[2 * h], a[2 * h + 1]);
}

// ---- workers: C fragments of column-block P of the trailing matrix -> Pb (local rows 32 (i - P) ..)
template <int P>
__device__ __forceinline__ void dr_dump(const DrSmem& sm, double (&accA)[10][2][2], int tr, int tcb, int g, int t) {
#pragma unroll
  for (int i = P; i < 4; ++i)
#pragma unroll
    for (int tc = 0; tc < 2; ++tc) {
      double* dst = sm.Pb + (32 * (i - P) + 8 * tr + g) * DR_S + 8 * (tcb + tc) + 2 * t;
      *reinterpret_cast<double2*>(dst) = make_double2(accA[dr_ablk(i, P)][tc][0], accA[dr_ablk(i, P)][tc][1]);
    }
}

// ---- workers: A_ij -= Y_iP Z_jP^T for i >= j > P
template <int P>
__device__ __forceinline__ void dr_update(const DrSmem& sm, double (&accA)[10][2][2], int tr, int tcb, int g, int t) {
  const double* Zp = sm.Zs + dr_zrow0(P) * DR_S;
#pragma unroll
  for (int bj = P + 1; bj < 4; ++bj) {
    double bz[2][8];
#pragma unroll
    for (int tc = 0; tc < 2; ++tc)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) bz[tc][ks] = Zp[(32 * (bj - P) + 8 * (tcb + tc) + g) * DR_S + 4 * ks + t];
#pragma unroll
    for (int bi = bj; bi < 4; ++bi) {
      double ay[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) ay[ks] = -sm.Pb[(32 * (bi - P) + 8 * tr + g) * DR_S + 4 * ks + t];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc)
          dmma884(accA[dr_ablk(bi, bj)][tc][0], accA[dr_ablk(bi, bj)][tc][1], ay[ks], bz[tc][ks]);
    }
  }
}

// ---- helper warp 9: V_qq = Z_qq^-1 by forward substitution (lane m = column m), sqrt(d) and 1/sqrt(d) of panel q.
// Runs next to the workers' transposed dump of T_qj; both meet at dr_bar_inverse().
__device__ __forceinline__ void dr_inverse_diag(int q, int lane, double* __restrict__ Winv) {
  const DrSmem sm = dr_smem();
  double* Vd = sm.Pb + 96 * DR_S;
  {
    const double sd = sqrt(sm.dvec[32 * q + lane]);
    sm.sdv[32 * q + lane] = sd;
    sm.rsv[32 * q + lane] = 1.0 / sd;
  }
  __syncwarp();
  double v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (i == lane) ? 1.0 : 0.0;
  const double* Zq = sm.Zs + dr_zrow0(q) * DR_S;
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    v[k + 1] = fma(-Zq[(k + 1) * DR_S + k], v[k], v[k + 1]);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i >= k + 2) {   // constant trip count so that the nest unrolls completely and v[] stays in registers
        const double2 zz = *reinterpret_cast<const double2*>(Zq + i * DR_S + k);
        v[i] = fma(-zz.x, v[k], v[i]);
        v[i] = fma(-zz.y, v[k + 1], v[i]);
      }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    Vd[i * DR_S + lane] = v[i];
    Winv[(32 * q + i) * 128 + 32 * q + lane] = sm.rsv[32 * q + i] * v[i];
  }
  if (q < 3) {
    double* vt = sm.Vt + (dr_ablk(q, q) * 32 + lane) * DR_S;
#pragma unroll
    for (int h = 0; h < 16; ++h) *reinterpret_cast<double2*>(vt + 2 * h) = make_double2(v[2 * h], v[2 * h + 1]);
  }
  dr_bar_inverse();
}

// ---- workers: inverse work of panel Q (runs while the serial warp sweeps panel Q+1; Q = 3 after the last sweep).
// Staging in the rows of Pb the next dump does not use: slot 3 = V_QQ row-major, slot 2-j = T_Qj transposed.
template <int Q>
__device__ __forceinline__ void dr_inverse(const DrSmem& sm, double (&accT)[6][2][2], int tr, int tcb, int g, int t,
                                           double* __restrict__ Winv) {
  const double* Vd = sm.Pb + 96 * DR_S;
  // (a) transposed dump of T_Qj, j < Q
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    double* Tt = sm.Pb + 32 * (2 - j) * DR_S;
#pragma unroll
    for (int tc = 0; tc < 2; ++tc)
#pragma unroll
      for (int e = 0; e < 2; ++e) Tt[(8 * (tcb + tc) + 2 * t + e) * DR_S + 8 * tr + g] = accT[dr_tblk(Q, j)][tc][e];
  }
  dr_bar_inverse();   // V_QQ (helper warp) and the dumps are in place
  // (c) V_Qj = -V_QQ T_Qj for j < Q; V_QQ is lower triangular: k-steps up to the row tile only
  if (Q > 0) {
    double av[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) av[ks] = (ks < 2 * (tr + 1)) ? -Vd[(8 * tr + g) * DR_S + 4 * ks + t] : 0.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      const double* Tt = sm.Pb + 32 * (2 - j) * DR_S;
      double cv[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        if (ks < 2 * (tr + 1)) {
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)
            dmma884(cv[tc][0], cv[tc][1], av[ks], Tt[(8 * (tcb + tc) + g) * DR_S + 4 * ks + t]);
        }
      const int row = 8 * tr + g;
      const double rs = sm.rsv[32 * Q + row];
#pragma unroll
      for (int tc = 0; tc < 2; ++tc) {
        const int col = 8 * (tcb + tc) + 2 * t;
        if (Q < 3) {
          double* vt = sm.Vt + dr_ablk(Q, j) * 32 * DR_S;
          vt[col * DR_S + row] = cv[tc][0];
          vt[(col + 1) * DR_S + row] = cv[tc][1];
        }
        *reinterpret_cast<double2*>(Winv + (32 * Q + row) * 128 + 32 * j + col) = make_double2(rs * cv[tc][0], rs * cv[tc][1]);
      }
    }
    if (Q < 3) dr_bar_workers();
  }
  // (e) T_ij += Z_iQ V_Qj for i > Q >= j
  if (Q < 3) {
    const double* Zq = sm.Zs + dr_zrow0(Q) * DR_S;
#pragma unroll
    for (int i = Q + 1; i < 4; ++i) {
      double az[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) az[ks] = Zq[(32 * (i - Q) + 8 * tr + g) * DR_S + 4 * ks + t];
#pragma unroll
      for (int j = 0; j <= Q; ++j) {
        const double* vt = sm.Vt + dr_ablk(Q, j) * 32 * DR_S;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)
            dmma884(accT[dr_tblk(i, j)][tc][0], accT[dr_tblk(i, j)][tc][1], az[ks],
                    vt[(8 * (tcb + tc) + g) * DR_S + 4 * ks + t]);
      }
    }
  }
}

// In / out exactly as potrf_diag4_kernel:
//   A (row-major, lda): lower triangle of the SPD block at rows/cols [k0, k0+128)  ->  L (strict upper zeroed)
//   Winv <- L^-1 (128x128 row-major, strict upper zeroed);  logdet[0] <- sum_j log L_jj over rows < n_valid;
//   info <- first non-positive pivot (global 1-based index), untouched otherwise.
__global__ void __launch_bounds__(DR_THREADS, 1)
potrf_diagr_kernel(double* __restrict__ A, long long lda, int k0, double* __restrict__ Winv, double* __restrict__ logdet,
                   int* __restrict__ info, int n_valid) {
  const DrSmem sm = dr_smem();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < DR_WORKERS;
  const int g = lane >> 2, t = lane & 3, tr = warp >> 1, tcb = 2 * (warp & 1);
  double* Ablk = A + (long long)k0 * lda + k0;
  pdl_launch_dependents();
  pdl_wait();
#ifdef GPB_TRACE
  const unsigned long long trace_t0 = trace_now();
#endif

  // Three role programs; every role passes the same sequence of CTA-wide barriers.
  if (worker) {
    double accA[10][2][2], accT[6][2][2];
#pragma unroll
    for (int bi = 0; bi < 4; ++bi)
#pragma unroll
      for (int bj = 0; bj <= bi; ++bj)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc) {
          const double2 v = *reinterpret_cast<const double2*>(Ablk + (long long)(32 * bi + 8 * tr + g) * lda + 32 * bj +
                                                              8 * (tcb + tc) + 2 * t);
          accA[dr_ablk(bi, bj)][tc][0] = v.x;
          accA[dr_ablk(bi, bj)][tc][1] = v.y;
        }
#pragma unroll
    for (int b = 0; b < 6; ++b)
#pragma unroll
      for (int tc = 0; tc < 2; ++tc) accT[b][tc][0] = accT[b][tc][1] = 0.0;
#define GPB_DR_PANEL(P)                                                                 \
  dr_dump<P>(sm, accA, tr, tcb, g, t);                                                  \
  dr_bar_all(); /* panel P is in Pb */                                                  \
  if (P > 0) dr_inverse<(P > 0 ? P - 1 : 0)>(sm, accT, tr, tcb, g, t, Winv);            \
  dr_bar_all(); /* sweep(P) done */                                                     \
  if (P < 3) {                                                                          \
    dr_bar_all(); /* panel rows solved */                                               \
    dr_update<(P < 3 ? P : 2)>(sm, accA, tr, tcb, g, t);                                \
    dr_bar_all(); /* Pb may be overwritten */                                           \
  }
    GPB_DR_PANEL(0) GPB_DR_PANEL(1) GPB_DR_PANEL(2) GPB_DR_PANEL(3)
#undef GPB_DR_PANEL
    dr_inverse<3>(sm, accT, tr, tcb, g, t, Winv);
  } else if (warp == DR_SERIAL) {
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
      dr_bar_all();
      dr_sweep(p, lane, k0, n_valid, info);
      dr_bar_all();
      if (p < 3) {
        dr_bar_all();
        dr_bar_all();
      }
    }
  } else {
    const int hrow = tid - 32 * DR_HELPER0;   // 0..95
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
      dr_bar_all();
      if (p > 0 && warp == DR_HELPER0) dr_inverse_diag(p - 1, lane, Winv);
      dr_bar_all();
      if (p < 3) {
        if (hrow < 96 - 32 * p) dr_trsm_row(p, 32 + hrow);
        dr_bar_all();
        dr_bar_all();
      }
    }
    if (warp == DR_HELPER0) dr_inverse_diag(3, lane, Winv);
  }
  __syncthreads();

  // ---- outputs: L = Z diag(sqrt d) in place (strict upper zeroed), zeros above the block diagonal of Winv, logdet
  for (int e = tid; e < 128 * 128; e += DR_THREADS) {
    const int r = e >> 7, c = e & 127;
    double l = 0.0;
    if (c < r) {
      const int pc = c >> 5;
      l = sm.Zs[(dr_zrow0(pc) + r - 32 * pc) * DR_S + (c & 31)] * sm.sdv[c];
    } else if (c == r) {
      l = sm.sdv[r];
    }
    Ablk[(long long)r * lda + c] = l;
    if ((c >> 5) > (r >> 5)) Winv[r * 128 + c] = 0.0;
  }
  if (logdet != nullptr && warp == 4) {
    double s = 0.0;
    for (int q = 0; q < 4; ++q) {
      const int j = lane + 32 * q;
      if (k0 + j < n_valid) s += 0.5 * log(sm.dvec[j]);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) logdet[0] = s;
  }
#ifdef GPB_TRACE
  if (tid == 0) trace_emit(trace_t0, 1u, 0xD1A9u);
#endif
}

}  // namespace gpb
