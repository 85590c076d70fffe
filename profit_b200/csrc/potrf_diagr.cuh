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
//               sub-block, one row per thread, replaying the published columns four at a time WHILE the sweep is
//               still running (the serial warp signals every fourth column on an mbarrier); warp 11 inverts Z_qq.
// While the serial warp sweeps panel p+1 the workers and warp 11 do the inverse work of panel p, so only the inverse
// of the last panel is left after the factorisation ends.
//
// Shared memory (8-byte fragment loads are conflict-free with a row stride of 36 doubles):
//   Zs  320 x 36  the four panels of Z (panel p: rows 32p..127)          Pb  128 x 36  panel dump / Y rows / V staging
//   Vt  6 x 32 x 36  transposed V blocks (B operands of the T updates)    Cd  2 x 32 x 32  published columns of the sweep
//                    (panel 0 keeps its own copy: its diagonal inverse is taken one panel later)
//   Vd  2 x 32 x 36  V_qq row-major (A operand of V_qj = -V_qq T_qj), double-buffered
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ptx.cuh"

namespace gpb {

constexpr int DR_S = 36;                         // shared-memory row stride in doubles
constexpr int DR_WORKERS = 8;
constexpr int DR_SERIAL = DR_WORKERS;            // warp index of the serial warp
constexpr int DR_HELPER0 = DR_WORKERS + 1;       // first helper warp
constexpr int DR_INVWARP = DR_WORKERS + 3;       // helper warp that inverts the diagonal sub-blocks of Z
constexpr int DR_THREADS = 32 * (DR_WORKERS + 4);
constexpr int DR_ZS = 320 * DR_S, DR_PB = 128 * DR_S, DR_VT = 6 * 32 * DR_S, DR_CD = 2 * 32 * 32;
constexpr int DR_VD = 2 * 32 * DR_S;              // V_qq row-major, double-buffered by q & 1
constexpr int DR_SMEM = (DR_ZS + DR_PB + DR_VT + DR_CD + DR_VD + 4 * 128 + 9) * (int)sizeof(double);   // + 9 mbarriers

__host__ __device__ constexpr int dr_zrow0(int p) { return p == 0 ? 0 : (p == 1 ? 128 : (p == 2 ? 224 : 288)); }
__host__ __device__ constexpr int dr_ablk(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }   // bi >= bj, 10 blocks
__host__ __device__ constexpr int dr_tblk(int i, int j) { return i * (i - 1) / 2 + j; }         // i > j, 6 blocks

struct DrSmem {
  double *Zs, *Pb, *Vt, *Cd, *Vd, *dvec, *invv, *sdv, *rsv;
  uint64_t* grp;   // mbarrier g < 8: columns 4g..4g+3 of the running sweep are published (phase parity = panel & 1);
                   // mbarrier 8: V_00 is in place
};

extern __shared__ double dr_sm[];

#ifdef GPB_DIAG_CLOCKS
// Debug build (tools/diagr_bench.cu): clock64 stamps of the three role programs, [role][slot].
__device__ long long g_dr_clk[3][64];
__device__ long long g_dr_step[4][36];
#define DR_STAMP(role, slot)                                  \
  do {                                                        \
    if ((threadIdx.x & 31) == 0) g_dr_clk[role][slot] = clock64(); \
  } while (0)
#else
#define DR_STAMP(role, slot) do {} while (0)
#endif

__device__ __forceinline__ DrSmem dr_smem() {
  DrSmem sm;
  sm.Zs = dr_sm;
  sm.Pb = sm.Zs + DR_ZS;
  sm.Vt = sm.Pb + DR_PB;
  sm.Cd = sm.Vt + DR_VT;
  sm.Vd = sm.Cd + DR_CD;
  sm.dvec = sm.Vd + DR_VD;
  sm.invv = sm.dvec + 128;
  sm.sdv = sm.invv + 128;
  sm.rsv = sm.sdv + 128;
  sm.grp = reinterpret_cast<uint64_t*>(sm.rsv + 128);
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

// ---- serial warp: factor the 32x32 diagonal sub-block of panel p (rows 0..31 of Pb), publish its columns.
__device__ __forceinline__ void dr_sweep(int p, int lane, int k0, int n_valid, int* info) {
  const DrSmem sm = dr_smem();
#ifdef GPB_DIAG_CLOCKS
  if (lane == 0) g_dr_step[p][33] = clock64();
#endif
  double a[32];
  const double* prow = sm.Pb + lane * DR_S;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    const double2 v = *reinterpret_cast<const double2*>(prow + 2 * h);
    a[2 * h] = v.x;
    a[2 * h + 1] = v.y;
  }
  double* Cd = sm.Cd + (p ? 1024 : 0);
  Cd[lane] = a[0];
  __syncwarp();
  double my_d = 1.0;   // pivot of column `lane`
#pragma unroll
  for (int j = 0; j < 32; ++j) {
#ifdef GPB_DIAG_CLOCKS
    if (lane == 0) g_dr_step[p][j] = clock64();
#endif
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
    }
    if (lane == j) {
      sm.invv[32 * p + j] = inv;
      sm.dvec[32 * p + j] = d;
      my_d = d;
    }
    __syncwarp();
    if ((j & 3) == 3 && lane == 0) mbar_arrive(&sm.grp[j >> 2]);   // columns .. j: rows of Cd and reciprocals are out
#pragma unroll
    for (int k = j + 2; k < 32; ++k) a[k] = fma(-f, c[k], a[k]);
  }
#ifdef GPB_DIAG_CLOCKS
  if (lane == 0) g_dr_step[p][32] = clock64();
#endif
  {
    // LAPACK-style info: the first non-positive pivot of this panel (lowest lane wins)
    const unsigned bad = __ballot_sync(0xffffffffu, !(my_d > 0.0) && (k0 + 32 * p + lane) < n_valid);
    if (bad != 0u && lane == __ffs(bad) - 1) atomicCAS(info, 0, k0 + 32 * p + lane + 1);
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
#ifdef GPB_DIAG_CLOCKS
  if (lane == 0) g_dr_step[p][34] = clock64();
#endif
}

// ---- helpers: one thread per row; replays the columns the sweep publishes, four at a time, while it is running.
// Panel rows below the diagonal sub-block (inverse == false): Pb row lr holds the A_ip row after the updates of the
// earlier panels; out: Y row (in place) and Z row (Zs).
// Identity rows (inverse == true, lane m starts from e_m): the same recurrence yields column m of V_pp = Z_pp^-1
// (v_i -= c_ik (v_k / d_k) is the forward substitution with Z_ik = c_ik / d_k); out: V_pp row-major (Vd), transposed
// (Vt, B operand of the T updates), W_pp = diag(1/sqrt d) V_pp (global) and sqrt(d), 1/sqrt(d) of the panel.
__device__ __forceinline__ void dr_solve_row(int p, int lr, bool inverse, bool wait, int lane, double* __restrict__ Winv) {
  const DrSmem sm = dr_smem();
  double a[32];
  double* prow = sm.Pb + lr * DR_S;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    double2 v = *reinterpret_cast<const double2*>(prow + 2 * h);
    if (inverse) v = make_double2((2 * h == lane) ? 1.0 : 0.0, (2 * h + 1 == lane) ? 1.0 : 0.0);
    a[2 * h] = v.x;
    a[2 * h + 1] = v.y;
  }
  const double* Cd = sm.Cd + (p ? 1024 : 0);
  const double* invp = sm.invv + 32 * p;
  double* zrow = sm.Zs + (dr_zrow0(p) + lr) * DR_S;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    if ((j & 3) == 0 && wait) mbar_wait(&sm.grp[j >> 2], (uint32_t)(p & 1));   // the sweep has published columns j..j+3
    const double2 iv = *reinterpret_cast<const double2*>(invp + j);
    const double f0 = a[j] * iv.x;
    a[j + 1] = fma(-f0, Cd[j * 32 + j + 1], a[j + 1]);
    const double f1 = a[j + 1] * iv.y;
    if (!inverse) *reinterpret_cast<double2*>(zrow + j) = make_double2(f0, f1);
#pragma unroll
    for (int h = j / 2 + 1; h < 16; ++h) {
      const double2 c0 = *reinterpret_cast<const double2*>(Cd + j * 32 + 2 * h);
      const double2 c1 = *reinterpret_cast<const double2*>(Cd + (j + 1) * 32 + 2 * h);
      a[2 * h] = fma(-f1, c1.x, fma(-f0, c0.x, a[2 * h]));
      a[2 * h + 1] = fma(-f1, c1.y, fma(-f0, c0.y, a[2 * h + 1]));
    }
  }
  if (!inverse) {
#pragma unroll
    for (int h = 0; h < 16; ++h) *reinterpret_cast<double2*>(prow + 2 * h) = make_double2(a[2 * h], a[2 * h + 1]);
    return;
  }
  {
    const double sd = sqrt(sm.dvec[32 * p + lane]);
    sm.sdv[32 * p + lane] = sd;
    sm.rsv[32 * p + lane] = 1.0 / sd;
  }
  __syncwarp();
  double* Vd = sm.Vd + (p & 1) * 32 * DR_S;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    Vd[i * DR_S + lane] = a[i];
    Winv[(32 * p + i) * 128 + 32 * p + lane] = sm.rsv[32 * p + i] * a[i];
  }
  if (p < 3) {
    double* vt = sm.Vt + (dr_ablk(p, p) * 32 + lane) * DR_S;
#pragma unroll
    for (int h = 0; h < 16; ++h) *reinterpret_cast<double2*>(vt + 2 * h) = make_double2(a[2 * h], a[2 * h + 1]);
  }
  if (p == 0) {   // the workers' inverse<0> waits for this
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.grp[8]);
  }
}

// ---- workers (256 threads), factor-only variant: columns 32p..32p+31 of L = Z diag(sqrt d), rows 32p..127, to global
// memory while the serial warp sweeps the next panel (the workers have nothing else to do then, and no DMMA stream
// competes with the scaling multiplications).  Only the lower triangle is written.
__device__ __forceinline__ void dr_store_panel(int p, int wt, double* __restrict__ Ablk, long long lda) {
  const DrSmem sm = dr_smem();
  const int cl = (wt & 15) * 2;                     // column pair inside the panel
  const double sd0 = sqrt(sm.dvec[32 * p + cl]), sd1 = sqrt(sm.dvec[32 * p + cl + 1]);
  const double* zcol = sm.Zs + dr_zrow0(p) * DR_S + cl;
  double* out = Ablk + (long long)(32 * p) * lda + 32 * p + cl;
#pragma unroll 4
  for (int lr = wt >> 4; lr < 128 - 32 * p; lr += 16) {
    if (cl <= lr) {
      const double2 z = *reinterpret_cast<const double2*>(zcol + lr * DR_S);
      double2 l;
      l.x = (cl < lr) ? z.x * sd0 : sd0;
      l.y = (cl + 1 < lr) ? z.y * sd1 : ((cl + 1 == lr) ? sd1 : 0.0);
      *reinterpret_cast<double2*>(out + (long long)lr * lda) = l;
    }
  }
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

// ---- workers: inverse work of panel Q (runs while the serial warp sweeps panel Q+1; Q = 3 after the last sweep).
// V_QQ (Vd, Vt) and 1/sqrt(d) of the panel are ready: warp 11 produced them behind sweep(Q).  Staging of the
// transposed T_Qj in the rows of Pb the next dump does not use (slot 2-j).
template <int Q>
__device__ __forceinline__ void dr_inverse(const DrSmem& sm, double (&accT)[6][2][2], int tr, int tcb, int g, int t,
                                           double* __restrict__ Winv) {
  const double* Vd = sm.Vd + (Q & 1) * 32 * DR_S;
  if (Q == 0) mbar_wait(&sm.grp[8], 0u);   // V_00 comes from warp 11 during this very sweep
  // (a) transposed dump of T_Qj, j < Q
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    double* Tt = sm.Pb + 32 * (2 - j) * DR_S;
#pragma unroll
    for (int tc = 0; tc < 2; ++tc)
#pragma unroll
      for (int e = 0; e < 2; ++e) Tt[(8 * (tcb + tc) + 2 * t + e) * DR_S + 8 * tr + g] = accT[dr_tblk(Q, j)][tc][e];
  }
  if (Q > 0) dr_bar_workers();   // the dumps are in place
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

// In / out as potrf_diag4_kernel:
//   A (row-major, lda): lower triangle of the SPD block at rows/cols [k0, k0+128)  ->  L (lower triangle)
//   Winv <- L^-1 (128x128 row-major, lower block triangle);  logdet[0] <- sum_j log L_jj over rows < n_valid;
//   info <- first non-positive pivot (global 1-based index), untouched otherwise.
// FULL_INV == false (the blocked Cholesky's panel chain): only the four 32x32 diagonal blocks of L^-1 are produced
// -- the panel solve is a blocked triangular solve (trsm_panel_kernel) that needs nothing else, and
// complete_inverse_kernel fills the rest of L^-1 off the chain when the triangular inverse asks for it.  The workers
// then have no inverse work and write the finished panel of L while the next one is being swept.
template <bool FULL_INV>
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
  if (warp == 0) DR_STAMP(0, 0);

  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 9; ++q) mbar_init(&sm.grp[q], 1);
  }
  __syncthreads();

  // Three role programs.  CTA-wide barriers per panel: S1 "panel is in Pb", S2 "sweep and panel solve done".
  if (worker) {
    double accA[10][2][2], accT[6][2][2];
    auto load_block = [&](int bi, int bj, double (&dst)[2][2]) {
#pragma unroll
      for (int tc = 0; tc < 2; ++tc) {
        const double2 v = *reinterpret_cast<const double2*>(Ablk + (long long)(32 * bi + 8 * tr + g) * lda + 32 * bj +
                                                            8 * (tcb + tc) + 2 * t);
        dst[tc][0] = v.x;
        dst[tc][1] = v.y;
      }
    };
    // the first panel first: the sweep starts while the other six blocks are still on their way
#pragma unroll
    for (int bi = 0; bi < 4; ++bi) load_block(bi, 0, accA[dr_ablk(bi, 0)]);
    dr_dump<0>(sm, accA, tr, tcb, g, t);
    dr_bar_all();   // S1(0)
    if (warp == 0) DR_STAMP(0, 1);
#pragma unroll
    for (int bi = 1; bi < 4; ++bi)
#pragma unroll
      for (int bj = 1; bj <= bi; ++bj) load_block(bi, bj, accA[dr_ablk(bi, bj)]);
#pragma unroll
    for (int b = 0; b < 6; ++b)
#pragma unroll
      for (int tc = 0; tc < 2; ++tc) accT[b][tc][0] = accT[b][tc][1] = 0.0;
#define GPB_DR_PANEL(P)                                                                 \
  if (P > 0) {                                                                          \
    dr_dump<P>(sm, accA, tr, tcb, g, t);                                                \
    dr_bar_all(); /* S1(P) */                                                           \
    if (warp == 0) DR_STAMP(0, 8 * P + 1);                                              \
    if (FULL_INV) dr_inverse<(P > 0 ? P - 1 : 0)>(sm, accT, tr, tcb, g, t, Winv);       \
    else dr_store_panel((P > 0 ? P - 1 : 0), tid, Ablk, lda);                           \
    if (warp == 0) DR_STAMP(0, 8 * P + 2);                                              \
  }                                                                                     \
  dr_bar_all(); /* S2(P) */                                                             \
  if (warp == 0) DR_STAMP(0, 8 * P + 3);                                                \
  if (P < 3) {                                                                          \
    dr_update<(P < 3 ? P : 2)>(sm, accA, tr, tcb, g, t);                                \
    if (warp == 0) DR_STAMP(0, 8 * P + 4);                                              \
    dr_bar_workers(); /* Y rows consumed: Pb may be overwritten by the next dump */     \
  }
    GPB_DR_PANEL(0) GPB_DR_PANEL(1) GPB_DR_PANEL(2) GPB_DR_PANEL(3)
#undef GPB_DR_PANEL
    if (FULL_INV) dr_inverse<3>(sm, accT, tr, tcb, g, t, Winv);
    else dr_store_panel(3, tid, Ablk, lda);
    if (warp == 0) DR_STAMP(0, 41);
  } else if (warp == DR_SERIAL) {
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
      dr_bar_all();   // S1(p)
      dr_sweep(p, lane, k0, n_valid, info);
      dr_bar_all();   // S2(p)
    }
  } else {
    const int hrow = tid - 32 * DR_HELPER0;   // 0..95
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
      dr_bar_all();   // S1(p)
      if (warp == DR_HELPER0) DR_STAMP(2, 4 * p);
      // Panel rows (96, 64, 32, 0 of them).  Warp 11 is free from panel 1 on: it inverts Z_pp behind the sweep, after
      // catching up on Z_00 (whose published columns have their own buffer) when p == 1.
      if (warp == DR_INVWARP && p > 0) {
        if (p == 1) dr_solve_row(0, 32 + hrow, true, false, lane, Winv);
        dr_solve_row(p, 32 + hrow, true, true, lane, Winv);
      } else if (hrow < 96 - 32 * p) {
        dr_solve_row(p, 32 + hrow, false, true, lane, Winv);
      }
      if (warp == DR_HELPER0) DR_STAMP(2, 4 * p + 1);
      dr_bar_all();   // S2(p)
    }
  }
  __syncthreads();
  if (warp == 0) DR_STAMP(0, 42);

  // ---- outputs: the lower triangle of L = Z diag(sqrt d) in place and the log-determinant.  Neither the strict upper
  // triangle of the block nor the part of Winv above the block diagonal is written: nothing reads the former (every
  // consumer of a diagonal block of L masks k <= i) and the caller provides a zero-initialised Winv.  (Storing the
  // panels earlier, behind S2(p), was measured slower: the scaling DMULs then queue behind the update's DMMAs.)
  if (FULL_INV) {
    const int c = (tid & 63) * 2, pc = c >> 5;   // 384 = 6 x 64: a thread keeps its column pair, rows r0, r0 + 6, ..
    const double2 sd = *reinterpret_cast<const double2*>(sm.sdv + c);
    const double* zcol = sm.Zs + (dr_zrow0(pc) - 32 * pc) * DR_S + (c & 31);
#pragma unroll 4
    for (int r = tid >> 6; r < 128; r += DR_THREADS / 64) {
      if (c <= r) {
        const double2 z = *reinterpret_cast<const double2*>(zcol + r * DR_S);
        double2 l;
        l.x = (c < r) ? z.x * sd.x : sd.x;
        l.y = (c + 1 < r) ? z.y * sd.y : ((c + 1 == r) ? sd.y : 0.0);
        *reinterpret_cast<double2*>(Ablk + (long long)r * lda + c) = l;
      }
    }
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
  if (warp == 0) DR_STAMP(0, 43);
}

}  // namespace gpb
