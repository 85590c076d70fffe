// libgpb: host orchestration and C ABI (include/gpb.h) of the B200-native GP backend.
//
// Data layout in HBM (all row-major fp64, leading dimensions padded to the 128 block size):
//   Lbuf (Np+128) x Np : K -> L (lower).  Rows [Np, Np+128) carry y^T so the panel solves also produce
//                        z = L^-1 y.  The strict upper triangle is scratch (T of the triangular inverse); the
//                        lower triangle is finally overwritten by K^-1 when a gradient is requested.
//   Ubuf Np x Np       : U = L^-T (upper)            Wbuf Np x Np : W = L^-1 = U^T (lower)
//   Dinv Np x 128      : inverses of the 128x128 diagonal blocks of L
//   stashG, stashK     : (gradient evaluations) radial derivative factor and k_pure of every pair of the lower triangle,
//                        packed 128x128 blocks, written by the assembly and read once by the gradient contraction
// Padding rows/cols (index >= N) hold the identity, which every stage maps to the identity.
//
// Algorithms (all dense O(N^3) work is tile tasks for dgemm_tasks_kernel, see dgemm_tasks.cuh):
//   potrf : two-level right-looking blocked Cholesky; panel solve L_ik = A_ik W_kk^T uses the inverted diagonal
//           block, so it is a GEMM too.                                   (np.linalg.cholesky, gp_functions.py:143)
//   trtri : recursive  U12 = -U11 L21^T U22  evaluated level by level as two batched NT GEMMs.
//   syrk  : K^-1 = U U^T (lower).                                          (invert_cholesky, gp_functions.py:304-322)
//   predict variance : ||W k*||^2 as a tile GEMM with a column sum-of-squares epilogue.
//   predict covariance: K** - Vt Vt^T, Vt = K*^T W^T (two tile GEMMs).
//
// Scheduling (build_potrf_plan / run_plan): two panels of look-ahead (the panel chain and the small update of the panel
// after next on high-priority streams), large trailing updates split over two streams by block-column parity, leading
// sub-trees of the triangular inverse started on a low-priority stream while the Cholesky is still running.  tests/test_plans.py replays and race-checks these lists on the host.
//
// Active learning (acquisition.cuh): acquisition losses + masked arg-max on device-resident predictions, and the
// O(N^2) extension of W, U, alpha, z, log|K| by one observed point (gpb_append_train).
#include "../../include/gpb.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "acquisition.cuh"
#include "assemble.cuh"
#include "dgemm_tasks.cuh"
#include "eig.cuh"
#include "potrf_diag.cuh"
#include "potrf_diagr.cuh"
#include "trsm_panel.cuh"

using namespace gpb;

namespace {

enum MatId { M_L = 0, M_U, M_W, M_DINV, M_KST, M_PART, M_DBG_A, M_DBG_B, M_DBG_C, M_NONE };
enum GemmCfg { CFG_128x128 = 0, CFG_128x64 = 1, CFG_64x128 = 2, CFG_32x128 = 3, CFG_32x64 = 4, NUM_CFG };

struct Launch {
  int type;  // 0 = diag block, 1 = tile GEMM
  int k;     // diag: block index
  int cfg, matA, matB, matC, matCt;
  double alpha, beta;
  size_t task_off;
  int ntasks;
  int epi;
  int stream = 0;    // 0 = main stream, 1 = panel (look-ahead) stream, 2 = second trailing-update stream,
                     // 3 = look-ahead update stream (T2a: the columns of the panel after next)
  int wait_ev = -1;  // event to wait for before the launch (index into the handle's event pool)
  int wait_ev2 = -1; // optional second event
  int wait_ev3 = -1; // optional third event
  int rec_ev = -1;   // event recorded after the launch
  int pdl = 0;       // launch with programmatic stream serialization (overlaps the launch with the previous kernel)
  int trsm = 0;      // diag block: factor-only variant (32x32 diagonal inverses);  panel solve: blocked triangular solve
};

struct Plan {
  std::vector<Launch> launches;
  std::vector<GemmTask> tasks;
  int num_events = 0;
  GemmTask* d_tasks = nullptr;
  void clear() {
    launches.clear();
    tasks.clear();
    num_events = 0;
    if (d_tasks) cudaFree(d_tasks);
    d_tasks = nullptr;
  }
};

// Frees the device task list of a function-local plan on every exit path.
struct PlanGuard {
  Plan& p;
  explicit PlanGuard(Plan& pl) : p(pl) {}
  ~PlanGuard() { p.clear(); }
};

struct MatDesc {
  double* ptr = nullptr;
  long long rows = 0, cols = 0, ld = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Workspace {
  long long N = 0, Np = 0, Nb = 0;
  MatDesc mat[M_NONE];
  double *Lbuf = nullptr, *Ubuf = nullptr, *Wbuf = nullptr, *Dinv = nullptr;
  double *logdet = nullptr, *alpha = nullptr;
  double *stashG = nullptr, *stashK = nullptr;   // radial-factor stashes of the assembly (assemble.cuh::stash_offset)
  size_t stash_cap = 0, stashK_cap = 0;          // doubles per array
  Plan potrf, trtri, syrk;
  // Early triangular inverse: the sub-trees of the trtri recursion that only need the leading blocks of L run on a
  // low-priority stream while the Cholesky is still in its chain-bound second half (see build_trtri_plan).
  static constexpr int MAX_PHASE = 4;
  int n_phase = 0;
  Plan trtri_phase[MAX_PHASE];   // phase i: nodes inside [0, phase_cut[i]) that are not in an earlier phase
  int phase_cut[MAX_PHASE] = {}; // in 128-blocks
  int phase_ev[MAX_PHASE] = {};  // potrf-plan event after which L[0:cut, 0:cut] is final
  Plan trtri_late;               // the remaining nodes
  std::map<std::pair<int, int>, CUtensorMap> maps;
  int dinv_partial_from = 1 << 30;   // first diagonal block of which Dinv only holds the 32x32 diagonal blocks of the inverse
};

}  // namespace

struct gpb_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // high-priority panel stream for the Cholesky look-ahead
  cudaStream_t stream3 = nullptr;   // low-priority stream of the early triangular inverse
  cudaStream_t stream4 = nullptr;   // second trailing-update stream (same priority as the main stream)
  cudaStream_t stream5 = nullptr;   // high-priority stream of the look-ahead updates T2a (see build_potrf_plan)
  cudaEvent_t ev_early = nullptr;   // recorded on stream3 after the last early phase
  cudaStream_t caller = nullptr;    // gpb_set_stream: producer stream of the caller's device buffers (nullptr = legacy default)
  cudaEvent_t ev_caller = nullptr;
  std::vector<cudaEvent_t> events;
  std::string err;
  EncodeTiledFn encode = nullptr;
  long long launches = 0;

  // training problem
  Workspace ws;       // bound to the training set
  Workspace tmp;      // scratch problem for the function-level twins (cholesky / invert / solve)
  int D = 0, DP = 0, n_out = 0;
  double *dX = nullptr, *dy = nullptr, *dXs = nullptr, *dXsT = nullptr;
  double* dscal = nullptr;   // device scalars: [0]=nll, [1..] gradient sums
  double* dpart = nullptr;   // gradient partials
  size_t dpart_cap = 0;
  int* dinfo = nullptr;
  KernelHyp hyp{};
  bool factor_ready = false;

  // predict chunk state
  long long Mc = 0;
  double *dXc = nullptr, *dKsT = nullptr, *dVpart = nullptr, *dmean = nullptr, *dvar = nullptr;
  Plan pred;
  long long pred_Nb = 0, pred_Mc = 0;

  // generic staging
  double* dstage = nullptr;
  size_t dstage_cap = 0;
  double* dstage2 = nullptr;
  size_t dstage2_cap = 0;
  double* dkm = nullptr;      // scaled inputs of the dense kernel-matrix kernel
  size_t dkm_cap = 0;
  double* dbbq = nullptr;     // marginal-variance scratch
  size_t dbbq_cap = 0;
  double* dapp = nullptr;     // append-point scratch: k row block, l, U l, scalars
  size_t dapp_cap = 0;
  double* dacq = nullptr;     // acquisition scratch: staged host inputs, block partials, statistics
  size_t dacq_cap = 0;

  // truncated eigen-decomposition (eig.cuh): Lanczos basis, projected problem, Ritz pairs
  struct Eig {
    double* K = nullptr;        // dense symmetric matrix (n x n, ld = n)
    size_t K_cap = 0;
    double* buf = nullptr;      // one allocation: Vt, w, c1, c2, alpha, beta, d, Zt, Qt, wsel, qy, t, avec
    size_t buf_cap = 0;
    double *Vt = nullptr, *w = nullptr, *c1 = nullptr, *c2 = nullptr, *alpha = nullptr, *beta = nullptr, *d = nullptr,
           *Zt = nullptr, *Qt = nullptr, *wsel = nullptr, *qy = nullptr, *t = nullptr, *avec = nullptr;
    int* ibuf = nullptr;        // sel[mcap], flag[2]
    int n = 0, mcap = 0, m = 0, k = 0, restarts = 0;
    double grad_jitter = 0.0;   // diagonal shift under which invert(Ky) of the last gpb_nll_eig gradient factorised
    bool matrix_ready = false;  // K holds the matrix described by key
    bool from_train = false;    // K = Ky(theta) of the resident training set (gpb_eigsh) rather than an uploaded matrix
    std::vector<double> key;    // kind, n_ell, theta...
    std::vector<double> w_host; // the k resident eigenvalues, ascending
    // the last diagonalisation of the projected problem (size solved_m): Ritz values, last components of their vectors,
    // descending order, so that a request for more pairs at the same size needs no second QL run
    std::vector<double> ritz, zlast;
    std::vector<int> order, sel;
    int solved_m = 0, vectors_k = 0, vectors_m = 0;
    double beta_last = 0.0, ritz_scale = 0.0;
  } eig;

  cudaEvent_t ev0[GPB_NUM_STAGES], ev1[GPB_NUM_STAGES];
  cudaEvent_t evx[3];   // scratch timing events (gradient tail, predict chunk phases, debug GEMM)
  bool ev_used[GPB_NUM_STAGES];
  double stage_ms[GPB_NUM_STAGES];
};

namespace {

typedef gpb_handle_s H;

int fail(H* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(h, -2, std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + std::to_string(__LINE__)); \
  } while (0)

long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// Order the handle's stream after everything the caller has queued on its own stream so far (gpb_set_stream): device
// buffers produced there are complete before any kernel of this call reads them.
int order_after_caller(H* h) {
  if (!h->caller) return 0;
  CK(cudaEventRecord(h->ev_caller, h->caller));
  CK(cudaStreamWaitEvent(h->stream, h->ev_caller, 0));
  return 0;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int pick_dp(int D) {
  const int opts[] = {4, 8, 12, 16, 24, 32};
  for (int o : opts)
    if (D <= o) return o;
  return -1;
}

int ensure_cap(H* h, double** p, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) CK(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  CK(cudaMalloc(p, need * sizeof(double)));
  *cap = need;
  return 0;
}

// ------------------------------------------------------------------------------------------------ tensor maps
int get_map(H* h, Workspace& w, int mat, int box_rows, CUtensorMap* out) {
  auto key = std::make_pair(mat, box_rows);
  auto it = w.maps.find(key);
  if (it != w.maps.end()) {
    *out = it->second;
    return 0;
  }
  const MatDesc& m = w.mat[mat];
  if (!m.ptr) return fail(h, -3, "tensor map requested for an unallocated matrix");
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows};
  cuuint64_t gstr[1] = {(cuuint64_t)m.ld * sizeof(double)};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, m.ptr, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, -2, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  w.maps[key] = tm;
  *out = tm;
  return 0;
}

void set_mat(Workspace& w, int id, double* p, long long rows, long long cols, long long ld) {
  w.mat[id].ptr = p;
  w.mat[id].rows = rows;
  w.mat[id].cols = cols;
  w.mat[id].ld = ld;
  for (auto it = w.maps.begin(); it != w.maps.end();)
    it = (it->first.first == id) ? w.maps.erase(it) : std::next(it);
}

// ------------------------------------------------------------------------------------------------ GEMM launch
template <int BM, int BN, int WM, int WN, int ST, int EPI, int MINB>
int launch_cfg(H* h, cudaStream_t stream, const CUtensorMap& tA, const CUtensorMap& tB, const GemmTask* tasks,
               int ntasks, const GemmEpilogue& ep, bool pdl) {
  auto kern = dgemm_tasks_kernel<BM, BN, WM, WN, ST, EPI, MINB>;
  constexpr int smem = GemmSmem<BM, BN, ST>::BYTES;
  static bool configured[64] = {};   // the attribute is per device; one process may hold handles on several GPUs
  if (!configured[h->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[h->device & 63] = true;
  }
  constexpr int threads = (BM / WM) * (BN / WN) * 32 + 32;
  if (pdl) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)ntasks);
    lc.blockDim = dim3((unsigned)threads);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    CK(cudaLaunchKernelEx(&lc, kern, tA, tB, tasks, ep));
  } else {
    kern<<<ntasks, threads, smem, stream>>>(tA, tB, tasks, ep);
  }
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

void cfg_tile(int cfg, int* bm, int* bn) {
  switch (cfg) {
    case CFG_128x128: *bm = 128; *bn = 128; break;
    case CFG_128x64: *bm = 128; *bn = 64; break;
    case CFG_32x128: *bm = 32; *bn = 128; break;
    case CFG_32x64: *bm = 32; *bn = 64; break;
    default: *bm = 64; *bn = 128; break;
  }
}

int launch_gemm(H* h, Workspace& w, int cfg, int epi, int matA, int matB, const GemmTask* d_tasks, int ntasks,
                const GemmEpilogue& ep, cudaStream_t stream = nullptr, bool pdl = false) {
  if (!stream) stream = h->stream;
  if (ntasks <= 0) return 0;
  int bm, bn;
  cfg_tile(cfg, &bm, &bn);
  CUtensorMap tA, tB;
  int rc;
  if ((rc = get_map(h, w, matA, bm, &tA))) return rc;
  if ((rc = get_map(h, w, matB, bn, &tB))) return rc;
  if (epi == EPI_STORE) {
    switch (cfg) {
      case CFG_128x128: return launch_cfg<128, 128, 64, 32, 4, EPI_STORE, 1>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      case CFG_128x64: return launch_cfg<128, 64, 32, 32, 4, EPI_STORE, 2>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      case CFG_32x128: return launch_cfg<32, 128, 16, 32, 4, EPI_STORE, 2>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      case CFG_32x64: return launch_cfg<32, 64, 16, 16, 4, EPI_STORE, 2>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      default: return launch_cfg<64, 128, 32, 32, 4, EPI_STORE, 2>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
    }
  } else {
    switch (cfg) {
      case CFG_128x128: return launch_cfg<128, 128, 64, 32, 4, EPI_COLSQ, 1>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      case CFG_128x64: return launch_cfg<128, 64, 32, 32, 4, EPI_COLSQ, 2>(h, stream, tA, tB, d_tasks, ntasks, ep, pdl);
      default: return fail(h, -1, "no column-norm epilogue for this tile configuration");
    }
  }
}

// ------------------------------------------------------------------------------------------------ plans
int g_cfg_update = CFG_128x64;  // trailing updates / trtri / syrk / predict   (env GPB_CFG_UPDATE)
int g_outer_blocks = 0;         // outer panel width of the Cholesky in 128-blocks; 0 = by size (env GPB_OUTER_BLOCKS)
int g_diag4 = 2;                // diagonal-block kernel: 2 = blocked DMMA kernel (potrf_diagr.cuh), 1 = rank-4 register
                                // sweep, 0 = rank-1 register sweep (env GPB_DIAG4)
int g_trsm = 1;                 // panel solve as a blocked triangular solve; the diagonal-block kernel then skips the 128x128
                                // inverse, which complete_inverse_kernel supplies off the chain (env GPB_TRSM; needs GPB_DIAG4=2)
int g_trsm_below = 40;          // ... for the block columns with at most this many block rows left (env GPB_TRSM_BELOW): while the
                                // panel is taller the chain is hidden behind the trailing update and the 64-row GEMM solve with
                                // the full inverse has the better throughput (N=16384: 130.8 vs 131.2 ms per evaluation)
int g_early_min_blocks = 16;    // smallest matrix (in 128-blocks) whose triangular inverse starts during the Cholesky (env GPB_EARLY_MIN_BLOCKS)
int g_super_min_blocks = 48;    // smallest matrix for the super-tile order of K^-1 = U U^T (below, the plain order is 5 % faster)
int g_pdl = 1;                  // programmatic dependent launch along the panel chain (env GPB_PDL)
int g_tail_blocks = 16;         // the last this-many block columns use single-block panels (env GPB_TAIL_BLOCKS)
int g_wide_above = 64;           // block columns left above which the panels are four blocks wide (env GPB_WIDE_ABOVE)
int g_lookahead = 2;            // panels of look-ahead: the panel chain runs this many panels ahead of the bulk trailing update
                                // (env GPB_LOOKAHEAD; 0 = one stream in dependency order)
int g_split_t2 = 1;             // trailing updates as two interleaved halves on two streams (env GPB_SPLIT_T2)
int g_early_trtri = 1;          // start the triangular inverse of the leading blocks during the Cholesky (env GPB_EARLY_TRTRI)
int g_early_cfg = CFG_32x128;   // tiles of the early phases: short CTAs delay the trailing updates least; -1 = as late (env GPB_EARLY_CFG)
int g_early_mask = 0xF;         // which of the cut points {1/4, 1/2, 3/4, 7/8} are used (env GPB_EARLY_MASK)

// Small launches are latency-bound (one CTA takes ~15 us for a 128x64x128 tile on its own SM): cut the tiles finer
// so the same work spreads over more SMs.  `coarse` is the configuration used for large launches.
int g_super = 12;                // super-tile edge of the K^-1 = U U^T launch in 128-blocks, ~sqrt(148) (env GPB_SUPER; 0 = off)
int g_pred_super_m = 12;         // super-tile of the predict column-norm GEMM: row tiles of W x candidate tiles (env GPB_PRED_SUPER_M / _C;
int g_pred_super_c = 12;         // M = 0: plain order).  12 x 12 blocks = 288 128x64 CTAs ~ the 296 resident ones
int g_eig_small = 1;             // one-CTA Lanczos kernel for matrices up to EIG_SMALL_N rows (env GPB_EIG_SMALL; 0 = always the multi-kernel path)
int g_super_trtri = 0;           // the same for the triangular-inverse launches (env GPB_SUPER_TRTRI): measured +7 % time at
                                 // N=16384 (the heaviest-first order matters more there than the L2 hit rate), so off
int g_stash = 1;                 // the assembly leaves k_pure and the derivative factor of every pair for the gradient
                                 // contraction (env GPB_STASH; 0 = the contraction recomputes them)
int g_fine_below = 300;          // launches with fewer 128-blocks than this use the fine tiles (env GPB_FINE_BELOW)
int cfg_for_blocks(int coarse, int nblocks) {
  if (nblocks >= g_fine_below) return coarse;
  return (coarse == CFG_64x128) ? CFG_32x128 : CFG_32x64;
}

struct TaskSink {
  Plan& p;
  size_t begin;
  std::vector<int> keys;   // optional sort key per task (finish(..., sort_heavy) orders by key, else by nk)
  int key = -1;            // key given to the tasks emitted next (-1: none)
  explicit TaskSink(Plan& pl) : p(pl), begin(pl.tasks.size()) {}
  // emit the tile(s) covering the 128x128 block (c_row, c_col) for the given config
  void block(int cfg, int a_row, int a_k, int b_row, int b_k, int nk, int c_row, int c_col, int aux = 0) {
    const size_t before = p.tasks.size();
    block_(cfg, a_row, a_k, b_row, b_k, nk, c_row, c_col, aux);
    if (key >= 0) keys.resize(keys.size() + (p.tasks.size() - before), key);
  }
  void block_(int cfg, int a_row, int a_k, int b_row, int b_k, int nk, int c_row, int c_col, int aux) {
    if (cfg == CFG_128x128) {
      p.tasks.push_back({a_row, a_k, b_row, b_k, nk, c_row, c_col, aux});
    } else if (cfg == CFG_128x64) {
      for (int hlf = 0; hlf < 2; ++hlf)
        p.tasks.push_back({a_row, a_k, b_row + 64 * hlf, b_k, nk, c_row, c_col + 64 * hlf, aux});
    } else if (cfg == CFG_32x128) {
      for (int q = 0; q < 4; ++q) p.tasks.push_back({a_row + 32 * q, a_k, b_row, b_k, nk, c_row + 32 * q, c_col, aux});
    } else if (cfg == CFG_32x64) {
      for (int q = 0; q < 4; ++q)
        for (int hlf = 0; hlf < 2; ++hlf)
          p.tasks.push_back({a_row + 32 * q, a_k, b_row + 64 * hlf, b_k, nk, c_row + 32 * q, c_col + 64 * hlf, aux});
    } else {
      for (int hlf = 0; hlf < 2; ++hlf)
        p.tasks.push_back({a_row + 64 * hlf, a_k, b_row, b_k, nk, c_row + 64 * hlf, c_col, aux});
    }
  }
  void finish(int cfg, int epi, int matA, int matB, int matC, int matCt, double alpha, double beta, bool sort_heavy) {
    const int n = (int)(p.tasks.size() - begin);
    if (n == 0) return;
    if (sort_heavy && (int)keys.size() == n) {
      // stable sort by group key (descending): super-tiles stay contiguous, heavy groups first
      std::vector<int> order(n);
      for (int t = 0; t < n; ++t) order[t] = t;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] > keys[b]; });
      std::vector<GemmTask> sorted(n);
      for (int t = 0; t < n; ++t) sorted[t] = p.tasks[begin + order[t]];
      std::copy(sorted.begin(), sorted.end(), p.tasks.begin() + begin);
    } else if (sort_heavy) {
      std::stable_sort(p.tasks.begin() + begin, p.tasks.end(),
                       [](const GemmTask& a, const GemmTask& b) { return a.nk > b.nk; });
    }
    p.launches.push_back({1, 0, cfg, matA, matB, matC, matCt, alpha, beta, begin, n, epi});
    last = (int)p.launches.size() - 1;
  }
  int last = -1;  // index of the launch emitted by finish(), -1 if the task list was empty
};

// Two-level right-looking Cholesky with one panel of look-ahead.
//   P(J)     : factor outer panel J (per 128-column: diagonal block, panel solve, update of the panel's own
//              remaining columns)                                                        -- panel stream
//   T1(J)    : update the columns of panel J+1 with panel J                              -- panel stream
//   T2(J)    : update every column right of panel J+1 with panel J                       -- main stream
// P(J+1) only depends on T1(J), so it runs concurrently with the large T2(J).  Events: evP[J] after P(J),
// evM[J] after T2(J);  T1(J) waits evM[J-1] (both touch panel J+1's columns), T2(J) waits evP[J].
// With g_lookahead == 0 everything is issued on the main stream in dependency order (no events).
void build_potrf_plan(Workspace& w) {
  Plan& p = w.potrf;
  p.clear();
  // Outer panel width: wide panels (K = 512 trailing updates) pay off when the trailing matrix is large; for small
  // matrices the factorisation is bound by the per-column-block chain and narrow panels shorten it
  // (measured: N=4096 potrf 3.15 / 3.02 / 2.84 ms with 4 / 2 / 1 blocks, N=16384 48.2 / 48.4 / 50.8 ms).
  const int Nb = (int)w.Nb, cfg = g_cfg_update;
  const int aug = Nb;  // block-row index of the y^T rows
  // Panel boundaries: wide panels while the trailing matrix is large, single-block panels for the last
  // g_tail_blocks block columns, where the factorisation is chain-bound anyway.
  std::vector<int> pb{0};
  while (pb.back() < Nb) {
    const int left = Nb - pb.back();
    int wdt = g_outer_blocks > 0 ? g_outer_blocks : (left > g_wide_above ? 4 : (left > g_tail_blocks ? 2 : 1));
    pb.push_back(std::min(Nb, pb.back() + wdt));
  }
  const int nJ = (int)pb.size() - 1;
  const bool la = g_lookahead != 0;
  const int ps = la ? 1 : 0;           // stream of the panel work
  auto evP = [&](int J) { return J; };
  auto evM = [&](int J) { return nJ + J; };
  auto evM2 = [&](int J) { return 2 * nJ + J; };   // second half of a split trailing update
  auto evA = [&](int J) { return 3 * nJ + J; };    // T2a(J): panel J+2's columns updated with panel J
  p.num_events = la ? 4 * nJ : 0;
  std::vector<char> split(nJ, 0);
  // Cut points of the early triangular inverse: L[0:c, 0:c] is final once the panel that contains block column
  // c - 1 has been factored and solved (event evP of that panel).
  w.n_phase = 0;
  if (la && g_early_trtri && Nb >= g_early_min_blocks) {   // below ~6k rows the whole factorisation is chain-bound: nothing to gain
    const int cuts[Workspace::MAX_PHASE] = {Nb / 4, Nb / 2, 3 * Nb / 4, 7 * Nb / 8};
    for (int ci = 0; ci < Workspace::MAX_PHASE; ++ci) {
      if (!((g_early_mask >> ci) & 1)) continue;
      const int c = cuts[ci];
      int J = 0;
      while (pb[J + 1] < c) ++J;
      w.phase_cut[w.n_phase] = c;
      w.phase_ev[w.n_phase] = evP(J);
      ++w.n_phase;
    }
  }

  // Programmatic dependent launch pays off only while the chain is the bottleneck (measured: potrf N=4096
  // 2.50 -> 2.39 ms, but N=8192 8.30 -> 8.60 ms: early-resident dependents then take SMs from the updates).
  const int pdl = (Nb <= 40) ? 1 : 0;
  auto emit_panel = [&](int J) {
    const int J0 = pb[J], Je = pb[J + 1];
    for (int k = J0; k < Je; ++k) {
      const bool trsm = g_trsm && g_diag4 == 2 && (Nb - k) <= g_trsm_below;
      Launch dg{0, k, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      dg.stream = ps;
      dg.pdl = pdl;
      dg.trsm = trsm;
      p.launches.push_back(dg);
      {  // panel solve: L_ik = A_ik W_kk^T, in place (64-row tiles own all the columns they read)
        TaskSink s(p);
        const int tc = trsm ? CFG_32x128 : cfg_for_blocks(CFG_64x128, aug - k);
        for (int i = k + 1; i <= aug; ++i) s.block(tc, i * 128, k * 128, k * 128, 0, 8, i * 128, k * 128);
        s.finish(tc, EPI_STORE, M_L, M_DINV, M_L, M_NONE, 1.0, 0.0, false);
        if (s.last >= 0) {
          p.launches[s.last].stream = ps;
          p.launches[s.last].pdl = pdl;
          p.launches[s.last].trsm = trsm;
        }
      }
      {  // update the remaining columns of the current outer panel with block column k
        TaskSink s(p);
        int nblk = 0;
        for (int j = k + 1; j < Je; ++j) nblk += aug - j + 1;
        const int uc = cfg_for_blocks(cfg, nblk);
        for (int j = k + 1; j < Je; ++j)
          for (int i = j; i <= aug; ++i) s.block(uc, i * 128, k * 128, j * 128, k * 128, 8, i * 128, j * 128);
        s.finish(uc, EPI_STORE, M_L, M_L, M_L, M_NONE, -1.0, 1.0, false);
        if (s.last >= 0) {
          p.launches[s.last].stream = ps;
          p.launches[s.last].pdl = pdl;
        }
      }
    }
    if (la) p.launches.back().rec_ev = evP(J);
  };
  // columns [jb0, jb1) (in 128-blocks) updated with panel J
  // parity: -1 = every block column, 0 / 1 = only the even / odd ones
  auto emit_update = [&](int J, int jb0, int jb1, int stream, int wait_ev, int rec_ev, int parity = -1, int wait_ev2 = -1,
                         int wait_ev3 = -1) {
    const int J0 = pb[J], Je = pb[J + 1];
    TaskSink s(p);
    int nblk = 0;
    for (int j = jb0; j < jb1; ++j)
      if (parity < 0 || (j & 1) == parity) nblk += aug - j + 1;
    const int uc = cfg_for_blocks(cfg, nblk);
    for (int j = jb0; j < jb1; ++j)
      if (parity < 0 || (j & 1) == parity)
        for (int i = j; i <= aug; ++i) s.block(uc, i * 128, J0 * 128, j * 128, J0 * 128, 8 * (Je - J0), i * 128, j * 128);
    s.finish(uc, EPI_STORE, M_L, M_L, M_L, M_NONE, -1.0, 1.0, false);
    if (s.last >= 0) {
      p.launches[s.last].stream = stream;
      p.launches[s.last].wait_ev = wait_ev;
      p.launches[s.last].wait_ev2 = wait_ev2;
      p.launches[s.last].wait_ev3 = wait_ev3;
      p.launches[s.last].rec_ev = rec_ev;
    }
    return s.last >= 0;
  };

  emit_panel(0);
  for (int J = 0; J + 1 < nJ; ++J) {
    const int next0 = pb[J + 1], next1 = pb[J + 2];
    if (la) {
      const bool has_t2 = next1 < Nb;
      // Two panels of look-ahead (g_lookahead == 2): the columns of panel J+2 are updated by a small launch of their own,
      // T2a(J), on a high-priority stream (same-priority grids are dispatched first come first served, so behind a bulk
      // update it would wait for the whole of it).  T1(J+1) then only waits for T2a(J) (event evA) instead of the whole
      // trailing update: the panel chain P(J+1) -> T1(J+1) -> P(J+2) runs up to two panels ahead of the bulk, and the
      // main stream always has the next bulk update queued.
      const int deep = (g_lookahead >= 2 && has_t2) ? 1 : 0;
      const int next2 = deep ? pb[std::min(J + 3, nJ)] : next1;
      const int cross = (J > 0 && split[J - 1]) ? evM2(J - 1) : -1;   // the previous bulk update of the other parity
      // T2(J) first so that the big launch is queued early on the main stream.  While the trailing matrix is large it
      // goes out as two halves (even / odd block columns) on two streams: the drain of one half is covered by the
      // other, and the column -> stream map is static, so each half only depends on its own predecessor.
      if (has_t2) {
        if (deep) emit_update(J, next1, next2, 3, evP(J), evA(J), -1, J > 0 ? evM(J - 1) : -1, cross);
        const bool bulk = next2 < Nb;
        split[J] = bulk && g_split_t2 && (Nb - next2) >= 16;
        if (split[J]) {
          emit_update(J, next2, Nb, 0, evP(J), evM(J), 0);
          emit_update(J, next2, Nb, 2, evP(J), evM2(J), 1);
        } else if (bulk) {
          // an unsplit update touches both parities: it follows the last split one on the other stream as well
          emit_update(J, next2, Nb, 0, evP(J), evM(J), -1, cross);
        }
      }
      // T1(J): with two panels of look-ahead its columns were last touched by T2a(J-1), otherwise by T2(J-1)
      if (g_lookahead >= 2 && J > 0) emit_update(J, next0, next1, 1, evA(J - 1), -1);
      else emit_update(J, next0, next1, 1, J > 0 ? evM(J - 1) : -1, -1, -1, (J > 0 && split[J - 1]) ? evM2(J - 1) : -1);
      emit_panel(J + 1);
    } else {
      emit_update(J, next0, Nb, 0, -1, -1);
      emit_panel(J + 1);
    }
  }
}

struct Node {
  int a, mid, b, height;
};

int build_nodes(int a, int b, std::vector<Node>& out) {
  if (b - a <= 1) return 0;
  const int mid = a + (b - a + 1) / 2;
  const int hl = build_nodes(a, mid, out), hr = build_nodes(mid, b, out);
  const int hgt = 1 + std::max(hl, hr);
  out.push_back({a, mid, b, hgt});
  return hgt;
}

// Emit the nodes selected by `pick` level by level into plan p.
template <class Pick>
void emit_trtri_nodes(Plan& p, const std::vector<Node>& nodes, int top, int cfg, Pick pick, int force_cfg = -1) {
  p.clear();
  for (int hgt = 1; hgt <= top; ++hgt) {
    int nblk = 0;
    for (const Node& nd : nodes)
      if (nd.height == hgt && pick(nd)) nblk += (nd.mid - nd.a) * (nd.b - nd.mid);
    if (nblk == 0) continue;
    const int hc = force_cfg >= 0 ? force_cfg : cfg_for_blocks(cfg, nblk);
    // Tiles of a node are emitted super-tile by super-tile (see build_syrk_plan) and the launch is ordered by the
    // heaviest tile of each super-tile, so that resident CTAs share operand panels and heavy groups still go first.
    const int S = g_super_trtri > 0 ? g_super_trtri : (1 << 20);
    {  // T = U11 L21^T  -> scratch in the strict upper triangle of Lbuf
      TaskSink s(p);
      for (const Node& nd : nodes)
        if (nd.height == hgt && pick(nd))
          for (int M0 = nd.a; M0 < nd.mid; M0 += S)
            for (int N0 = nd.mid; N0 < nd.b; N0 += S) {
              s.key = g_super_trtri > 0 ? 8 * (nd.mid - M0) : -1;
              for (int m = M0; m < std::min(nd.mid, M0 + S); ++m)
                for (int n = N0; n < std::min(nd.b, N0 + S); ++n)
                  s.block(hc, m * 128, m * 128, n * 128, m * 128, 8 * (nd.mid - m), m * 128, n * 128);
            }
      s.finish(hc, EPI_STORE, M_U, M_L, M_L, M_NONE, 1.0, 0.0, true);
    }
    {  // U12 = -T W22^T, mirrored into W21
      TaskSink s(p);
      for (const Node& nd : nodes)
        if (nd.height == hgt && pick(nd))
          for (int M0 = nd.a; M0 < nd.mid; M0 += S)
            for (int N0 = nd.mid; N0 < nd.b; N0 += S) {
              s.key = g_super_trtri > 0 ? 8 * (std::min(nd.b, N0 + S) - nd.mid) : -1;
              for (int m = M0; m < std::min(nd.mid, M0 + S); ++m)
                for (int n = N0; n < std::min(nd.b, N0 + S); ++n)
                  s.block(hc, m * 128, nd.mid * 128, n * 128, nd.mid * 128, 8 * (n - nd.mid + 1), m * 128, n * 128);
            }
      s.finish(hc, EPI_STORE, M_L, M_W, M_U, M_W, -1.0, 0.0, true);
    }
  }
}

// w.trtri: every node (used when L is given, e.g. gpb_invert_cholesky).  w.trtri_phase[i] + w.trtri_late: the same
// nodes split by the right end of their block range, so that phase i can run as soon as the Cholesky has finished
// the block columns left of phase_cut[i] (a node [a, b) only reads L[a:b, a:b] and the U / W of its children).
void build_trtri_plan(Workspace& w) {
  const int Nb = (int)w.Nb, cfg = g_cfg_update;
  std::vector<Node> nodes;
  const int top = build_nodes(0, Nb, nodes);
  emit_trtri_nodes(w.trtri, nodes, top, cfg, [](const Node&) { return true; });
  auto phase_of = [&](const Node& nd) {
    for (int i = 0; i < w.n_phase; ++i)
      if (nd.b <= w.phase_cut[i]) return i;
    return w.n_phase;
  };
  for (int i = 0; i < w.n_phase; ++i)
    emit_trtri_nodes(w.trtri_phase[i], nodes, top, cfg, [&](const Node& nd) { return phase_of(nd) == i; }, g_early_cfg);
  emit_trtri_nodes(w.trtri_late, nodes, top, cfg, [&](const Node& nd) { return phase_of(nd) == w.n_phase; });
}

// Super-tile rasterisation (L2 reuse).  A tile (i, j) of an NT product streams the row panels i and j of its operands
// through the k range; CTAs are dispatched in task order, so the ~296 resident CTAs of a launch should form a compact
// g_super x g_super block of tiles: they then share 2 g_super row panels instead of reading one or two each, and
// because they walk k in step the shared slabs are still in L2 when the next CTA asks for them.
void build_syrk_plan(Workspace& w) {
  Plan& p = w.syrk;
  p.clear();
  const int Nb = (int)w.Nb, cfg = cfg_for_blocks(g_cfg_update, (int)(w.Nb * (w.Nb + 1) / 2));
  TaskSink s(p);
  auto tile = [&](int i, int j) { s.block(cfg, i * 128, i * 128, j * 128, i * 128, 8 * (Nb - i), i * 128, j * 128); };
  if (g_super <= 0 || Nb < g_super_min_blocks) {
    for (int i = 0; i < Nb; ++i)
      for (int j = 0; j <= i; ++j) tile(i, j);
    s.finish(cfg, EPI_STORE, M_U, M_U, M_L, M_NONE, 1.0, 0.0, true);
    return;
  }
  // rows ascending = longest k first; the super-tiles of one block row run left to right
  for (int I = 0; I < Nb; I += g_super)
    for (int J = 0; J <= I; J += g_super)
      for (int i = I; i < std::min(Nb, I + g_super); ++i)
        for (int j = J; j < std::min(J + g_super, i + 1); ++j) tile(i, j);
  s.finish(cfg, EPI_STORE, M_U, M_U, M_L, M_NONE, 1.0, 0.0, false);
}

int upload_plan(H* h, Plan& p) {
  if (p.d_tasks) {
    cudaFree(p.d_tasks);
    p.d_tasks = nullptr;
  }
  if (p.tasks.empty()) return 0;
  CK(cudaMalloc(&p.d_tasks, p.tasks.size() * sizeof(GemmTask)));
  CK(cudaMemcpyAsync(p.d_tasks, p.tasks.data(), p.tasks.size() * sizeof(GemmTask), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int run_plan(H* h, Workspace& w, Plan& p, bool factor_diag, cudaStream_t only_stream = nullptr) {
  const bool multi = p.num_events > 0 && !only_stream;
  if (factor_diag) w.dinv_partial_from = 1 << 30;   // set again below by the factor-only diagonal blocks of this plan
  if (multi) {
    while ((int)h->events.size() < p.num_events + 4) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->events.push_back(e);
    }
    // the panel stream starts after everything already queued on the main stream
    CK(cudaEventRecord(h->events[p.num_events], h->stream));
    CK(cudaStreamWaitEvent(h->stream2, h->events[p.num_events], 0));
    CK(cudaStreamWaitEvent(h->stream4, h->events[p.num_events], 0));
    CK(cudaStreamWaitEvent(h->stream5, h->events[p.num_events], 0));
  }
  for (const Launch& L : p.launches) {
    cudaStream_t st = only_stream ? only_stream
                                  : ((multi && L.stream == 1)   ? h->stream2
                                     : (multi && L.stream == 2) ? h->stream4
                                     : (multi && L.stream == 3) ? h->stream5
                                                                : h->stream);
    if (multi && L.wait_ev >= 0) CK(cudaStreamWaitEvent(st, h->events[L.wait_ev], 0));
    if (multi && L.wait_ev2 >= 0) CK(cudaStreamWaitEvent(st, h->events[L.wait_ev2], 0));
    if (multi && L.wait_ev3 >= 0) CK(cudaStreamWaitEvent(st, h->events[L.wait_ev3], 0));
    if (L.type == 0) {
      double* winv = w.Dinv + (long long)L.k * 128 * 128;
      if (factor_diag) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(1);
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = (L.pdl && g_pdl) ? 1 : 0;
        if (g_diag4 == 2) {
          lc.blockDim = dim3(DR_THREADS);
          lc.dynamicSmemBytes = DR_SMEM;
          if (L.trsm) w.dinv_partial_from = std::min(w.dinv_partial_from, L.k);
          if (L.trsm)
            CK(cudaLaunchKernelEx(&lc, potrf_diagr_kernel<false>, w.Lbuf, (long long)w.Np, L.k * 128, winv,
                                  w.logdet + L.k, h->dinfo, (int)w.N));
          else
            CK(cudaLaunchKernelEx(&lc, potrf_diagr_kernel<true>, w.Lbuf, (long long)w.Np, L.k * 128, winv,
                                  w.logdet + L.k, h->dinfo, (int)w.N));
        } else if (g_diag4 == 1) {
          lc.blockDim = dim3(DIAG_THREADS);
          lc.dynamicSmemBytes = DIAG4_SMEM;
          CK(cudaLaunchKernelEx(&lc, potrf_diag4_kernel, w.Lbuf, (long long)w.Np, L.k * 128, winv, w.logdet + L.k,
                                h->dinfo, (int)w.N));
        } else {
          lc.blockDim = dim3(DIAG_THREADS);
          lc.dynamicSmemBytes = DIAG_SMEM;
          CK(cudaLaunchKernelEx(&lc, potrf_diag_kernel<true>, w.Lbuf, (long long)w.Np, L.k * 128, winv, w.logdet + L.k,
                                h->dinfo, (int)w.N));
        }
      } else
        potrf_diag_kernel<false><<<1, DIAG_THREADS, DIAG_SMEM, st>>>(w.Lbuf, w.Np, L.k * 128, winv, w.logdet + L.k,
                                                                     h->dinfo, (int)w.N);
      h->launches++;
      CK(cudaGetLastError());
    } else if (factor_diag && L.trsm) {
      // panel solve L_ik = A_ik W_kk^T, executed as a blocked triangular solve on the same stripe tasks
      cudaLaunchConfig_t lc{};
      lc.gridDim = dim3((unsigned)L.ntasks);
      lc.blockDim = dim3(TP_THREADS);
      lc.dynamicSmemBytes = TP_SMEM;
      lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = at;
      lc.numAttrs = (L.pdl && g_pdl && !(multi && L.wait_ev >= 0)) ? 1 : 0;
      CK(cudaLaunchKernelEx(&lc, trsm_panel_kernel, w.Lbuf, (long long)w.Np, (const double*)w.Dinv,
                            (const GemmTask*)(p.d_tasks + L.task_off)));
      h->launches++;
      CK(cudaGetLastError());
    } else {
      GemmEpilogue ep;
      ep.C = w.mat[L.matC].ptr;
      ep.ldc = w.mat[L.matC].ld;
      ep.Ct = (L.matCt == M_NONE) ? nullptr : w.mat[L.matCt].ptr;
      ep.ldct = (L.matCt == M_NONE) ? 0 : w.mat[L.matCt].ld;
      ep.alpha = L.alpha;
      ep.beta = L.beta;
      int rc = launch_gemm(h, w, L.cfg, L.epi, L.matA, L.matB, p.d_tasks + L.task_off, L.ntasks, ep, st,
                           L.pdl && g_pdl && !(multi && L.wait_ev >= 0));
      if (rc) return rc;
    }
    if (multi && L.rec_ev >= 0) CK(cudaEventRecord(h->events[L.rec_ev], st));
  }
  if (multi) {
    // join: the main stream continues only after the panel stream has drained
    CK(cudaEventRecord(h->events[p.num_events + 1], h->stream2));
    CK(cudaStreamWaitEvent(h->stream, h->events[p.num_events + 1], 0));
    CK(cudaEventRecord(h->events[p.num_events + 2], h->stream4));
    CK(cudaStreamWaitEvent(h->stream, h->events[p.num_events + 2], 0));
    CK(cudaEventRecord(h->events[p.num_events + 3], h->stream5));
    CK(cudaStreamWaitEvent(h->stream, h->events[p.num_events + 3], 0));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ workspace
void free_ws(Workspace& w) {
  cudaFree(w.stashG);
  cudaFree(w.stashK);
  w.stashG = w.stashK = nullptr;
  w.stash_cap = w.stashK_cap = 0;
  cudaFree(w.Lbuf);
  cudaFree(w.Ubuf);
  cudaFree(w.Wbuf);
  cudaFree(w.Dinv);
  cudaFree(w.logdet);
  cudaFree(w.alpha);
  w.Lbuf = w.Ubuf = w.Wbuf = w.Dinv = w.logdet = w.alpha = nullptr;
  w.potrf.clear();
  w.trtri.clear();
  w.trtri_late.clear();
  for (Plan& p : w.trtri_phase) p.clear();
  w.n_phase = 0;
  w.syrk.clear();
  w.maps.clear();
  w.N = w.Np = w.Nb = 0;
}

int resize_ws(H* h, Workspace& w, long long N) {
  const long long Np = round_up(std::max<long long>(N, 1), 128);
  w.N = N;
  if (Np == w.Np) return 0;
  const long long keepN = N;
  free_ws(w);
  w.N = keepN;
  w.Np = Np;
  w.Nb = Np / 128;
  CK(cudaMalloc(&w.Lbuf, (size_t)(Np + 128) * Np * sizeof(double)));
  CK(cudaMalloc(&w.Dinv, (size_t)Np * 128 * sizeof(double)));
  // potrf_diagr_kernel only writes the lower block triangle of an inverted diagonal block
  CK(cudaMemsetAsync(w.Dinv, 0, (size_t)Np * 128 * sizeof(double), h->stream));
  CK(cudaMalloc(&w.logdet, (size_t)w.Nb * sizeof(double)));
  CK(cudaMalloc(&w.alpha, (size_t)128 * Np * sizeof(double)));
  set_mat(w, M_L, w.Lbuf, Np + 128, Np, Np);
  set_mat(w, M_DINV, w.Dinv, Np, 128, 128);
  build_potrf_plan(w);
  int rc = upload_plan(h, w.potrf);
  if (rc) return rc;
  return 0;
}

int ensure_inverse_buffers(H* h, Workspace& w) {
  if (w.Ubuf) return 0;
  const long long Np = w.Np;
  CK(cudaMalloc(&w.Ubuf, (size_t)Np * Np * sizeof(double)));
  CK(cudaMalloc(&w.Wbuf, (size_t)Np * Np * sizeof(double)));
  set_mat(w, M_U, w.Ubuf, Np, Np, Np);
  set_mat(w, M_W, w.Wbuf, Np, Np, Np);
  build_trtri_plan(w);
  build_syrk_plan(w);
  int rc = upload_plan(h, w.trtri);
  if (rc) return rc;
  for (int i = 0; i < w.n_phase; ++i)
    if ((rc = upload_plan(h, w.trtri_phase[i]))) return rc;
  if ((rc = upload_plan(h, w.trtri_late))) return rc;
  return upload_plan(h, w.syrk);
}

// ------------------------------------------------------------------------------------------------ small kernels
// Seed the diagonal tiles of U (= W_kk^T, upper) and W (= W_kk, lower) from the inverted diagonal blocks.
__global__ void seed_diag_kernel(const double* __restrict__ Dinv, double* __restrict__ U, double* __restrict__ W,
                                 long long ld, int k0) {
  const int k = k0 + blockIdx.x;
  const double* src = Dinv + (long long)k * 128 * 128;
  for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
    const int i = e >> 7, j = e & 127;
    const long long off = ((long long)k * 128 + i) * ld + (long long)k * 128 + j;
    W[off] = src[i * 128 + j];
    U[off] = src[j * 128 + i];
  }
}

// out[o*ldo + i] = sum_{k>=i} U[i][k] * z[o*ldz + k]        (alpha = U z, U upper triangular)
// or, lower == true:  sum_{k<=i} W[i][k] * z[k*sz_k + o*sz_o]
__global__ void __launch_bounds__(256) trigemv_kernel(const double* __restrict__ T, long long ld, long long n,
                                                      bool lower, const double* __restrict__ z, long long sz_k,
                                                      long long sz_o, int n_out, double* __restrict__ out,
                                                      long long so_i, long long so_o) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + warp;
  if (i >= n) return;
  const long long k0 = lower ? 0 : i, k1 = lower ? i + 1 : n;
  for (int o = 0; o < n_out; ++o) {
    // four independent partial sums: four row segments in flight per lane (the kernel is bound by load latency:
    // ncu long-scoreboard stall 129 cycles per issue with one), combined in a fixed order
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long k = k0 + lane;
    for (; k + 96 < k1; k += 128) {
      const double t0 = T[i * ld + k], t1 = T[i * ld + k + 32], t2 = T[i * ld + k + 64], t3 = T[i * ld + k + 96];
      s0 = fma(t0, z[k * sz_k + o * sz_o], s0);
      s1 = fma(t1, z[(k + 32) * sz_k + o * sz_o], s1);
      s2 = fma(t2, z[(k + 64) * sz_k + o * sz_o], s2);
      s3 = fma(t3, z[(k + 96) * sz_k + o * sz_o], s3);
    }
    for (; k < k1; k += 32) s0 = fma(T[i * ld + k], z[k * sz_k + o * sz_o], s0);
    double s = (s0 + s1) + (s2 + s3);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[i * so_i + o * so_o] = s;
  }
}

// nll = 0.5 * sum_o ||z_o||^2 + n_out * sum_k logdet[k] + 0.5 * n * n_out * log(2 pi)   (gp_functions.py:147-164)
__global__ void __launch_bounds__(1024) nll_kernel(const double* __restrict__ z, long long ldz, int n_out, long long n,
                                                   long long n_data, const double* __restrict__ logdet, int nb,
                                                   double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int o = 0; o < n_out; ++o)
    for (long long k = threadIdx.x; k < n; k += 1024) {
      const double v = z[o * ldz + k];
      s = fma(v, v, s);
    }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double ld = 0.0;
    for (int k = 0; k < nb; ++k) ld += logdet[k];
    out[0] = 0.5 * sh[0] + n_out * ld + 0.5 * (double)n_data * n_out * 1.8378770664093454835606594728112;
  }
}

// dst (rows x cols, ldd) <- padded copy of src (n x n, lds): identity outside, optionally lower only.
__global__ void pad_copy_kernel(const double* __restrict__ src, long long n, long long lds, double* __restrict__ dst,
                                long long np, long long ldd) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np * np) return;
  const long long i = e / np, j = e % np;
  double v = (i == j) ? 1.0 : 0.0;
  if (i < n && j < n) v = src[i * lds + j];
  dst[i * ldd + j] = v;
}

// mode 0: out = lower(src) with zero upper (Cholesky factor); mode 1: out = symmetric from lower (K^-1)
__global__ void unpack_kernel(const double* __restrict__ src, long long lds, long long n, int mode,
                              double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const long long i = e / n, j = e % n;
  double v;
  if (mode == 0) v = (j <= i) ? src[i * lds + j] : 0.0;
  else v = (j <= i) ? src[i * lds + j] : src[j * lds + i];
  out[e] = v;
}

__global__ void zero_aug_kernel(double* __restrict__ L, long long np) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 128 * np) L[np * np + e] = 0.0;
}

// K(Xc, Xc) (+ sn^2 on the diagonal, np.eye as in python_kernels.py:41) into the padded covariance buffer; the
// predictive covariance then subtracts Vt Vt^T from it (gp_functions.predict_f, :484-487).
template <int KIND>
__global__ void cov_prior_kernel(const double* __restrict__ Xc, int m, KernelHyp h, double* __restrict__ C, long long ldc,
                                 int m_pad) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)m_pad * m_pad) return;
  const int i = (int)(e / m_pad), j = (int)(e % m_pad);
  double v = 0.0;
  if (i < m && j < m) {
    double d2 = 0.0;
    for (int d = 0; d < h.D; ++d) {
      const double l = ell_of(h, d);
      const double u = Xc[(long long)i * h.D + d] / l - Xc[(long long)j * h.D + d] / l;
      d2 = fma(u, u, d2);
    }
    double gd;
    kern_eval<KIND>(d2, h.sf * h.sf, v, gd);
    if (i == j) v += h.sn * h.sn;
  }
  C[(long long)i * ldc + j] = v;
}

// out (m x m, contiguous) = max(C, 1e-10) element-wise (gp_functions.py:488)
__global__ void cov_finish_kernel(const double* __restrict__ C, long long ldc, int m, double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)m * m) return;
  const int i = (int)(e / m), j = (int)(e % m);
  out[e] = fmax(C[(long long)i * ldc + j], 1e-10);
}

// mean_ij |x_i - x_j| over all ordered pairs (GaussianProcess.infer_hyperparameters, gaussian_process.py:126-128):
// one block per 64 rows, fixed-order reductions, partial sums per block.
__global__ void __launch_bounds__(256) pairwise_distance_sum_kernel(const double* __restrict__ X, long long n, int d,
                                                                   double* __restrict__ part) {
  __shared__ double xi[64][MAX_D];
  __shared__ double red[256];
  const long long r0 = (long long)blockIdx.x * 64;
  for (int e = threadIdx.x; e < 64 * d; e += 256) {
    const long long r = r0 + e / d;
    xi[e / d][e % d] = r < n ? X[r * d + e % d] : 0.0;
  }
  __syncthreads();
  const int nrow = (n - r0 < 64) ? (int)(n - r0) : 64;
  double s = 0.0;
  for (long long j = threadIdx.x; j < n; j += 256) {
    double xj[MAX_D];
    for (int k = 0; k < d; ++k) xj[k] = X[j * d + k];
    for (int r = 0; r < nrow; ++r) {
      double a = 0.0;
      for (int k = 0; k < d; ++k) {
        const double u = xi[r][k] - xj[k];
        a = fma(u, u, a);
      }
      s += sqrt(a);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// ------------------------------------------------------------------------------------------------ helpers
void stage_begin(H* h, int s) {
  cudaEventRecord(h->ev0[s], h->stream);
  h->ev_used[s] = true;
}
void stage_end(H* h, int s) { cudaEventRecord(h->ev1[s], h->stream); }
void stages_reset(H* h) {
  for (int s = 0; s < GPB_NUM_STAGES; ++s) {
    h->ev_used[s] = false;
    h->stage_ms[s] = 0.0;
  }
}
void stages_collect(H* h) {
  for (int s = 0; s < GPB_NUM_STAGES; ++s)
    if (h->ev_used[s]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->ev0[s], h->ev1[s]) == cudaSuccess) h->stage_ms[s] = ms;
    }
}

int set_hyp(H* h, int kind, const double* theta, int n_ell) {
  if (kind != GPB_KERNEL_RBF && kind != GPB_KERNEL_MATERN52) return fail(h, -1, "unknown kernel kind");
  if (!(n_ell == 1 || n_ell == h->D)) return fail(h, -1, "n_ell must be 1 or the input dimension");
  h->hyp.kind = kind;
  h->hyp.D = h->D;
  h->hyp.n_ell = n_ell;
  for (int d = 0; d < n_ell; ++d) h->hyp.ell[d] = theta[d];
  h->hyp.sf = theta[n_ell];
  h->hyp.sn = theta[n_ell + 1];
  return 0;
}

// Radial-factor stashes for the gradient contraction (one array for RBF, two for Matern-5/2), Nb (Nb + 1) / 2 blocks of
// 128 x 128 each.  Returns false -- and the caller recomputes the factors in the contraction -- when they are switched off
// (GPB_STASH=0) or the device cannot spare the memory.
bool ensure_stash(Workspace& w, bool need_kp) {
  if (!g_stash) return false;
  // one spare block row: the last tall tile of the contraction may reach 128 rows past the padded matrix
  const size_t need = (size_t)((w.Nb + 1) * (w.Nb + 2) / 2) * 128 * 128;
  auto grow = [&](double** p, size_t* cap) {
    if (*cap >= need) return true;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (cudaMalloc(p, need * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      *p = nullptr;
      return false;
    }
    *cap = need;
    return true;
  };
  if (!grow(&w.stashG, &w.stash_cap)) return false;
  if (need_kp && !grow(&w.stashK, &w.stashK_cap)) return false;
  return true;
}

template <int KIND>
int assemble_dispatch(H* h, Workspace& w, bool stash) {
  double* G = stash ? w.stashG : nullptr;
  double* Kp = stash ? w.stashK : nullptr;
  const double sf2 = h->hyp.sf * h->hyp.sf, sn2 = h->hyp.sn * h->hyp.sn;
  dim3 grid((unsigned)(w.Np / AS_TN), (unsigned)(w.Np / AS_TM));
#define GPB_AS(DPV)                                                                                              \
  case DPV:                                                                                                      \
    assemble_sym_kernel<KIND, DPV><<<grid, AS_THREADS, 0, h->stream>>>(h->dXs, h->dXsT, w.Np, (int)w.N, (int)w.Np, \
                                                                      sf2, sn2, w.Lbuf, w.Np, G, Kp);           \
    break;
  switch (h->DP) {
    GPB_AS(4) GPB_AS(8) GPB_AS(12) GPB_AS(16) GPB_AS(24) GPB_AS(32)
    default: return fail(h, -1, "unsupported input dimension");
  }
#undef GPB_AS
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <int KIND>
int contract_dispatch(H* h, Workspace& w, dim3 grid) {
  const double sf2 = h->hyp.sf * h->hyp.sf;
#define GPB_GC(DPV)                                                                                               \
  case DPV:                                                                                                       \
    grad_contract_kernel<KIND, DPV><<<grid, AS_THREADS, 0, h->stream>>>(h->dXs, h->dXsT, w.Np, (int)w.N, sf2, w.Lbuf, \
                                                                       w.Np, w.alpha, w.Np, h->n_out, h->dpart); \
    break;
  switch (h->DP) {
    GPB_GC(4) GPB_GC(8) GPB_GC(12) GPB_GC(16) GPB_GC(24) GPB_GC(32)
    default: return fail(h, -1, "unsupported input dimension");
  }
#undef GPB_GC
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <int KIND, int DPV, int TM>
int contract_stash_launch_tm(H* h, Workspace& w) {
  dim3 grid((unsigned)(w.Np / AS_TN), (unsigned)((w.Np + TM - 1) / TM));
  const size_t ncta = (size_t)grid.x * grid.y;
  if (int rc = ensure_cap(h, &h->dpart, &h->dpart_cap, ncta * (DPV + 2))) return rc;
  auto kern = grad_contract_stash_kernel<KIND, DPV, TM>;
  constexpr int smem = GcSmem<KIND, DPV, TM>::BYTES;
  static bool configured[64] = {};   // per device, like launch_cfg
  if (!configured[h->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[h->device & 63] = true;
  }
  kern<<<grid, AS_THREADS, smem, h->stream>>>(h->dXs, h->dXsT, w.Np, (int)w.N, (int)w.Np, w.Lbuf, w.Np, w.alpha, w.Np,
                                              h->n_out, w.stashG, w.stashK, h->dpart);
  h->launches++;
  CK(cudaGetLastError());
  reduce_partials_kernel<<<DPV + 2, 256, 0, h->stream>>>(h->dpart, (int)ncta, DPV + 2, h->dscal + 1);
  h->launches++;
  return 0;
}

template <int KIND, int DPV>
int contract_stash_launch(H* h, Workspace& w) {
  // tall tiles once they still give every SM several CTAs (64 block columns x 32 tile rows / 2 at 8192 rows)
  if (w.Nb >= 64) return contract_stash_launch_tm<KIND, DPV, GcTile<DPV>::TM>(h, w);
  return contract_stash_launch_tm<KIND, DPV, GcTile<DPV>::TM_SMALL>(h, w);
}

// Gradient sums into dscal[1 ...] (DP + 2 slots): from the stash when the assembly of this evaluation left one,
// otherwise with the radial factors recomputed pair by pair.
template <int KIND>
int contract_and_reduce(H* h, Workspace& w, bool stash) {
  if (stash) {
    switch (h->DP) {
      case 4: return contract_stash_launch<KIND, 4>(h, w);
      case 8: return contract_stash_launch<KIND, 8>(h, w);
      case 12: return contract_stash_launch<KIND, 12>(h, w);
      case 16: return contract_stash_launch<KIND, 16>(h, w);
      case 24: return contract_stash_launch<KIND, 24>(h, w);
      case 32: return contract_stash_launch<KIND, 32>(h, w);
      default: return fail(h, -1, "unsupported input dimension");
    }
  }
  const int width = h->DP + 2;
  dim3 grid((unsigned)(w.Np / AS_TN), (unsigned)(w.Np / AS_TM));
  const size_t ncta = (size_t)grid.x * grid.y;
  if (int rc = ensure_cap(h, &h->dpart, &h->dpart_cap, ncta * width)) return rc;
  if (int rc = contract_dispatch<KIND>(h, w, grid)) return rc;
  reduce_partials_kernel<<<width, 256, 0, h->stream>>>(h->dpart, (int)ncta, width, h->dscal + 1);
  h->launches++;
  return 0;
}

template <int KIND>
int cross_dispatch(H* h, Workspace& w, const double* dXc, long long mc, double* KsT, long long ldk, double* mean) {
  const unsigned grid = (unsigned)((mc + CR_ROWS - 1) / CR_ROWS);
#define GPB_CM(DPV)                                                                                                  \
  case DPV:                                                                                                          \
    cross_mean_kernel<KIND, DPV><<<grid, AS_THREADS, 0, h->stream>>>(dXc, (int)mc, h->D, h->hyp, h->dXsT, w.Np, (int)w.N, \
                                                                    (int)w.Np, w.alpha, w.Np, mean ? h->n_out : 0, KsT, ldk, mean); \
    break;
  switch (h->DP) {
    GPB_CM(4) GPB_CM(8) GPB_CM(12) GPB_CM(16) GPB_CM(24) GPB_CM(32)
    default: return fail(h, -1, "unsupported input dimension");
  }
#undef GPB_CM
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// Stages shared by gpb_nll and gpb_factorize: K(theta) -> L, z, logdet.
// early_inverse: also queue the early phases of the triangular inverse (the caller must have called
// ensure_inverse_buffers and must finish with triangular_inverse(h, w, true)).
int assemble_and_factor(H* h, bool early_inverse = false, bool stash = false) {
  Workspace& w = h->ws;
  stage_begin(h, GPB_STAGE_ASSEMBLE);
  CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  {
    const long long tot = w.Np * h->DP;
    prescale_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->dX, (int)w.N, h->hyp, h->DP, h->dXs,
                                                                        h->dXsT, w.Np, (int)w.Np);
    h->launches++;
  }
  int rc = (h->hyp.kind == KIND_RBF) ? assemble_dispatch<KIND_RBF>(h, w, stash) : assemble_dispatch<KIND_MATERN52>(h, w, stash);
  if (rc) return rc;
  {
    const long long tot = 128 * w.Np;
    fill_aug_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->dy, (int)w.N, h->n_out, (int)w.Np, w.Lbuf,
                                                                        w.Np);
    h->launches++;
  }
  CK(cudaGetLastError());
  stage_end(h, GPB_STAGE_ASSEMBLE);
  stage_begin(h, GPB_STAGE_POTRF);
  rc = run_plan(h, w, w.potrf, true);
  if (rc) return rc;
  if (early_inverse && w.n_phase > 0) {
    // The Cholesky launches above are only queued; stream3 picks up each phase when its event fires.
    int k0 = 0;
    for (int i = 0; i < w.n_phase; ++i) {
      CK(cudaStreamWaitEvent(h->stream3, h->events[w.phase_ev[i]], 0));
      if (w.phase_cut[i] > k0) {
        const int c0 = std::max(k0, w.dinv_partial_from);
        if (c0 < w.phase_cut[i]) {
          complete_inverse_kernel<<<(unsigned)(w.phase_cut[i] - c0), TP_THREADS, 0, h->stream3>>>(w.Lbuf, w.Np, c0, w.Dinv);
          h->launches++;
        }
        seed_diag_kernel<<<(unsigned)(w.phase_cut[i] - k0), 256, 0, h->stream3>>>(w.Dinv, w.Ubuf, w.Wbuf, w.Np, k0);
        h->launches++;
        CK(cudaGetLastError());
        k0 = w.phase_cut[i];
      }
      if ((rc = run_plan(h, w, w.trtri_phase[i], false, h->stream3))) return rc;
    }
    CK(cudaEventRecord(h->ev_early, h->stream3));
  }
  stage_end(h, GPB_STAGE_POTRF);
  return 0;
}

// U = L^-T, W = L^-1 from L and the inverted diagonal blocks.
// early_done: the phases queued by assemble_and_factor(h, true) cover the leading blocks; finish the rest.
int triangular_inverse(H* h, Workspace& w, bool early_done = false) {
  int rc = ensure_inverse_buffers(h, w);
  if (rc) return rc;
  const bool early = early_done && w.n_phase > 0;
  const int k0 = early ? w.phase_cut[w.n_phase - 1] : 0;
  const int c0 = std::max(k0, w.dinv_partial_from);
  if (c0 < (int)w.Nb) {   // the chain only left the 32x32 diagonal blocks of these inverted diagonal blocks
    complete_inverse_kernel<<<(unsigned)(w.Nb - c0), TP_THREADS, 0, h->stream>>>(w.Lbuf, w.Np, c0, w.Dinv);
    h->launches++;
  }
  seed_diag_kernel<<<(unsigned)(w.Nb - k0), 256, 0, h->stream>>>(w.Dinv, w.Ubuf, w.Wbuf, w.Np, k0);
  h->launches++;
  CK(cudaGetLastError());
  if (!early) return run_plan(h, w, w.trtri, false);
  CK(cudaStreamWaitEvent(h->stream, h->ev_early, 0));
  return run_plan(h, w, w.trtri_late, false);
}

int compute_alpha(H* h, Workspace& w, int n_out) {
  // alpha_o = U z_o with z_o = row o of the y^T block of Lbuf
  trigemv_kernel<<<(unsigned)((w.Np + 7) / 8), 256, 0, h->stream>>>(w.Ubuf, w.Np, w.Np, false, w.Lbuf + w.Np * w.Np, 1,
                                                                    w.Np, n_out, w.alpha, 1, w.Np);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int check_info(H* h) {
  int info = 0;
  CK(cudaMemcpyAsync(&info, h->dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (info > 0) {
    h->err = "matrix is not positive definite: non-positive pivot at row " + std::to_string(info);
    return info;
  }
  return 0;
}

int need_train(H* h) {
  if (!h) return -1;
  if (h->ws.N <= 0 || !h->dX) return fail(h, -3, "gpb_set_train has not been called");
  return 0;
}

int load_matrix_into_L(H* h, Workspace& w, const double* A, long long n) {
  // stage A on the device if it is a host pointer, then pad-copy into Lbuf
  int rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)n * n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->dstage, A, (size_t)n * n * sizeof(double), cudaMemcpyDefault, h->stream));
  const long long tot = w.Np * w.Np;
  pad_copy_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->dstage, n, n, w.Lbuf, w.Np, w.Np);
  zero_aug_kernel<<<(unsigned)((128 * w.Np + 255) / 256), 256, 0, h->stream>>>(w.Lbuf, w.Np);
  h->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int invert_diag_blocks(H* h, Workspace& w) {
  w.dinv_partial_from = 1 << 30;
  potrf_diag_kernel<false><<<(unsigned)w.Nb, DIAG_THREADS, DIAG_SMEM, h->stream>>>(w.Lbuf, w.Np, 0, w.Dinv, w.logdet,
                                                                                   h->dinfo, (int)w.N);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ active learning
// Device address of `count` doubles at `p`: device pointers pass through, host data is staged at *cursor.
int acq_stage_in(H* h, const double* p, size_t count, double** cursor, const double** out) {
  *out = nullptr;
  if (!p) return 0;
  if (is_device_ptr(p)) {
    *out = p;
    return 0;
  }
  CK(cudaMemcpyAsync(*cursor, p, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  *out = *cursor;
  *cursor += count;
  return 0;
}

int acq_grid(long long M) {
  const long long want = (M + ACQ_THREADS - 1) / ACQ_THREADS;
  return (int)std::max<long long>(1, std::min<long long>(want, 148 * 8));
}

// stats[2c], stats[2c+1] = column-wise min / max of v (M, ncol) [after sqrt when transform == 1]
int acq_minmax(H* h, const double* v, long long M, int ncol, int transform, double* part, double* stats) {
  const int nblk = acq_grid(M);
  acq_minmax_partial_kernel<<<dim3(nblk, ncol), ACQ_THREADS, 0, h->stream>>>(v, M, ncol, transform, part);
  acq_minmax_final_kernel<<<ncol, ACQ_THREADS, 0, h->stream>>>(part, nblk, stats);
  h->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int fill_acq_args(H* h, AcqArgs& a, int kind, long long M, int n_out, const double* params, int n_params) {
  memset(&a, 0, sizeof(a));
  a.kind = kind;
  a.M = M;
  a.n_out = n_out;
  auto need = [&](int n) { return n_params >= n && (n == 0 || params); };
  switch (kind) {
    case ACQ_VARIANCE: return 0;
    case ACQ_DISTANCE_PENALTY:   // [c1, d, last_0 .. last_{d-1}]
      if (!need(2)) return fail(h, -1, "distance penalty: params = [weight, d, last point...]");
      a.p0 = params[0];
      a.d = (int)params[1];
      if (a.d < 1 || a.d > ACQ_MAX_DIM || !need(2 + a.d)) return fail(h, -1, "distance penalty: bad dimension");
      for (int k = 0; k < a.d; ++k) a.last[k] = params[2 + k];
      return 0;
    case ACQ_WEIGHTED:           // [weight]
      if (!need(1)) return fail(h, -1, "weighted exploration: params = [weight]");
      a.p0 = params[0];
      return 0;
    case ACQ_POI:                // [ymax_0 ..]
      if (!need(n_out)) return fail(h, -1, "probability of improvement: params = [ymax per output]");
      for (int o = 0; o < n_out; ++o) a.ymax[o] = params[o];
      return 0;
    case ACQ_EI:                 // loss: none;  prepare: [xi, find_min, ymax_0 ..]
      if (n_params == 0) return 0;
      // fallthrough
    case ACQ_EI2:                // [xi, find_min, ymax_0 ..]
      if (!need(2 + n_out)) return fail(h, -1, "expected improvement: params = [xi, find_min, ymax per output]");
      a.p0 = params[0];
      a.p1 = params[1];
      for (int o = 0; o < n_out; ++o) a.ymax[o] = params[2 + o];
      return 0;
    default: return fail(h, -1, "unknown acquisition kind");
  }
}

// Dense K(Xa, Xb) [+ dK] from device-resident raw inputs into device outputs (kernel_matrix_kernel).
int dense_kernel_device(H* h, const KernelHyp& hy, const double* dXa, long long na, const double* dXb, long long nb,
                        double* dKo, double* ddK, bool timed) {
  const int P = hy.n_ell + 2;
  const int DP = pick_dp(hy.D);
  const long long na_pad = round_up(na, KM_ROWS), nb_pad = round_up(nb, KM_COLS);
  int rc = ensure_cap(h, &h->dkm, &h->dkm_cap, (size_t)na_pad * DP + (size_t)DP * nb_pad);
  if (rc) return rc;
  double* dSa = h->dkm;                          // scaled rows of Xa (na_pad x DP)
  double* dSb = h->dkm + (size_t)na_pad * DP;    // scaled columns of Xb (DP x nb_pad)
  prescale_kernel<<<(unsigned)((na_pad * DP + 255) / 256), 256, 0, h->stream>>>(dXa, (int)na, hy, DP, dSa, nullptr, 0,
                                                                              (int)na_pad);
  prescale_kernel<<<(unsigned)((nb_pad * DP + 255) / 256), 256, 0, h->stream>>>(dXb, (int)nb, hy, DP, nullptr, dSb,
                                                                              nb_pad, (int)nb_pad);
  h->launches += 2;
  if (timed) stage_begin(h, GPB_STAGE_ASSEMBLE);
  dim3 grid((unsigned)(nb_pad / KM_COLS), (unsigned)(na_pad / KM_ROWS));
  const size_t stage_doubles = (size_t)4 * KM_COLS * P;
  // Two stages (2 barriers per pass, 4 CTAs/SM) measured faster than three (1 barrier, 2 CTAs/SM): 6.98 vs 6.01 TB/s.
  int nbuf = 2;
  if (const char* e = getenv("GPB_KM_NBUF"))
    if (atoi(e) == 3 && 3 * stage_doubles * sizeof(double) <= 74 * 1024) nbuf = 3;
  const size_t smem = ((size_t)KM_ROWS * DP + DP + (DP & 1) + (ddK ? nbuf * stage_doubles : 0)) * sizeof(double);
#define GPB_KM(KINDV, DPV)                                                                                          \
  {                                                                                                                 \
    auto kern = kernel_matrix_kernel<KINDV, DPV>;                                                                   \
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                         \
    kern<<<grid, KM_THREADS, smem, h->stream>>>(dSa, (int)na, dSb, nb_pad, (int)nb, hy, dKo, ddK, nbuf);            \
  }
#define GPB_KM_DP(KINDV)                                                                                            \
  switch (DP) {                                                                                                     \
    case 4: GPB_KM(KINDV, 4) break;                                                                                 \
    case 8: GPB_KM(KINDV, 8) break;                                                                                 \
    case 12: GPB_KM(KINDV, 12) break;                                                                               \
    case 16: GPB_KM(KINDV, 16) break;                                                                               \
    case 24: GPB_KM(KINDV, 24) break;                                                                               \
    default: GPB_KM(KINDV, 32) break;                                                                               \
  }
  if (hy.kind == KIND_RBF) GPB_KM_DP(KIND_RBF) else GPB_KM_DP(KIND_MATERN52)
#undef GPB_KM_DP
#undef GPB_KM
  h->launches++;
  CK(cudaGetLastError());
  if (timed) stage_end(h, GPB_STAGE_ASSEMBLE);
  return 0;
}

// out[r*so_r + p*so_p] = sum_c T[r][c][p] * a[c]  +  scale * sum_c Kmat[r][c] * B[p*ldb + c]     (B, Kmat optional)
// One warp per row r; lanes stride over c.  Used by the marginal-variance path (small problems, not a hot kernel).
__global__ void __launch_bounds__(256) rowwise_contract_kernel(const double* __restrict__ T, const double* __restrict__ Kmat,
                                                               int R, int C, int P, const double* __restrict__ a,
                                                               const double* __restrict__ B, long long ldb, double scale,
                                                               double* __restrict__ out, long long so_r, long long so_p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= R) return;
  const double* Tr = T + r * (long long)C * P;
  for (int p = 0; p < P; ++p) {
    double s = 0.0;
    for (int c = lane; c < C; c += 32) {
      s = fma(Tr[(long long)c * P + p], a[c], s);
      if (B != nullptr) s = fma(scale * Kmat[r * (long long)C + c], B[(long long)p * ldb + c], s);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[r * so_r + p * so_p] = s;
  }
}

// out[c] = dm_c^T H dm_c + pv[c]
__global__ void quadform_kernel(const double* __restrict__ dm, int M, int P, const double* __restrict__ H,
                                const double* __restrict__ pv, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= M) return;
  double s = 0.0;
  for (int p = 0; p < P; ++p) {
    double hq = 0.0;
    for (int q = 0; q < P; ++q) hq = fma(H[p * P + q], dm[(long long)c * P + q], hq);
    s = fma(dm[(long long)c * P + p], hq, s);
  }
  out[c] = s + pv[c];
}

}  // namespace


// ------------------------------------------------------------------------------------------------ eigen-decomposition
constexpr int EIG_MCAP = 4096;   // most Lanczos steps (= largest projected problem the QL kernel holds in shared memory)

// (Re)allocate the work buffers for an n x n matrix; invalidates the resident matrix when the size changes.
int eig_reserve(H* h, int n) {
  H::Eig& e = h->eig;
  const int mcap = std::min(n, EIG_MCAP);
  if (e.n == n && e.buf) return 0;
  const size_t nn = (size_t)n, mm = (size_t)mcap;
  const size_t need = (mm + 1) * nn + nn + 5 * mm + 8 + mm * mm + mm * nn + 3 * mm + nn;
  int rc = ensure_cap(h, &e.buf, &e.buf_cap, need);
  if (rc) return rc;
  double* p = e.buf;
  e.Vt = p, p += (mm + 1) * nn;
  e.w = p, p += nn;
  e.c1 = p, p += mm;
  e.c2 = p, p += mm;
  e.alpha = p, p += mm;
  e.beta = p, p += mm + 8;
  e.d = p, p += mm;
  e.Zt = p, p += mm * mm;
  e.Qt = p, p += mm * nn;
  e.wsel = p, p += mm;
  e.qy = p, p += mm;
  e.t = p, p += mm;
  e.avec = p, p += nn;
  if (e.ibuf) CK(cudaFree(e.ibuf));
  e.ibuf = nullptr;
  CK(cudaMalloc(&e.ibuf, (mm + 4) * sizeof(int)));
  if ((rc = ensure_cap(h, &e.K, &e.K_cap, nn * nn))) return rc;
  e.n = n;
  e.mcap = mcap;
  e.m = e.k = 0;
  e.matrix_ready = false;
  return 0;
}

// Start a Lanczos run on the resident matrix: v_0 = normalised deterministic pseudo-random vector.
int eig_begin(H* h) {
  H::Eig& e = h->eig;
  CK(cudaMemsetAsync(e.ibuf + e.mcap, 0, 4 * sizeof(int), h->stream));
  eig_start_kernel<<<(unsigned)((e.n + 255) / 256), 256, 0, h->stream>>>(e.w, e.n, 0x5DEECE66Dull);
  eig_next_kernel<<<1, 1024, 0, h->stream>>>(e.w, e.n, e.Vt, nullptr);
  h->launches += 2;
  CK(cudaGetLastError());
  e.m = e.k = 0;
  e.solved_m = e.vectors_k = e.vectors_m = 0;
  e.restarts = 0;
  return 0;
}

// One pass of classical Gram-Schmidt of e.w against the first cnt basis vectors (coefficients into c).
void eig_gs_pass(H* h, int cnt, double* c, const double* c_first, double* alpha_out) {
  H::Eig& e = h->eig;
  const int n = e.n;
  eig_dots_kernel<<<(unsigned)cnt, EIG_BLOCK, 0, h->stream>>>(e.Vt, n, n, e.w, c);
  eig_update_kernel<<<(unsigned)((n + EIG_BLOCK - 1) / EIG_BLOCK), EIG_BLOCK, 0, h->stream>>>(e.Vt, n, n, cnt, c, e.w, c_first,
                                                                                               alpha_out);
  h->launches += 2;
}

// Orthogonalise e.w against the first cnt basis vectors and leave the normalised result in row cnt of the basis.
// Two Gram-Schmidt passes, then more while a pass still removes more than half of the vector (a vector that lies
// numerically inside span(V) needs them: "twice is enough" only holds for vectors that do not).  norms <- {|w| before,
// after the first pass, at the end}; alpha_out (device, may be nullptr) <- coefficient along the last basis vector.
int eig_orthonormalise(H* h, int cnt, double* alpha_out, double norms[3], double* alpha_host = nullptr) {
  H::Eig& e = h->eig;
  const int n = e.n;
  double* vnext = e.Vt + (size_t)cnt * n;       // the basis buffer has mcap + 1 rows
  double* dn = e.beta + e.mcap;                 // 8 spare doubles behind the beta array
  eig_next_kernel<<<1, 1024, 0, h->stream>>>(e.w, n, nullptr, dn);
  eig_gs_pass(h, cnt, e.c1, nullptr, nullptr);
  eig_next_kernel<<<1, 1024, 0, h->stream>>>(e.w, n, nullptr, dn + 1);
  eig_gs_pass(h, cnt, e.c2, e.c1, alpha_out);
  eig_next_kernel<<<1, 1024, 0, h->stream>>>(e.w, n, vnext, dn + 2);
  h->launches += 3;
  CK(cudaMemcpyAsync(norms, dn, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (alpha_out && alpha_host) CK(cudaMemcpyAsync(alpha_host, alpha_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // the one host synchronisation of a regular step
  double prev = norms[1];
  for (int extra = 0; extra < 4 && norms[2] > 0.0 && norms[2] < 0.5 * prev; ++extra) {
    prev = norms[2];
    eig_gs_pass(h, cnt, e.c2, nullptr, nullptr);
    eig_next_kernel<<<1, 1024, 0, h->stream>>>(e.w, n, vnext, dn + 2);
    h->launches++;
    CK(cudaMemcpyAsync(norms + 2, dn + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaGetLastError());
  return 0;
}

// Lanczos steps m .. m_target-1 with full re-orthogonalisation.  After every step the host reads alpha_j and the norms of
// the residual: a residual at the rounding level of the vector it was computed from (64 eps |A v_j|) means that span(V)
// is invariant under the matrix (rank-deficient or repeated spectrum) -- what is left is rounding noise, not a Krylov
// direction.  The run then continues from a fresh pseudo-random vector orthogonalised against V, with beta_j = 0 in T,
// so that T stays the projection V^T A V (block diagonal) to working accuracy.
int eig_extend(H* h, int m_target) {
  H::Eig& e = h->eig;
  const int n = e.n;
  if (n <= EIG_SMALL_N && g_eig_small) {   // small matrix: the whole recurrence in one CTA, no per-step round trips
    if (e.m >= m_target) return 0;
    const size_t smem = (size_t)(n + m_target) * sizeof(double);
    eig_lanczos_small_kernel<<<1, 1024, smem, h->stream>>>(e.K, n, e.Vt, e.alpha, e.beta, e.m, m_target, 0x5DEECE66Dull,
                                                          e.ibuf + e.mcap);
    h->launches++;
    CK(cudaGetLastError());
    int info[2] = {0, 0};
    CK(cudaMemcpyAsync(info, e.ibuf + e.mcap, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (info[1]) return fail(h, -4, "eigsh: non-finite Lanczos coefficients at step " + std::to_string(info[1]));
    e.restarts = info[0];
    e.m = m_target;
    return 0;
  }
  const unsigned gr = (unsigned)((n + 7) / 8);
  for (int j = e.m; j < m_target; ++j) {
    const double* vj = e.Vt + (size_t)j * n;
    eig_symv_kernel<<<gr, EIG_BLOCK, 0, h->stream>>>(e.K, n, n, vj, e.w);
    h->launches++;
    double norms[3], alpha_j = 0.0;
    int rc = eig_orthonormalise(h, j + 1, e.alpha + j, norms, &alpha_j);
    if (rc) return rc;
    CK(cudaMemcpyAsync(e.beta + j, e.beta + e.mcap + 2, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (!std::isfinite(alpha_j) || !std::isfinite(norms[2]))
      return fail(h, -4, "eigsh: non-finite Lanczos coefficients at step " + std::to_string(j + 1));
    e.m = j + 1;
    if (j + 1 < n && norms[2] <= 64.0 * 2.220446049250313e-16 * norms[0]) {
      CK(cudaMemsetAsync(e.beta + j, 0, sizeof(double), h->stream));
      eig_start_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(e.w, n, 0x5DEECE66Dull + 7919ull * (unsigned long long)(j + 1));
      h->launches++;
      if ((rc = eig_orthonormalise(h, j + 1, nullptr, norms))) return rc;
      e.restarts++;
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// The k largest eigenpairs of the resident matrix; leaves the eigenvalues (ascending, like eigsh) in e.wsel / e.w_host
// and the orthonormal Ritz vectors in the rows of e.Qt.  steps_out <- Lanczos steps used.
int eig_topk(H* h, int k, int* steps_out, int k_hint = 0) {
  H::Eig& e = h->eig;
  const int n = e.n;
  if (k < 1) return fail(h, -1, "eigsh: k must be positive");
  k = std::min(k, n);
  if (k > e.mcap)
    return fail(h, -4, "eigsh: " + std::to_string(k) + " eigenpairs requested, the device solver holds at most " +
                           std::to_string(e.mcap) + " Lanczos vectors");
  std::vector<double>& d = e.ritz;
  std::vector<double>& zlast = e.zlast;
  std::vector<int>& order = e.order;
  // k_hint: the largest request the caller expects for this matrix -- size the Krylov space (and diagonalise it) once
  int m = std::min(std::min(n, e.mcap), std::max(e.m, 2 * std::max(k, std::min(k_hint, e.mcap / 2)) + 16));
  while (true) {
    int rc = eig_extend(h, m);
    if (rc) return rc;
    if (e.solved_m != m) {   // the projected problem of this size has not been diagonalised yet
      const size_t smem = 4 * (size_t)m * sizeof(double);
      CK(cudaFuncSetAttribute(eig_tql2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      eig_tql2_kernel<<<1, EIG_QL_THREADS, smem, h->stream>>>(e.alpha, e.beta, m, e.d, e.Zt, m, e.ibuf + e.mcap + 2);
      h->launches++;
      CK(cudaGetLastError());
      d.resize(m);
      zlast.resize(m);
      int flags[2] = {0, 0};   // flags[1] <- QL failure
      CK(cudaMemcpyAsync(d.data(), e.d, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpy2DAsync(zlast.data(), sizeof(double), e.Zt + (m - 1), (size_t)m * sizeof(double), sizeof(double), m,
                           cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(&e.beta_last, e.beta + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(flags, e.ibuf + e.mcap + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (flags[1]) return fail(h, -4, "eigsh: the QL iteration on the projected tridiagonal matrix did not converge");
      for (int i = 0; i < m; ++i)
        if (!std::isfinite(d[i])) return fail(h, -4, "eigsh: non-finite Ritz values");
      order.resize(m);
      for (int i = 0; i < m; ++i) order[i] = i;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return d[a] > d[b] || (d[a] == d[b] && a < b); });
      e.ritz_scale = 0.0;
      for (int i = 0; i < m; ++i) e.ritz_scale = std::max(e.ritz_scale, std::fabs(d[i]));
      e.solved_m = m;
    }
    // residual of Ritz pair i: |beta_m z_{m,i}|; the basis spans everything once m == n
    bool ok = (m >= n);
    if (!ok) {
      ok = true;
      for (int i = 0; i < k && ok; ++i) ok = std::fabs(e.beta_last * zlast[order[i]]) <= 5e-13 * e.ritz_scale;
    }
    if (ok) break;
    if (m >= e.mcap)
      return fail(h, -4, "eigsh: " + std::to_string(k) + " eigenpairs did not converge within " + std::to_string(m) +
                             " Lanczos steps");
    m = std::min(std::min(n, e.mcap), m + std::max(16, m / 2));
  }
  // eigsh order: ascending, the largest k.  The Ritz vectors are formed when somebody needs them (eig_vectors).
  e.sel.resize(k);
  e.w_host.resize(k);
  for (int i = 0; i < k; ++i) {
    e.sel[i] = order[k - 1 - i];
    e.w_host[i] = d[e.sel[i]];
  }
  e.k = k;
  e.vectors_k = 0;
  if (steps_out) *steps_out = m;
  return 0;
}

// Ritz vectors of the k pairs selected by the last eig_topk: rows of e.Qt (orthonormal), eigenvalues in e.wsel.
int eig_vectors(H* h) {
  H::Eig& e = h->eig;
  if (e.vectors_k == e.k && e.vectors_m == e.solved_m) return 0;
  const int n = e.n, k = e.k, m = e.solved_m;
  CK(cudaMemcpyAsync(e.ibuf, e.sel.data(), k * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(e.wsel, e.w_host.data(), k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((n + EIG_BLOCK - 1) / EIG_BLOCK), (unsigned)k);
  eig_ritz_kernel<<<grid, EIG_BLOCK, 0, h->stream>>>(e.Vt, n, n, m, e.Zt, m, e.ibuf, e.Qt, n);
  eig_project_kernel<<<(unsigned)k, EIG_BLOCK, 0, h->stream>>>(e.Qt, n, n, nullptr, e.wsel, 0.0, nullptr, nullptr);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));   // e.sel / e.w_host are pageable host memory
  e.vectors_k = k;
  e.vectors_m = m;
  return 0;
}

// ================================================================================================ C ABI
extern "C" {

int gpb_version(void) { return 100; }

int gpb_create(int device, gpb_handle_t* out) {
  if (!out) return -1;
  *out = nullptr;
  H* h = new H();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete h;
    return -2;
  }
  // A blocking stream: implicitly ordered with the legacy default stream, which is what torch uses by default,
  // so device pointers handed in by the caller are safe to read without an explicit cross-stream event.
  // Priorities: panel chain (stream2) highest, main stream in the middle, early triangular inverse (stream3) lowest,
  // so that the fill-in work never delays a trailing update the chain is waiting for.
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int mid = (hi < lo - 1) ? lo - 1 : lo;   // numerically lower = higher priority
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamDefault, mid) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream2, cudaStreamDefault, hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, mid) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream5, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_caller, cudaEventDisableTiming) != cudaSuccess) {
      delete h;
      return -2;
    }
  }
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete h;
    return -2;
  }
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (cudaFuncSetAttribute(potrf_diagr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DR_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(potrf_diagr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DR_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(potrf_diag4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG4_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(potrf_diag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(potrf_diag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM) != cudaSuccess) {
    delete h;
    return -2;
  }
  for (int s = 0; s < GPB_NUM_STAGES; ++s) {
    cudaEventCreate(&h->ev0[s]);
    cudaEventCreate(&h->ev1[s]);
  }
  for (int s = 0; s < 3; ++s) cudaEventCreate(&h->evx[s]);
  stages_reset(h);
  cudaMalloc(&h->dscal, 64 * sizeof(double));
  cudaMalloc(&h->dinfo, sizeof(int));
  if (const char* e = getenv("GPB_CFG_UPDATE")) g_cfg_update = std::max(0, std::min(1, atoi(e)));
  if (const char* e = getenv("GPB_OUTER_BLOCKS")) g_outer_blocks = std::max(0, std::min(16, atoi(e)));
  if (const char* e = getenv("GPB_LOOKAHEAD")) g_lookahead = std::max(0, std::min(2, atoi(e)));
  if (const char* e = getenv("GPB_EARLY_TRTRI")) g_early_trtri = atoi(e) != 0;
  if (const char* e = getenv("GPB_SPLIT_T2")) g_split_t2 = atoi(e) != 0;
  if (const char* e = getenv("GPB_EARLY_CFG")) g_early_cfg = atoi(e);
  if (const char* e = getenv("GPB_EARLY_MASK")) g_early_mask = atoi(e);
  if (const char* e = getenv("GPB_FINE_BELOW")) g_fine_below = atoi(e);
  if (const char* e = getenv("GPB_STASH")) g_stash = atoi(e) != 0;
  if (const char* e = getenv("GPB_WIDE_ABOVE")) g_wide_above = atoi(e);
  if (const char* e = getenv("GPB_PDL")) g_pdl = atoi(e) != 0;
  if (const char* e = getenv("GPB_TRSM")) g_trsm = atoi(e) != 0;
  if (const char* e = getenv("GPB_TRSM_BELOW")) g_trsm_below = std::max(0, atoi(e));
  if (const char* e = getenv("GPB_EARLY_MIN_BLOCKS")) g_early_min_blocks = std::max(8, atoi(e));
  if (const char* e = getenv("GPB_TAIL_BLOCKS")) g_tail_blocks = std::max(0, atoi(e));
  if (const char* e = getenv("GPB_SUPER")) g_super = std::max(0, std::min(64, atoi(e)));
  if (const char* e = getenv("GPB_SUPER_TRTRI")) g_super_trtri = std::max(0, std::min(64, atoi(e)));
  if (const char* e = getenv("GPB_EIG_SMALL")) g_eig_small = atoi(e) != 0;
  if (const char* e = getenv("GPB_PRED_SUPER_M")) g_pred_super_m = std::max(0, std::min(128, atoi(e)));
  if (const char* e = getenv("GPB_PRED_SUPER_C")) g_pred_super_c = std::max(1, std::min(512, atoi(e)));
  if (const char* e = getenv("GPB_DIAG4")) g_diag4 = std::max(0, std::min(2, atoi(e)));
  *out = h;
  return 0;
}

int gpb_destroy(gpb_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_ws(h->ws);
  free_ws(h->tmp);
  h->pred.clear();
  cudaFree(h->dX);
  cudaFree(h->dy);
  cudaFree(h->dXs);
  cudaFree(h->dXsT);
  cudaFree(h->dscal);
  cudaFree(h->dpart);
  cudaFree(h->dinfo);
  cudaFree(h->dXc);
  cudaFree(h->dKsT);
  cudaFree(h->dVpart);
  cudaFree(h->dmean);
  cudaFree(h->dvar);
  cudaFree(h->dstage);
  cudaFree(h->dstage2);
  cudaFree(h->dkm);
  cudaFree(h->dbbq);
  cudaFree(h->dapp);
  cudaFree(h->dacq);
  cudaFree(h->eig.K);
  cudaFree(h->eig.buf);
  cudaFree(h->eig.ibuf);
  for (int s = 0; s < GPB_NUM_STAGES; ++s) {
    cudaEventDestroy(h->ev0[s]);
    cudaEventDestroy(h->ev1[s]);
  }
  for (int s = 0; s < 3; ++s) cudaEventDestroy(h->evx[s]);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream2);
  cudaStreamDestroy(h->stream3);
  cudaStreamDestroy(h->stream4);
  cudaStreamDestroy(h->stream5);
  cudaEventDestroy(h->ev_early);
  cudaEventDestroy(h->ev_caller);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int gpb_set_stream(gpb_handle_t h, void* cuda_stream) {
  if (!h) return -1;
  h->caller = reinterpret_cast<cudaStream_t>(cuda_stream);
  return 0;
}

const char* gpb_error_string(gpb_handle_t h) { return h ? h->err.c_str() : "null handle"; }

long long gpb_launch_count(gpb_handle_t h) { return h ? h->launches : 0; }

int gpb_stage_ms(gpb_handle_t h, double* out) {
  if (!h || !out) return -1;
  for (int s = 0; s < GPB_NUM_STAGES; ++s) out[s] = h->stage_ms[s];
  return 0;
}

int gpb_set_train(gpb_handle_t h, const double* X, const double* y, long long n, int d, int n_out) {
  if (!h || !X || !y || n < 1 || d < 1 || n_out < 1) return fail(h, -1, "gpb_set_train: bad argument");
  if (d > MAX_D) return fail(h, -1, "gpb_set_train: input dimension above 32 is not supported");
  if (n_out > 128) return fail(h, -1, "gpb_set_train: more than 128 outputs are not supported");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  const long long oldNp = h->ws.Np;
  const int oldDP = h->DP;
  int rc = resize_ws(h, h->ws, n);
  if (rc) return rc;
  h->D = d;
  h->DP = pick_dp(d);
  h->n_out = n_out;
  const long long Np = h->ws.Np;
  if (Np != oldNp || h->DP != oldDP || !h->dX) {
    cudaFree(h->dX);
    cudaFree(h->dy);
    cudaFree(h->dXs);
    cudaFree(h->dXsT);
    h->dX = h->dy = h->dXs = h->dXsT = nullptr;
    CK(cudaMalloc(&h->dX, (size_t)Np * MAX_D * sizeof(double)));
    CK(cudaMalloc(&h->dy, (size_t)Np * 128 * sizeof(double)));
    CK(cudaMalloc(&h->dXs, (size_t)Np * h->DP * sizeof(double)));
    CK(cudaMalloc(&h->dXsT, (size_t)Np * h->DP * sizeof(double)));
  }
  CK(cudaMemcpyAsync(h->dX, X, (size_t)n * d * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemcpyAsync(h->dy, y, (size_t)n * n_out * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->factor_ready = false;
  if (h->eig.from_train) h->eig.matrix_ready = false;
  return 0;
}

int gpb_nll(gpb_handle_t h, int kind, const double* theta, int n_ell, int want_grad, double* nll, double* grad) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!theta || !nll || (want_grad && !grad)) return fail(h, -1, "gpb_nll: null argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  if ((rc = set_hyp(h, kind, theta, n_ell))) return rc;
  Workspace& w = h->ws;
  h->factor_ready = false;
  stages_reset(h);
  stage_begin(h, GPB_STAGE_TOTAL);
  if (want_grad && (rc = ensure_inverse_buffers(h, w))) return rc;
  const bool stash = want_grad && ensure_stash(w, kind != KIND_RBF);
  if ((rc = assemble_and_factor(h, want_grad != 0, stash))) return rc;
  nll_kernel<<<1, 1024, 0, h->stream>>>(w.Lbuf + w.Np * w.Np, w.Np, h->n_out, w.Np, w.N, w.logdet, (int)w.Nb, h->dscal);
  h->launches++;
  const int P = n_ell + 2, width = h->DP + 2;
  if (want_grad) {
    stage_begin(h, GPB_STAGE_TRTRI);
    if ((rc = triangular_inverse(h, w, true))) return rc;
    stage_end(h, GPB_STAGE_TRTRI);
    stage_begin(h, GPB_STAGE_GRAD);
    if ((rc = compute_alpha(h, w, h->n_out))) return rc;
    stage_end(h, GPB_STAGE_GRAD);
    stage_begin(h, GPB_STAGE_SYRK);
    if ((rc = run_plan(h, w, w.syrk, false))) return rc;
    stage_end(h, GPB_STAGE_SYRK);
    cudaEvent_t g0 = h->evx[0], g1 = h->evx[1];  // the GRAD stage has two pieces; time the second with its own pair
    cudaEventRecord(g0, h->stream);
    rc = (kind == KIND_RBF) ? contract_and_reduce<KIND_RBF>(h, w, stash) : contract_and_reduce<KIND_MATERN52>(h, w, stash);
    if (rc) return rc;
    cudaEventRecord(g1, h->stream);
    stage_end(h, GPB_STAGE_TOTAL);
    double host[64];
    CK(cudaMemcpyAsync(host, h->dscal, (1 + width) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if ((rc = check_info(h))) return rc;
    stages_collect(h);
    float gms = 0.f;
    cudaEventElapsedTime(&gms, g0, g1);
    h->stage_ms[GPB_STAGE_GRAD] += gms;
    *nll = host[0];
    // dnll_p = 0.5 * sum_ij A_ij dK_ij,p   (gp_functions.py:180) with the per-parameter factors of
    // python_kernels.py:49-55 pulled out of the sums.
    const double* acc = host + 1;
    if (n_ell == 1) {
      double s = 0.0;
      for (int d = 0; d < h->D; ++d) s += acc[d];
      grad[0] = 0.5 * s / h->hyp.ell[0];
    } else {
      for (int d = 0; d < h->D; ++d) grad[d] = 0.5 * acc[d] / h->hyp.ell[d];
    }
    grad[P - 2] = 0.5 * (2.0 / h->hyp.sf) * acc[h->DP];
    grad[P - 1] = 0.5 * (2.0 * h->hyp.sn) * acc[h->DP + 1];
    h->factor_ready = true;  // W and alpha are valid for this theta
    return 0;
  }
  stage_end(h, GPB_STAGE_TOTAL);
  double host0 = 0.0;
  CK(cudaMemcpyAsync(&host0, h->dscal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = check_info(h))) return rc;
  stages_collect(h);
  *nll = host0;
  return 0;
}

// Largest-k eigenpairs of Ky(theta) of the resident training set -- scipy.sparse.linalg.eigsh(Ky, neig, tol=clip_eig) inside
// the loop of gp_functions.negative_log_likelihood (gp_functions.py:273-275).  Repeated calls with the same (kind, theta)
// continue the same Lanczos run.
int gpb_eigsh(gpb_handle_t h, int kind, const double* theta, int n_ell, int k, int k_hint, double* w_out, int* steps_out) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!theta || !w_out) return fail(h, -1, "gpb_eigsh: null argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  if ((rc = set_hyp(h, kind, theta, n_ell))) return rc;
  const int n = (int)h->ws.N;
  if ((rc = eig_reserve(h, n))) return rc;
  H::Eig& e = h->eig;
  std::vector<double> key = {(double)kind, (double)n_ell};
  key.insert(key.end(), theta, theta + n_ell + 2);
  if (!(e.matrix_ready && e.from_train && e.key == key)) {
    e.matrix_ready = false;
    if ((rc = dense_kernel_device(h, h->hyp, h->dX, n, h->dX, n, e.K, nullptr, false))) return rc;
    if ((rc = eig_begin(h))) return rc;
    e.key = key;
    e.from_train = true;
    e.matrix_ready = true;
  }
  if ((rc = eig_topk(h, k, steps_out, k_hint))) {
    e.matrix_ready = false;
    return rc;
  }
  std::memcpy(w_out, e.w_host.data(), e.k * sizeof(double));
  return 0;
}

// The same for a caller-provided symmetric matrix A (n x n) -- eigsh(K, neig, tol=tol) in gp_functions.invert
// (gp_functions.py:351,359).  reuse != 0: A is the matrix of the previous call (continue its Lanczos run, A is not read).
// Qt_out (k x n, row i = eigenvector of w_out[i]) may be NULL.
int gpb_eigsh_matrix(gpb_handle_t h, const double* A, long long n, int k, int reuse, double* w_out, double* Qt_out,
                     int* steps_out) {
  if (!h || !w_out || n < 1 || n > 2000000000LL || (!A && !reuse)) return fail(h, -1, "gpb_eigsh_matrix: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  H::Eig& e = h->eig;
  int rc;
  if (!(reuse && e.matrix_ready && !e.from_train && e.n == (int)n)) {
    if (!A) return fail(h, -3, "gpb_eigsh_matrix: nothing to reuse");
    if ((rc = eig_reserve(h, (int)n))) return rc;
    e.matrix_ready = false;
    CK(cudaMemcpyAsync(e.K, A, (size_t)n * n * sizeof(double), cudaMemcpyDefault, h->stream));
    if ((rc = eig_begin(h))) return rc;
    e.from_train = false;
    e.key.clear();
    e.matrix_ready = true;
  }
  if ((rc = eig_topk(h, k, steps_out))) {
    e.matrix_ready = false;
    return rc;
  }
  std::memcpy(w_out, e.w_host.data(), e.k * sizeof(double));
  if (Qt_out) {
    if ((rc = eig_vectors(h))) return rc;
    CK(cudaMemcpyAsync(Qt_out, e.Qt, (size_t)e.k * n * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

// K^-1 = Q diag(1/w) Q^T from the resident Ritz vectors and the caller's (adjusted) eigenvalues -- the last line of
// gp_functions.invert (gp_functions.py:366-369).  w: host, k entries; Kinv: (n, n).
int gpb_eig_inverse(gpb_handle_t h, const double* w, double* Kinv) {
  if (!h || !w || !Kinv) return fail(h, -1, "gpb_eig_inverse: null argument");
  H::Eig& e = h->eig;
  if (!e.matrix_ready || e.k < 1) return fail(h, -3, "gpb_eig_inverse: no resident eigenpairs (call gpb_eigsh_matrix first)");
  CK(cudaSetDevice(h->device));
  const int n = e.n;
  int rc;
  if ((rc = eig_vectors(h))) return rc;
  CK(cudaMemcpyAsync(e.wsel, w, e.k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const bool dev = is_device_ptr(Kinv);
  if (!dev && (rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, (size_t)n * n))) return rc;
  double* out = dev ? Kinv : h->dstage2;
  dim3 grid((unsigned)((n + EIG_BLOCK - 1) / EIG_BLOCK), (unsigned)std::min(n, 65535));
  eig_outer_kernel<<<grid, EIG_BLOCK, 0, h->stream>>>(e.Qt, n, n, e.k, e.wsel, out);
  h->launches++;
  CK(cudaGetLastError());
  if (!dev) CK(cudaMemcpyAsync(Kinv, out, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// NLL (and gradient) on the truncated eigen-decomposition exit of gp_functions.negative_log_likelihood
// (gp_functions.py:283-300) from the eigenpairs left resident by gpb_eigsh:
//   w <- max(w, 1e-10), alpha = Q diag(1/w) Q^T y, nll = 1/2 (y^T alpha + sum log w + neig_term log 2 pi);
//   gradient = 1/2 tr((invert(Ky) - alpha alpha^T) dKy) with invert(Ky) through the Cholesky factorisation, retried once
//   with 1e-6 on the diagonal (invert, gp_functions.py:340-348); a second failure returns the pivot index like gpb_nll.
int gpb_nll_eig(gpb_handle_t h, int neig_term, int want_grad, double* nll, double* grad) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!nll || (want_grad && !grad)) return fail(h, -1, "gpb_nll_eig: null argument");
  H::Eig& e = h->eig;
  if (!e.matrix_ready || !e.from_train || e.k < 1 || e.n != (int)h->ws.N)
    return fail(h, -3, "gpb_nll_eig: no resident eigenpairs of the training covariance (call gpb_eigsh first)");
  if (h->n_out != 1) return fail(h, -1, "gpb_nll_eig: the eigen-decomposition branch handles one output column (nll.item())");
  CK(cudaSetDevice(h->device));
  Workspace& w = h->ws;
  const int n = e.n, k = e.k;
  h->factor_ready = false;
  stages_reset(h);
  if ((rc = eig_vectors(h))) return rc;
  CK(cudaMemcpyAsync(e.wsel, e.w_host.data(), k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  eig_project_kernel<<<(unsigned)k, EIG_BLOCK, 0, h->stream>>>(e.Qt, n, n, h->dy, e.wsel, 1e-10, e.qy, e.t);
  eig_alpha_kernel<<<(unsigned)((n + EIG_BLOCK - 1) / EIG_BLOCK), EIG_BLOCK, 0, h->stream>>>(e.Qt, n, n, k, e.t, e.avec);
  eig_nll_kernel<<<1, EIG_BLOCK, 0, h->stream>>>(e.wsel, e.qy, e.t, k, 1e-10, (double)neig_term, h->dscal);
  h->launches += 3;
  CK(cudaGetLastError());
  double host[64];
  if (!want_grad) {
    CK(cudaMemcpyAsync(host, h->dscal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *nll = host[0];
    return 0;
  }
  CK(cudaMemcpyAsync(host, h->dscal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *nll = host[0];
  if ((rc = ensure_inverse_buffers(h, w))) return rc;
  const double sn_true = h->hyp.sn;
  const int P = h->hyp.n_ell + 2, width = h->DP + 2;
  // invert(Ky): Cholesky, then Cholesky of Ky + 1e-6 I (gp_functions.py:340-343).  Where the reference would now raise
  // inside np.linalg.cholesky, the drop-in keeps the optimiser alive with the relative jitter ladder of the "jitter"
  // policy (1e-8 .. 1e-2 sigma_f^2); only if that fails too is the pivot index returned.
  const double sf2 = h->hyp.sf * h->hyp.sf;
  const double jitter[6] = {0.0, 1e-6, 1e-8 * sf2, 1e-6 * sf2, 1e-4 * sf2, 1e-2 * sf2};
  double used = 0.0;
  int info = 0;
  const bool stash = ensure_stash(w, h->hyp.kind != KIND_RBF);   // the radial factors do not depend on the jittered sigma_n
  for (int attempt = 0; attempt < 6; ++attempt) {
    if (attempt >= 2 && jitter[attempt] <= used) continue;
    used = jitter[attempt];
    h->hyp.sn = std::sqrt(sn_true * sn_true + used);
    rc = assemble_and_factor(h, true, stash);
    if (!rc) rc = triangular_inverse(h, w, true);
    if (!rc) rc = run_plan(h, w, w.syrk, false);
    h->hyp.sn = sn_true;
    if (rc) return rc;
    info = check_info(h);
    if (info <= 0) break;
  }
  if (info) return info;
  h->eig.grad_jitter = used;
  CK(cudaMemsetAsync(w.alpha, 0, (size_t)w.Np * sizeof(double), h->stream));
  CK(cudaMemcpyAsync(w.alpha, e.avec, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  rc = (h->hyp.kind == KIND_RBF) ? contract_and_reduce<KIND_RBF>(h, w, stash) : contract_and_reduce<KIND_MATERN52>(h, w, stash);
  if (rc) return rc;
  CK(cudaMemcpyAsync(host, h->dscal, (1 + width) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const double* acc = host + 1;
  if (h->hyp.n_ell == 1) {
    double sum = 0.0;
    for (int d = 0; d < h->D; ++d) sum += acc[d];
    grad[0] = 0.5 * sum / h->hyp.ell[0];
  } else {
    for (int d = 0; d < h->D; ++d) grad[d] = 0.5 * acc[d] / h->hyp.ell[d];
  }
  grad[P - 2] = 0.5 * (2.0 / h->hyp.sf) * acc[h->DP];
  grad[P - 1] = 0.5 * (2.0 * sn_true) * acc[h->DP + 1];
  return 0;
}

int gpb_factorize(gpb_handle_t h, int kind, const double* theta, int n_ell) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!theta) return fail(h, -1, "gpb_factorize: null argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  if ((rc = set_hyp(h, kind, theta, n_ell))) return rc;
  Workspace& w = h->ws;
  h->factor_ready = false;
  stages_reset(h);
  stage_begin(h, GPB_STAGE_TOTAL);
  if ((rc = ensure_inverse_buffers(h, w))) return rc;
  if ((rc = assemble_and_factor(h, true))) return rc;
  stage_begin(h, GPB_STAGE_TRTRI);
  if ((rc = triangular_inverse(h, w, true))) return rc;
  stage_end(h, GPB_STAGE_TRTRI);
  stage_begin(h, GPB_STAGE_GRAD);
  if ((rc = compute_alpha(h, w, h->n_out))) return rc;
  stage_end(h, GPB_STAGE_GRAD);
  stage_end(h, GPB_STAGE_TOTAL);
  if ((rc = check_info(h))) return rc;
  stages_collect(h);
  h->factor_ready = true;
  return 0;
}

int gpb_predict(gpb_handle_t h, const double* Xc, long long M, int add_data_variance, double* mean, double* var) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!h->factor_ready) return fail(h, -3, "gpb_predict: call gpb_factorize (or gpb_nll with gradient) first");
  if (!Xc || M < 0) return fail(h, -1, "gpb_predict: bad argument");
  if (M == 0) return 0;
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->ws;
  const int cfg = g_cfg_update;
  int bm, bn;
  cfg_tile(cfg, &bm, &bn);
  // chunk of candidates: K*^T chunk (Mc x Np doubles) bounded to ~2 GiB
  long long Mc = ((1LL << 28) / w.Np) / 128 * 128;
  Mc = std::max<long long>(128, std::min<long long>(Mc, 65536));
  Mc = std::min(Mc, round_up(M, 128));
  if (Mc != h->pred_Mc || w.Nb != h->pred_Nb) {
    cudaFree(h->dXc);
    cudaFree(h->dKsT);
    cudaFree(h->dVpart);
    cudaFree(h->dmean);
    cudaFree(h->dvar);
    h->dXc = h->dKsT = h->dVpart = h->dmean = h->dvar = nullptr;
    h->pred_Mc = h->pred_Nb = 0;
    CK(cudaMalloc(&h->dXc, (size_t)Mc * MAX_D * sizeof(double)));  // sized for any D: the chunk buffers outlive set_train
    CK(cudaMalloc(&h->dKsT, (size_t)Mc * w.Np * sizeof(double)));
    CK(cudaMalloc(&h->dVpart, (size_t)w.Nb * Mc * sizeof(double)));
    CK(cudaMalloc(&h->dmean, (size_t)Mc * 128 * sizeof(double)));
    CK(cudaMalloc(&h->dvar, (size_t)Mc * sizeof(double)));
    // candidate rows beyond the last real one are never written by cross_mean_kernel: keep them finite
    CK(cudaMemsetAsync(h->dKsT, 0, (size_t)Mc * w.Np * sizeof(double), h->stream));
    h->pred.clear();
    {  // launch 0: row tiles of W descending (longest k first), candidate tiles ascending -- the ragged last chunk
       // launches a prefix of every row from this list
      TaskSink s(h->pred);
      for (int mt = (int)w.Nb - 1; mt >= 0; --mt)
        for (long long ct = 0; ct < Mc / 128; ++ct) s.block(cfg, mt * 128, 0, (int)ct * 128, 0, 8 * (mt + 1), 0, (int)ct * 128, mt);
      s.finish(cfg, EPI_COLSQ, M_W, M_KST, M_PART, M_NONE, 1.0, 0.0, false);
    }
    {  // launch 1 (full chunks): the same tiles super-tile by super-tile (see build_syrk_plan).  In the plain order every
       // candidate panel is streamed from HBM once per row tile of W (ncu: 70.9 GB per 32768 candidates at N=8192); a
       // g_pred_super_m x g_pred_super_c group of resident CTAs shares its W and candidate slabs through L2 instead.
      TaskSink s(h->pred);
      const int Gm = std::max(1, g_pred_super_m), Gc = std::max(1, g_pred_super_c), nct = (int)(Mc / 128);
      for (int M1 = (int)w.Nb; M1 > 0; M1 -= Gm)
        for (int C0 = 0; C0 < nct; C0 += Gc)
          for (int mt = M1 - 1; mt >= std::max(0, M1 - Gm); --mt)
            for (int ct = C0; ct < std::min(nct, C0 + Gc); ++ct)
              s.block(cfg, mt * 128, 0, ct * 128, 0, 8 * (mt + 1), 0, ct * 128, mt);
      s.finish(cfg, EPI_COLSQ, M_W, M_KST, M_PART, M_NONE, 1.0, 0.0, false);
    }
    if ((rc = upload_plan(h, h->pred))) return rc;
    h->pred_Mc = Mc;
    h->pred_Nb = w.Nb;
  }
  set_mat(w, M_KST, h->dKsT, Mc, w.Np, w.Np);
  set_mat(w, M_PART, h->dVpart, w.Nb, Mc, Mc);
  const double sf2 = h->hyp.sf * h->hyp.sf;
  const double add = add_data_variance ? h->hyp.sn * h->hyp.sn : 0.0;
  stages_reset(h);
  stage_begin(h, GPB_STAGE_TOTAL);
  double cross_ms = 0.0, var_ms = 0.0;
  cudaEvent_t e0 = h->evx[0], e1 = h->evx[1], e2 = h->evx[2];
  for (long long c0 = 0; c0 < M; c0 += Mc) {
    const long long mc = std::min(Mc, M - c0);
    CK(cudaMemcpyAsync(h->dXc, Xc + c0 * h->D, (size_t)mc * h->D * sizeof(double), cudaMemcpyDefault, h->stream));
    cudaEventRecord(e0, h->stream);
    rc = (h->hyp.kind == KIND_RBF)
             ? cross_dispatch<KIND_RBF>(h, w, h->dXc, mc, var ? h->dKsT : nullptr, w.Np, mean ? h->dmean : nullptr)
             : cross_dispatch<KIND_MATERN52>(h, w, h->dXc, mc, var ? h->dKsT : nullptr, w.Np, mean ? h->dmean : nullptr);
    if (rc) return rc;
    cudaEventRecord(e1, h->stream);
    if (var) {
      // only the candidate tiles that hold real rows
      const long long ctiles = (mc + 127) / 128;
      const Launch& L = h->pred.launches[0];
      const int per_block = (cfg == CFG_128x128) ? 1 : 2;
      GemmEpilogue ep;
      ep.C = h->dVpart;
      ep.ldc = Mc;
      ep.Ct = nullptr;
      ep.ldct = 0;
      ep.alpha = 1.0;
      ep.beta = 0.0;
      if (ctiles == Mc / 128) {
        const Launch& LS = h->pred.launches[g_pred_super_m > 0 ? 1 : 0];
        if ((rc = launch_gemm(h, w, cfg, EPI_COLSQ, M_W, M_KST, h->pred.d_tasks + LS.task_off, LS.ntasks, ep))) return rc;
      } else {
        // ragged last chunk: tasks are ordered (mt desc, ct asc); launch the first ctiles of every mt row
        for (int r = 0; r < (int)w.Nb; ++r) {
          const GemmTask* base = h->pred.d_tasks + L.task_off + (size_t)r * (Mc / 128) * per_block;
          if ((rc = launch_gemm(h, w, cfg, EPI_COLSQ, M_W, M_KST, base, (int)ctiles * per_block, ep))) return rc;
        }
      }
      finish_variance_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, h->stream>>>(h->dVpart, Mc, (int)w.Nb, (int)mc, sf2,
                                                                                add, add_data_variance == 2, h->dvar);
      h->launches++;
      CK(cudaGetLastError());
    }
    cudaEventRecord(e2, h->stream);
    if (mean)
      CK(cudaMemcpyAsync(mean + c0 * h->n_out, h->dmean, (size_t)mc * h->n_out * sizeof(double), cudaMemcpyDefault,
                         h->stream));
    if (var) CK(cudaMemcpyAsync(var + c0, h->dvar, (size_t)mc * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&b, e1, e2);
    cross_ms += a;
    var_ms += b;
  }
  stage_end(h, GPB_STAGE_TOTAL);
  CK(cudaStreamSynchronize(h->stream));
  stages_collect(h);
  h->stage_ms[GPB_STAGE_CROSS] = cross_ms;
  h->stage_ms[GPB_STAGE_VAR] = var_ms;
  return 0;
}

int gpb_predict_cov(gpb_handle_t h, const double* Xc, long long M, int noise_on_diagonal, double* mean, double* cov) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!h->factor_ready) return fail(h, -3, "gpb_predict_cov: call gpb_factorize first");
  if (!Xc || !cov || M < 1) return fail(h, -1, "gpb_predict_cov: bad argument");
  if (M > 16384) return fail(h, -1, "gpb_predict_cov: at most 16384 test points (the covariance is M x M)");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->ws;
  const long long Np = w.Np, Mp = round_up(M, 128);
  const int D = h->D, cfg = cfg_for_blocks(g_cfg_update, (int)((Mp / 128) * w.Nb));
  // scratch: candidates | K*^T (Mp x Np) | Vt = K*^T W^T (Mp x Np) | covariance (Mp x Mp) | mean
  const size_t need = (size_t)M * D + 16 + 2 * (size_t)Mp * Np + (size_t)Mp * Mp + (size_t)Mp * 128;
  if ((rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, need))) return rc;
  double* dXc = h->dstage2;
  double* kst = dXc + (((size_t)M * D + 15) & ~(size_t)15);   // keep the matrices 128-byte aligned for TMA
  double* vt = kst + (size_t)Mp * Np;
  double* cv = vt + (size_t)Mp * Np;
  double* dmean = cv + (size_t)Mp * Mp;
  CK(cudaMemcpyAsync(dXc, Xc, (size_t)M * D * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemsetAsync(kst, 0, (size_t)Mp * Np * sizeof(double), h->stream));
  rc = (h->hyp.kind == KIND_RBF) ? cross_dispatch<KIND_RBF>(h, w, dXc, M, kst, Np, mean ? dmean : nullptr)
                                 : cross_dispatch<KIND_MATERN52>(h, w, dXc, M, kst, Np, mean ? dmean : nullptr);
  if (rc) return rc;
  KernelHyp hp = h->hyp;
  if (!noise_on_diagonal) hp.sn = 0.0;
  {
    const long long tot = Mp * Mp;
    if (hp.kind == KIND_RBF)
      cov_prior_kernel<KIND_RBF><<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(dXc, (int)M, hp, cv, Mp, (int)Mp);
    else
      cov_prior_kernel<KIND_MATERN52><<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(dXc, (int)M, hp, cv, Mp, (int)Mp);
    h->launches++;
    CK(cudaGetLastError());
  }
  set_mat(w, M_KST, kst, Mp, Np, Np);
  set_mat(w, M_DBG_A, vt, Mp, Np, Np);
  Plan p;
  PlanGuard guard(p);
  {
    // Vt[c][i] = sum_{k <= i} K*^T[c][k] W[i][k]
    TaskSink s(p);
    for (int ib = (int)w.Nb - 1; ib >= 0; --ib)
      for (long long cb = 0; cb < Mp / 128; ++cb) s.block(cfg, (int)cb * 128, 0, ib * 128, 0, 8 * (ib + 1), (int)cb * 128, ib * 128);
    s.finish(cfg, EPI_STORE, M_KST, M_W, M_DBG_A, M_NONE, 1.0, 0.0, false);
  }
  const size_t n1 = p.tasks.size();
  {
    // C -= Vt Vt^T
    TaskSink s(p);
    for (long long i = 0; i < Mp / 128; ++i)
      for (long long j = 0; j < Mp / 128; ++j) s.block(cfg, (int)i * 128, 0, (int)j * 128, 0, (int)(Np / 16), (int)i * 128, (int)j * 128);
    s.finish(cfg, EPI_STORE, M_DBG_A, M_DBG_A, M_DBG_C, M_NONE, -1.0, 1.0, false);
  }
  if ((rc = upload_plan(h, p))) return rc;
  GemmEpilogue ep;
  ep.C = vt;
  ep.ldc = Np;
  ep.Ct = nullptr;
  ep.ldct = 0;
  ep.alpha = 1.0;
  ep.beta = 0.0;
  if ((rc = launch_gemm(h, w, cfg, EPI_STORE, M_KST, M_W, p.d_tasks, (int)n1, ep))) return rc;
  ep.C = cv;
  ep.ldc = Mp;
  ep.alpha = -1.0;
  ep.beta = 1.0;
  if ((rc = launch_gemm(h, w, cfg, EPI_STORE, M_DBG_A, M_DBG_A, p.d_tasks + n1, (int)(p.tasks.size() - n1), ep))) return rc;
  const bool cov_dev = is_device_ptr(cov);
  double* dout = cov;
  if (!cov_dev) {
    if ((rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)M * M))) return rc;
    dout = h->dstage;
  }
  {
    const long long tot = M * M;
    cov_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(cv, Mp, (int)M, dout);
    h->launches++;
    CK(cudaGetLastError());
  }
  if (!cov_dev) CK(cudaMemcpyAsync(cov, dout, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (mean) CK(cudaMemcpyAsync(mean, dmean, (size_t)M * h->n_out * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_argmax(gpb_handle_t h, const double* v, long long M, long long* idx, double* val) {
  if (!h || !v || M < 1 || !idx) return fail(h, -1, "gpb_argmax: bad argument");
  // the plain variance acquisition without a mask is exactly np.argmax
  return gpb_acquire(h, ACQ_VARIANCE, nullptr, v, nullptr, M, 1, nullptr, 0, nullptr, 0, nullptr, idx, val, nullptr);
}

int gpb_kernel_matrix(gpb_handle_t h, int kind, const double* Xa, long long na, const double* Xb, long long nb, int d,
                      const double* ell, int n_ell, double sigma_f, double sigma_n, double* K, double* dK) {
  if (!h || !Xa || !Xb || !ell || !K || na < 1 || nb < 1 || d < 1 || d > MAX_D)
    return fail(h, -1, "gpb_kernel_matrix: bad argument");
  if (!(n_ell == 1 || n_ell == d)) return fail(h, -1, "gpb_kernel_matrix: n_ell must be 1 or d");
  if (kind != GPB_KERNEL_RBF && kind != GPB_KERNEL_MATERN52) return fail(h, -1, "unknown kernel kind");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  KernelHyp hy{};
  hy.kind = kind;
  hy.D = d;
  hy.n_ell = n_ell;
  for (int i = 0; i < n_ell; ++i) hy.ell[i] = ell[i];
  hy.sf = sigma_f;
  hy.sn = sigma_n;
  const int P = n_ell + 2;
  const size_t nK = (size_t)na * nb, ndK = dK ? nK * P : 0;
  int rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)(na + nb) * d);
  if (rc) return rc;
  const bool k_dev = is_device_ptr(K), dk_dev = dK ? is_device_ptr(dK) : true;
  const size_t need_out = (k_dev ? 0 : nK) + (dk_dev ? 0 : ndK);
  if (need_out && (rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, need_out))) return rc;
  double* dXa = h->dstage;
  double* dXb = h->dstage + (size_t)na * d;
  CK(cudaMemcpyAsync(dXa, Xa, (size_t)na * d * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemcpyAsync(dXb, Xb, (size_t)nb * d * sizeof(double), cudaMemcpyDefault, h->stream));
  double* dKo = k_dev ? K : h->dstage2;
  double* ddK = dK ? (dk_dev ? dK : h->dstage2 + (k_dev ? 0 : nK)) : nullptr;
  stages_reset(h);
  if ((rc = dense_kernel_device(h, hy, dXa, na, dXb, nb, dKo, ddK, true))) return rc;
  if (!k_dev) CK(cudaMemcpyAsync(K, dKo, nK * sizeof(double), cudaMemcpyDefault, h->stream));
  if (dK && !dk_dev) CK(cudaMemcpyAsync(dK, ddK, ndK * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  stages_collect(h);
  return 0;
}

// python_kernels.LinearEmbedding (python_kernels.py:60-114): K (na, nb) and, if dK != NULL, dK (na, nb, d*D + 2).
int gpb_linear_embedding(gpb_handle_t h, const double* Xa, long long na, const double* Xb, long long nb, int D,
                         const double* R, int d, double sigma_f, double sigma_n, double* K, double* dK) {
  if (!h || !Xa || !Xb || !R || !K || na < 1 || nb < 1 || D < 1 || D > MAX_D || d < 1 || d > MAX_D)
    return fail(h, -1, "gpb_linear_embedding: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  const int P = d * D + 2;
  const size_t nK = (size_t)na * nb, ndK = dK ? nK * P : 0;
  int rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)(na + nb) * D + (size_t)d * D);
  if (rc) return rc;
  const bool k_dev = is_device_ptr(K), dk_dev = dK ? is_device_ptr(dK) : true;
  const size_t need_out = (k_dev ? 0 : nK) + (dk_dev ? 0 : ndK);
  if (need_out && (rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, need_out))) return rc;
  double* dXa = h->dstage;
  double* dXb = dXa + (size_t)na * D;
  double* dR = dXb + (size_t)nb * D;
  CK(cudaMemcpyAsync(dXa, Xa, (size_t)na * D * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemcpyAsync(dXb, Xb, (size_t)nb * D * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemcpyAsync(dR, R, (size_t)d * D * sizeof(double), cudaMemcpyDefault, h->stream));
  double* dKo = k_dev ? K : h->dstage2;
  double* ddK = dK ? (dk_dev ? dK : h->dstage2 + (k_dev ? 0 : nK)) : nullptr;
  linear_embedding_kernel<<<(unsigned)((nK + 255) / 256), 256, 0, h->stream>>>(dXa, (int)na, dXb, (int)nb, D, dR, d, sigma_f,
                                                                            sigma_n, dKo, ddK);
  h->launches++;
  CK(cudaGetLastError());
  if (!k_dev) CK(cudaMemcpyAsync(K, dKo, nK * sizeof(double), cudaMemcpyDefault, h->stream));
  if (dK && !dk_dev) CK(cudaMemcpyAsync(dK, ddK, ndK * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// Marginal (hyper-parameter-uncertainty) variance after Osborne 2012 -- gp_functions.marginal_variance_BBQ
// (gp_functions.py:376-454):  out[c] = dm_c^T H^-1 dm_c + pred_var[c],
//   dm_c[p] = dK*/dtheta_p[c,:] . alpha + K*[c,:] . dalpha/dtheta_p,   dalpha/dtheta_p = -K^-1 (dK/dtheta_p alpha).
// Needs the factor of the current hyper-parameters (gpb_factorize).  hess_inv is (P, P) row-major, P = n_ell + 2.
int gpb_marginal_variance(gpb_handle_t h, const double* Xc, long long M, const double* hess_inv, const double* pred_var,
                          double* out) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!h->factor_ready) return fail(h, -3, "gpb_marginal_variance: call gpb_factorize first");
  if (!Xc || !hess_inv || !out || M < 1) return fail(h, -1, "gpb_marginal_variance: bad argument");
  if (h->n_out != 1) return fail(h, -1, "gpb_marginal_variance: single-output models only");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->ws;
  const long long N = w.N, Np = w.Np;
  const int P = h->hyp.n_ell + 2, D = h->D;
  // candidate chunk: (Mc x N x (P + 1)) doubles bounded to ~1 GiB
  long long Mc = std::max<long long>(1, (1LL << 27) / (N * (P + 1)));
  Mc = std::min<long long>(Mc, M);
  const size_t n_train_dk = (size_t)N * N * (P + 1), n_cand_dk = (size_t)Mc * N * (P + 1);
  const size_t vec = (size_t)3 * P * Np + (size_t)P * P + (size_t)Mc * (P + 2 + D);
  if ((rc = ensure_cap(h, &h->dbbq, &h->dbbq_cap, std::max(n_train_dk, n_cand_dk) + vec))) return rc;
  double* big = h->dbbq;                       // K | dK of the current block
  double* v = big + std::max(n_train_dk, n_cand_dk);   // [P][Np]  dK_p alpha
  double* t = v + (size_t)P * Np;                      // [P][Np]  W v_p
  double* da = t + (size_t)P * Np;                     // [P][Np]  K^-1 v_p  ( = -dalpha/dtheta_p )
  double* dH = da + (size_t)P * Np;                    // [P][P]
  double* dm = dH + (size_t)P * P;                     // [Mc][P]
  double* dpv = dm + (size_t)Mc * P;                   // [Mc]
  double* dout = dpv + Mc;                             // [Mc]
  double* dXc = dout + Mc;                             // [Mc][D]
  CK(cudaMemcpyAsync(dH, hess_inv, (size_t)P * P * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemsetAsync(v, 0, (size_t)3 * P * Np * sizeof(double), h->stream));
  // training block: dK(X, X) with the noise plane, then v_p = dK_p alpha
  if ((rc = dense_kernel_device(h, h->hyp, h->dX, N, h->dX, N, big, big + (size_t)N * N, false))) return rc;
  rowwise_contract_kernel<<<(unsigned)((N + 7) / 8), 256, 0, h->stream>>>(big + (size_t)N * N, nullptr, (int)N, (int)N, P,
                                                                          w.alpha, nullptr, Np, 1.0, v, 1, Np);
  // K^-1 v_p = U (W v_p)
  trigemv_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, h->stream>>>(w.Wbuf, Np, Np, true, v, 1, Np, P, t, 1, Np);
  trigemv_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, h->stream>>>(w.Ubuf, Np, Np, false, t, 1, Np, P, da, 1, Np);
  h->launches += 3;
  CK(cudaGetLastError());
  KernelHyp hs = h->hyp;
  hs.sn = 0.0;   // K* and dK* are evaluated without the noise term (gp_functions.py:427-429)
  for (long long c0 = 0; c0 < M; c0 += Mc) {
    const long long mc = std::min(Mc, M - c0);
    CK(cudaMemcpyAsync(dXc, Xc + c0 * D, (size_t)mc * D * sizeof(double), cudaMemcpyDefault, h->stream));
    if (pred_var) CK(cudaMemcpyAsync(dpv, pred_var + c0, (size_t)mc * sizeof(double), cudaMemcpyDefault, h->stream));
    else CK(cudaMemsetAsync(dpv, 0, (size_t)mc * sizeof(double), h->stream));
    if ((rc = dense_kernel_device(h, hs, dXc, mc, h->dX, N, big, big + (size_t)mc * N, false))) return rc;
    // dm[c][p] = sum_i dK*[c][i][p] alpha_i - K*[c][i] (K^-1 v_p)_i
    rowwise_contract_kernel<<<(unsigned)((mc + 7) / 8), 256, 0, h->stream>>>(big + (size_t)mc * N, big, (int)mc, (int)N, P,
                                                                           w.alpha, da, Np, -1.0, dm, P, 1);
    quadform_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, h->stream>>>(dm, (int)mc, P, dH, dpv, dout);
    h->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + c0, dout, (size_t)mc * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int gpb_append_train(gpb_handle_t h, const double* Xnew, const double* ynew, long long m) {
  int rc = need_train(h);
  if (rc) return rc;
  if (!h->factor_ready) return fail(h, -3, "gpb_append_train: call gpb_factorize first");
  if (!Xnew || !ynew || m < 1) return fail(h, -1, "gpb_append_train: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  const int D = h->D, n_out = h->n_out;
  for (long long p = 0; p < m; ++p) {
    Workspace& w = h->ws;
    const long long N = w.N, Np = w.Np;
    if (N == Np) {
      // The padded block is full: grow by re-uploading [X; Xnew[p:]] and refactorising at the current theta.
      const long long rem = m - p, tot = N + rem;
      double *tx = nullptr, *ty = nullptr;
      if (cudaMalloc(&tx, (size_t)tot * D * sizeof(double)) != cudaSuccess ||
          cudaMalloc(&ty, (size_t)tot * n_out * sizeof(double)) != cudaSuccess) {
        cudaFree(tx);
        return fail(h, -2, "gpb_append_train: out of device memory while growing the training set");
      }
      cudaError_t e1 = cudaMemcpyAsync(tx, h->dX, (size_t)N * D * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
      cudaError_t e2 = cudaMemcpyAsync(ty, h->dy, (size_t)N * n_out * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
      cudaError_t e3 = cudaMemcpyAsync(tx + N * D, Xnew + p * D, (size_t)rem * D * sizeof(double), cudaMemcpyDefault, h->stream);
      cudaError_t e4 = cudaMemcpyAsync(ty + N * n_out, ynew + p * n_out, (size_t)rem * n_out * sizeof(double), cudaMemcpyDefault, h->stream);
      cudaError_t e5 = cudaStreamSynchronize(h->stream);
      double theta[MAX_D + 2];
      const int n_ell = h->hyp.n_ell, kind = h->hyp.kind;
      for (int d = 0; d < n_ell; ++d) theta[d] = h->hyp.ell[d];
      theta[n_ell] = h->hyp.sf;
      theta[n_ell + 1] = h->hyp.sn;
      rc = (e1 || e2 || e3 || e4 || e5) ? fail(h, -2, "gpb_append_train: staging copy failed") : 0;
      if (!rc) rc = gpb_set_train(h, tx, ty, tot, D, n_out);
      if (!rc) rc = gpb_factorize(h, kind, theta, n_ell);
      cudaFree(tx);
      cudaFree(ty);
      return rc;
    }
    // scratch: [32 x Np] K* rows (cross_mean_kernel writes whole 32-row groups) | l | U l | scalars | new y
    if ((rc = ensure_cap(h, &h->dapp, &h->dapp_cap, (size_t)34 * Np + 320))) return rc;
    double* kst = h->dapp;
    double* lvec = kst + (size_t)32 * Np;
    double* tvec = lvec + Np;
    double* scal = tvec + Np;        // [0] = 1/d, [1 + o] = z_o[N]
    double* dynew = scal + 129;      // n_out values
    double* dxnew = dynew + 128;     // D values
    CK(cudaMemcpyAsync(dxnew, Xnew + p * D, (size_t)D * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(dynew, ynew + p * n_out, (size_t)n_out * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
    rc = (h->hyp.kind == KIND_RBF) ? cross_dispatch<KIND_RBF>(h, w, dxnew, 1, kst, Np, nullptr)
                                   : cross_dispatch<KIND_MATERN52>(h, w, dxnew, 1, kst, Np, nullptr);
    if (rc) return rc;
    // l = W[0:N, 0:N] k
    trigemv_kernel<<<(unsigned)((N + 7) / 8), 256, 0, h->stream>>>(w.Wbuf, Np, N, true, kst, 1, 0, 1, lvec, 1, 0);
    const double kss = h->hyp.sf * h->hyp.sf + h->hyp.sn * h->hyp.sn;
    append_pivot_kernel<<<1, 1024, 0, h->stream>>>(lvec, N, kss, dynew, n_out, w.Lbuf + Np * Np, Np,
                                                  w.logdet + N / 128, scal, h->dinfo);
    h->launches += 2;
    CK(cudaGetLastError());
    if ((rc = check_info(h))) return rc;   // not SPD with the new point: nothing has been modified
    // t = U[0:N, 0:N] l;  new row of W / column of U;  alpha update
    trigemv_kernel<<<(unsigned)((N + 7) / 8), 256, 0, h->stream>>>(w.Ubuf, Np, N, false, lvec, 1, 0, 1, tvec, 1, 0);
    append_update_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, h->stream>>>(tvec, N, scal, n_out, w.Wbuf, w.Ubuf, Np,
                                                                                w.alpha, Np);
    // the raw point joins the resident training set; refresh the scaled copies
    CK(cudaMemcpyAsync(h->dX + N * D, dxnew, (size_t)D * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dy + N * n_out, dynew, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    w.N = N + 1;
    {
      const long long tot = Np * h->DP;
      prescale_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->dX, (int)w.N, h->hyp, h->DP, h->dXs,
                                                                          h->dXsT, Np, (int)Np);
    }
    h->launches += 3;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_acq_stats(gpb_handle_t h, const double* v, long long M, int ncol, int transform, double* out) {
  if (!h || !v || !out || M < 1 || ncol < 1 || ncol > ACQ_MAX_OUT || (transform != 0 && transform != 1))
    return fail(h, -1, "gpb_acq_stats: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  const size_t tot = (size_t)M * ncol;
  const int nblk = acq_grid(M);
  int rc = ensure_cap(h, &h->dacq, &h->dacq_cap, tot + (size_t)2 * nblk * ncol + 2 * ACQ_MAX_OUT + 64);
  if (rc) return rc;
  double* cur = h->dacq;
  const double* dv = nullptr;
  if ((rc = acq_stage_in(h, v, tot, &cur, &dv))) return rc;
  double* part = cur;
  double* stats = part + (size_t)2 * nblk * ncol;
  if ((rc = acq_minmax(h, dv, M, ncol, transform, part, stats))) return rc;
  CK(cudaMemcpyAsync(out, stats, (size_t)2 * ncol * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_acq_prepare(gpb_handle_t h, int kind, const double* mean, long long M, int n_out, const double* params,
                    int n_params, double* out, const double* stats_in, double* stats_out) {
  if (!h || !mean || !out || M < 1 || n_out < 1 || n_out > ACQ_MAX_OUT) return fail(h, -1, "gpb_acq_prepare: bad argument");
  if (kind != ACQ_WEIGHTED && kind != ACQ_EI) return fail(h, -1, "gpb_acq_prepare: kind has no prepare step");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  AcqArgs a;
  int rc = 0;
  if (kind == ACQ_EI) {
    if (n_params == 0) return fail(h, -1, "expected improvement: params = [xi, find_min, ymax per output]");
    if ((rc = fill_acq_args(h, a, kind, M, n_out, params, n_params))) return rc;
  } else {   // normalize(mu) has no parameters
    memset(&a, 0, sizeof(a));
    a.kind = kind;
    a.M = M;
    a.n_out = n_out;
  }
  const size_t tot = (size_t)M * n_out;
  const int nblk = acq_grid(M);
  const bool out_dev = is_device_ptr(out);
  if ((rc = ensure_cap(h, &h->dacq, &h->dacq_cap, 2 * tot + (size_t)2 * nblk * n_out + 2 * ACQ_MAX_OUT + 64))) return rc;
  double* cur = h->dacq;
  const double* dmean = nullptr;
  if ((rc = acq_stage_in(h, mean, tot, &cur, &dmean))) return rc;
  double* dout = out_dev ? out : cur;
  if (!out_dev) cur += tot;
  double* part = cur;
  double* stats = part + (size_t)2 * nblk * n_out;
  const int grid = acq_grid((long long)tot);
  acq_prepare_kernel<<<grid, ACQ_THREADS, 0, h->stream>>>(0, a, dmean, stats, dout);
  h->launches++;
  if (!stats_in || stats_out) {
    if ((rc = acq_minmax(h, dout, M, n_out, 0, part, stats))) return rc;
    if (stats_out) CK(cudaMemcpyAsync(stats_out, stats, (size_t)2 * n_out * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  // statistics supplied by the caller (the global ones when the candidate set is sharded over ranks) win
  if (stats_in) CK(cudaMemcpyAsync(stats, stats_in, (size_t)2 * n_out * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  acq_prepare_kernel<<<grid, ACQ_THREADS, 0, h->stream>>>(1, a, dmean, stats, dout);
  h->launches++;
  CK(cudaGetLastError());
  if (!out_dev) CK(cudaMemcpyAsync(out, dout, tot * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_acquire(gpb_handle_t h, int kind, const double* term, const double* var, const double* extra, long long M,
                int n_out, const double* params, int n_params, const long long* visited, int n_visited, double* loss_out,
                long long* idx, double* val, const double* stats_in) {
  if (!h || !var || !idx || M < 1 || n_out < 1 || n_out > ACQ_MAX_OUT || n_visited < 0 || (n_visited > 0 && !visited))
    return fail(h, -1, "gpb_acquire: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  AcqArgs a;
  int rc = fill_acq_args(h, a, kind, M, n_out, params, n_params);
  if (rc) return rc;
  a.n_visited = n_visited;
  const bool need_term = kind == ACQ_WEIGHTED || kind == ACQ_POI || kind == ACQ_EI || kind == ACQ_EI2;
  const bool need_extra = kind == ACQ_POI || kind == ACQ_DISTANCE_PENALTY;
  if ((need_term && !term) || (need_extra && !extra)) return fail(h, -1, "gpb_acquire: missing input array for this kind");
  const size_t n_term = need_term ? (size_t)M * n_out : 0;
  const size_t n_extra = need_extra ? (size_t)M * (kind == ACQ_POI ? n_out : a.d) : 0;
  const int nblk = acq_grid(M);
  const bool loss_dev = loss_out && is_device_ptr(loss_out);
  const size_t cap = n_term + n_extra + (size_t)2 * M + (size_t)4 * nblk + n_visited + 64;
  if ((rc = ensure_cap(h, &h->dacq, &h->dacq_cap, cap))) return rc;
  double* cur = h->dacq;
  const double *dterm = nullptr, *dvar = nullptr, *dextra = nullptr;
  if ((rc = acq_stage_in(h, need_term ? term : nullptr, n_term, &cur, &dterm))) return rc;
  if ((rc = acq_stage_in(h, var, (size_t)M, &cur, &dvar))) return rc;
  if ((rc = acq_stage_in(h, need_extra ? extra : nullptr, n_extra, &cur, &dextra))) return rc;
  double* dloss = nullptr;
  if (loss_out) {
    dloss = loss_dev ? loss_out : cur;
    if (!loss_dev) cur += M;
  }
  double* part = cur;                 // min/max partials, then arg-max partial values
  long long* pidx = reinterpret_cast<long long*>(part + (size_t)2 * nblk);
  double* stats = part + (size_t)3 * nblk;       // [0..1] min/max, [2] winner value, [3] winner index
  long long* dvis = reinterpret_cast<long long*>(stats + 8);
  if (n_visited > 0) {
    for (int k = 0; k < n_visited; ++k)
      if (visited[k] < 0 || visited[k] >= M) return fail(h, -1, "gpb_acquire: visited index out of range");
    CK(cudaMemcpyAsync(dvis, visited, (size_t)n_visited * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  }
  if (kind == ACQ_DISTANCE_PENALTY || kind == ACQ_WEIGHTED || kind == ACQ_EI) {
    if (stats_in) CK(cudaMemcpyAsync(stats, stats_in, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else if ((rc = acq_minmax(h, dvar, M, 1, kind == ACQ_EI ? 1 : 0, part, stats))) return rc;
  }
  acq_loss_kernel<<<nblk, ACQ_THREADS, 0, h->stream>>>(a, dterm, dvar, dextra, stats, dvis, dloss, part, pidx);
  acq_argmax_final_kernel<<<1, ACQ_THREADS, 0, h->stream>>>(part, pidx, nblk, stats + 2, reinterpret_cast<long long*>(stats + 3));
  h->launches += 2;
  CK(cudaGetLastError());
  double hv[2];
  CK(cudaMemcpyAsync(hv, stats + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (loss_out && !loss_dev) CK(cudaMemcpyAsync(loss_out, dloss, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (val) *val = hv[0];
  memcpy(idx, &hv[1], sizeof(long long));
  return 0;
}

int gpb_mean_pairwise_distance(gpb_handle_t h, const double* X, long long n, int d, double* out) {
  if (!h || !X || !out || n < 1 || d < 1 || d > MAX_D) return fail(h, -1, "gpb_mean_pairwise_distance: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  const long long nblk = (n + 63) / 64;
  int rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)n * d + (size_t)nblk + 8);
  if (rc) return rc;
  double* dX = h->dstage;
  double* part = dX + (size_t)n * d;
  CK(cudaMemcpyAsync(dX, X, (size_t)n * d * sizeof(double), cudaMemcpyDefault, h->stream));
  pairwise_distance_sum_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>(dX, n, d, part);
  reduce_partials_kernel<<<1, 256, 0, h->stream>>>(part, (int)nblk, 1, part + nblk);
  h->launches += 2;
  CK(cudaGetLastError());
  double sum = 0.0;
  CK(cudaMemcpyAsync(&sum, part + nblk, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *out = sum / ((double)n * (double)n);
  return 0;
}

int gpb_cholesky(gpb_handle_t h, const double* A, long long n, double* L) {
  if (!h || !A || !L || n < 1) return fail(h, -1, "gpb_cholesky: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->tmp;
  int rc = resize_ws(h, w, n);
  if (rc) return rc;
  CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  if ((rc = load_matrix_into_L(h, w, A, n))) return rc;
  if ((rc = run_plan(h, w, w.potrf, true))) return rc;
  if ((rc = check_info(h))) return rc;
  if ((rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, (size_t)n * n))) return rc;
  unpack_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, h->stream>>>(w.Lbuf, w.Np, n, 0, h->dstage2);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(L, h->dstage2, (size_t)n * n * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_invert_cholesky(gpb_handle_t h, const double* L, long long n, double* Kinv) {
  if (!h || !L || !Kinv || n < 1) return fail(h, -1, "gpb_invert_cholesky: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->tmp;
  int rc = resize_ws(h, w, n);
  if (rc) return rc;
  CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  if ((rc = load_matrix_into_L(h, w, L, n))) return rc;
  if ((rc = invert_diag_blocks(h, w))) return rc;
  if ((rc = triangular_inverse(h, w))) return rc;
  if ((rc = run_plan(h, w, w.syrk, false))) return rc;
  if ((rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, (size_t)n * n))) return rc;
  unpack_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, h->stream>>>(w.Lbuf, w.Np, n, 1, h->dstage2);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(Kinv, h->dstage2, (size_t)n * n * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_solve_cholesky(gpb_handle_t h, const double* L, long long n, const double* b, int nrhs, double* x) {
  if (!h || !L || !b || !x || n < 1 || nrhs < 1 || nrhs > 128) return fail(h, -1, "gpb_solve_cholesky: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->tmp;
  int rc = resize_ws(h, w, n);
  if (rc) return rc;
  CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  if ((rc = load_matrix_into_L(h, w, L, n))) return rc;
  if ((rc = invert_diag_blocks(h, w))) return rc;
  if ((rc = triangular_inverse(h, w))) return rc;
  // z = W b ; x = U z.  b and x are (n, nrhs) row-major; z lives in w.alpha as (nrhs, Np).
  if ((rc = ensure_cap(h, &h->dstage2, &h->dstage2_cap, (size_t)2 * n * nrhs))) return rc;
  double* db = h->dstage2;
  double* dx = h->dstage2 + (size_t)n * nrhs;
  CK(cudaMemcpyAsync(db, b, (size_t)n * nrhs * sizeof(double), cudaMemcpyDefault, h->stream));
  trigemv_kernel<<<(unsigned)((n + 7) / 8), 256, 0, h->stream>>>(w.Wbuf, w.Np, n, true, db, nrhs, 1, nrhs, w.alpha, 1,
                                                                 w.Np);
  trigemv_kernel<<<(unsigned)((n + 7) / 8), 256, 0, h->stream>>>(w.Ubuf, w.Np, n, false, w.alpha, 1, w.Np, nrhs, dx, nrhs,
                                                                 1);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(x, dx, (size_t)n * nrhs * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int gpb_debug_plan(int n_blocks, int which, int* tasks_out, long long task_cap, double* launches_out, long long launch_cap,
                   long long* counts) {
  if (n_blocks < 1 || !counts) return -1;
  Workspace w;
  w.Nb = n_blocks;
  w.Np = (long long)n_blocks * 128;
  w.N = w.Np;
  build_potrf_plan(w);     // also fixes the phase cuts that build_trtri_plan splits by
  build_trtri_plan(w);
  build_syrk_plan(w);
  const Plan* p = nullptr;
  if (which == 0) p = &w.potrf;
  else if (which == 1) p = &w.trtri;
  else if (which == 2) p = &w.syrk;
  else if (which == 20) p = &w.trtri_late;
  else if (which >= 10 && which < 10 + w.n_phase) p = &w.trtri_phase[which - 10];
  counts[0] = p ? (long long)p->tasks.size() : 0;
  counts[1] = p ? (long long)p->launches.size() : 0;
  counts[2] = w.n_phase;
  for (int i = 0; i < Workspace::MAX_PHASE; ++i) {
    counts[3 + i] = i < w.n_phase ? w.phase_cut[i] : -1;
    counts[7 + i] = i < w.n_phase ? w.phase_ev[i] : -1;
  }
  if (!p) return which >= 10 && which < 10 + Workspace::MAX_PHASE ? 0 : -1;
  if (tasks_out && task_cap >= counts[0])
    for (size_t t = 0; t < p->tasks.size(); ++t) {
      const GemmTask& k = p->tasks[t];
      const int row[8] = {k.a_row, k.a_k, k.b_row, k.b_k, k.nk, k.c_row, k.c_col, k.aux};
      memcpy(tasks_out + 8 * t, row, sizeof(row));
    }
  if (launches_out && launch_cap >= counts[1])
    for (size_t l = 0; l < p->launches.size(); ++l) {
      const Launch& L = p->launches[l];
      const double row[17] = {(double)L.type, (double)L.k, (double)L.cfg, (double)L.matA, (double)L.matB, (double)L.matC,
                              (double)L.matCt, L.alpha, L.beta, (double)L.task_off, (double)L.ntasks, (double)L.epi,
                              (double)L.stream, (double)L.wait_ev, (double)L.wait_ev2, (double)L.rec_ev, (double)L.wait_ev3};
      memcpy(launches_out + 17 * l, row, sizeof(row));
    }
  return 0;
}

int gpb_debug_gemm(gpb_handle_t h, const double* A, const double* B, double* C, long long M, long long N, long long K,
                   int cfg, double alpha, double beta, int reps, double* ms) {
  if (!h || !A || !B || !C || M % 128 || N % 128 || K % 128 || cfg < 0 || cfg >= NUM_CFG || reps < 1)
    return fail(h, -1, "gpb_debug_gemm: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  Workspace& w = h->tmp;
  set_mat(w, M_DBG_A, const_cast<double*>(A), M, K, K);
  set_mat(w, M_DBG_B, const_cast<double*>(B), N, K, K);
  Plan p;
  PlanGuard guard(p);
  TaskSink s(p);
  for (long long i = 0; i < M / 128; ++i)
    for (long long j = 0; j < N / 128; ++j)
      s.block(cfg, (int)i * 128, 0, (int)j * 128, 0, (int)(K / 16), (int)i * 128, (int)j * 128);
  s.finish(cfg, EPI_STORE, M_DBG_A, M_DBG_B, M_DBG_C, M_NONE, 1.0, 0.0, false);
  int rc = upload_plan(h, p);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.C = C;
  ep.ldc = N;
  ep.Ct = nullptr;
  ep.ldct = 0;
  ep.alpha = alpha;
  ep.beta = beta;
  cudaEvent_t e0 = h->evx[0], e1 = h->evx[1];
  cudaEventRecord(e0, h->stream);
  for (int r = 0; r < reps; ++r)
    if ((rc = launch_gemm(h, w, cfg, EPI_STORE, M_DBG_A, M_DBG_B, p.d_tasks, (int)p.tasks.size(), ep))) return rc;
  cudaEventRecord(e1, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t / reps;
  return 0;
}

// Factor one 128x128 SPD block with the diagonal-block kernel selected by `mode` (0 / 1 / 2 as GPB_DIAG4), `reps`
// times back to back on the handle's stream; returns L, L^-1, sum log L_jj and the mean kernel time.
int gpb_debug_diag(gpb_handle_t h, const double* A, double* L, double* W, double* logdet, int mode, int reps, double* ms) {
  if (!h || !A || !L || !W || reps < 1 || mode < 0 || mode > 3) return fail(h, -1, "gpb_debug_diag: bad argument");
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  int rc = ensure_cap(h, &h->dstage, &h->dstage_cap, (size_t)3 * 128 * 128 + 8);
  if (rc) return rc;
  double* dA = h->dstage;
  double* dB = dA + 128 * 128;
  double* dW = dB + 128 * 128;
  double* dl = dW + 128 * 128;
  CK(cudaMemcpyAsync(dA, A, 128 * 128 * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  CK(cudaMemsetAsync(dW, 0, 128 * 128 * sizeof(double), h->stream));   // mode 2 leaves the strict upper blocks alone
  cudaEvent_t e0 = h->evx[0], e1 = h->evx[1];
  float total = 0.f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemcpyAsync(dB, dA, 128 * 128 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    cudaEventRecord(e0, h->stream);
    if (mode == 2) potrf_diagr_kernel<true><<<1, DR_THREADS, DR_SMEM, h->stream>>>(dB, 128, 0, dW, dl, h->dinfo, 128);
    else if (mode == 3) potrf_diagr_kernel<false><<<1, DR_THREADS, DR_SMEM, h->stream>>>(dB, 128, 0, dW, dl, h->dinfo, 128);
    else if (mode == 1) potrf_diag4_kernel<<<1, DIAG_THREADS, DIAG4_SMEM, h->stream>>>(dB, 128, 0, dW, dl, h->dinfo, 128);
    else potrf_diag_kernel<true><<<1, DIAG_THREADS, DIAG_SMEM, h->stream>>>(dB, 128, 0, dW, dl, h->dinfo, 128);
    cudaEventRecord(e1, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    total += t;
  }
  if (mode == 3) {   // what the library does off the chain: the blocks of L^-1 below the block diagonal
    complete_inverse_kernel<<<1, TP_THREADS, 0, h->stream>>>(dB, 128, 0, dW);
    h->launches++;
    CK(cudaGetLastError());
  }
  // the factor as the library's consumers see it: lower triangle (mode 2 does not touch the strict upper one)
  unpack_kernel<<<(128 * 128 + 255) / 256, 256, 0, h->stream>>>(dB, 128, 128, 0, dA);
  h->launches++;
  CK(cudaMemcpyAsync(L, dA, 128 * 128 * sizeof(double), cudaMemcpyDefault, h->stream));
  CK(cudaMemcpyAsync(W, dW, 128 * 128 * sizeof(double), cudaMemcpyDefault, h->stream));
  if (logdet) CK(cudaMemcpyAsync(logdet, dl, sizeof(double), cudaMemcpyDefault, h->stream));
  if (ms) *ms = total / reps;
  return check_info(h);
}

#ifdef GPB_TRACE
// Debug build only: point the CTA trace ring at a caller-provided device buffer of cap records (4 x u64 each),
// or read back how many records were written.
int gpb_debug_trace(gpb_handle_t h, unsigned long long* buf, unsigned int cap, unsigned int* count) {
  if (!h) return -1;
  CK(cudaSetDevice(h->device));
  if (int rc_order = order_after_caller(h)) return rc_order;
  CK(cudaDeviceSynchronize());
  if (count) CK(cudaMemcpyFromSymbol(count, g_trace_count, sizeof(unsigned int)));
  unsigned int zero = 0;
  CK(cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf)));
  CK(cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
  CK(cudaMemcpyToSymbol(g_trace_count, &zero, sizeof(zero)));
  return 0;
}
#endif

}  // extern "C"
