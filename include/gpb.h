/*
 * gpb.h -- C ABI of the B200-native Gaussian-process backend (libgpb.so).
 *
 * This is the drop-in boundary for proFit's Custom-GP numerics.  The reference has no FFI on this path (it is
 * numpy/scipy called in-process), so each entry point names the reference *function* it replaces; paths are
 * relative to the reference root.  The Python host layer (profit_b200/) binds these with ctypes and mirrors the
 * reference's operator interface on top (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - All matrices are C-order (row-major) IEEE fp64, exactly like the numpy arrays the reference passes.
 *   - Every pointer may be a host pointer (numpy) or a device pointer (torch CUDA tensor); the library detects
 *     which.  Caller owns all buffers passed in; the library owns its device workspaces inside the handle.
 *   - Return value: 0 = ok; > 0 = LAPACK-style info (1-based index of the first non-positive pivot: the
 *     matrix is not SPD -> the Python layer raises numpy.linalg.LinAlgError like gp_functions.py:143 would);
 *     < 0 = bad argument (-1), CUDA failure (-2), not ready (-3), eigen-solver failure (-4).  gpb_error_string() describes the last error.
 *   - One handle = one device + one stream; calls on a handle must be serialised by the caller.
 *   - Stream semantics for device pointers: the handle's stream is a blocking stream, i.e. it is ordered after work
 *     already queued on the legacy default stream (the one torch uses unless told otherwise).  A caller that
 *     produces device buffers on another stream names it with gpb_set_stream(); every later call is then ordered
 *     after the work queued on that stream at call time (event record + wait, no host synchronisation), which is
 *     what the Python layer does with torch.cuda.current_stream().  Every entry point is host-synchronous: when it
 *     returns, its outputs (host or device) are complete.
 *   - There is no CPU fallback anywhere behind this interface.
 */
#ifndef GPB_H_
#define GPB_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpb_handle_s* gpb_handle_t;

enum { GPB_KERNEL_RBF = 0, GPB_KERNEL_MATERN52 = 1 };

/* stage indices for gpb_stage_ms() */
enum {
  GPB_STAGE_ASSEMBLE = 0, /* prescale + K assembly (+ y rows)                       */
  GPB_STAGE_POTRF = 1,    /* blocked Cholesky incl. forward substitution of y        */
  GPB_STAGE_TRTRI = 2,    /* U = L^-T by recursive tile GEMMs                        */
  GPB_STAGE_SYRK = 3,     /* K^-1 = U U^T                                            */
  GPB_STAGE_GRAD = 4,     /* alpha, gradient contraction, reductions                 */
  GPB_STAGE_CROSS = 5,    /* predict: K(X*,X) + mean                                 */
  GPB_STAGE_VAR = 6,      /* predict: ||L^-1 k*||^2 tile GEMM + finish               */
  GPB_STAGE_TOTAL = 7,
  GPB_NUM_STAGES = 8
};

int gpb_version(void);

int gpb_create(int device, gpb_handle_t* out);
int gpb_destroy(gpb_handle_t h);
const char* gpb_error_string(gpb_handle_t h);

/* Name the CUDA stream (a cudaStream_t, passed as void*) on which the caller produces the device buffers it hands to
 * this handle; NULL (the default) = the legacy default stream.  See "Stream semantics" above.  The stream is
 * borrowed, never destroyed, and must stay alive while it is set. */
int gpb_set_stream(gpb_handle_t h, void* cuda_stream);

/* Upload the training set once; it stays resident across optimiser iterations.
 * Replaces the (X, y) arguments threaded through gp_functions.optimize (gp_functions.py:8-69) and the
 * surrogate's Xtrain/ytrain (custom_surrogate.py:122-142).  X: (n, d), y: (n, n_out), n_out <= 128. */
int gpb_set_train(gpb_handle_t h, const double* X, const double* y, long long n, int d, int n_out);

/* Negative log marginal likelihood and its gradient w.r.t. theta = [length_scale(n_ell), sigma_f, sigma_n]
 * on the LINEAR scale -- gp_functions.negative_log_likelihood_cholesky (gp_functions.py:103-187) with
 * log_scale_hyp=False, fixed_sigma_n=False; the host layer applies the log10 chain rule and the sigma_n slicing
 * (:181-186).  grad has n_ell + 2 entries and may be NULL when want_grad == 0.
 * After a call with want_grad != 0 the handle also holds alpha and L^-1 (usable by gpb_predict). */
int gpb_nll(gpb_handle_t h, int kind, const double* theta, int n_ell, int want_grad, double* nll, double* grad);

/* Factorise K(theta) and cache alpha = K^-1 y and W = L^-1 for prediction.  Replaces the per-call
 * GPSurrogate.Ky / .alpha / gp_functions.invert(Ky) of custom_surrogate.py:24-45,114-116. */
int gpb_factorize(gpb_handle_t h, int kind, const double* theta, int n_ell);

/* Predictive mean (M, n_out) and variance (M) -- GPSurrogate.predict (custom_surrogate.py:95-120):
 * mean = K*^T alpha, var = max(sf^2 - diag(K*^T K^-1 K*), 1e-10) (+ sn^2 if add_data_variance == 1).
 * add_data_variance == 2 adds sn^2 before the clamp, which is gp_functions.predict_f (gp_functions.py:484-488).
 * Either output may be NULL. */
int gpb_predict(gpb_handle_t h, const double* Xc, long long M, int add_data_variance, double* mean, double* var);

/* Full predictive covariance -- gp_functions.predict_f with return_full_cov=True (gp_functions.py:484-490):
 * cov = max(K(X*, X*) [+ sn^2 I if noise_on_diagonal] - K*^T K^-1 K*, 1e-10) element-wise, (M, M) row-major, and
 * optionally the mean (M, n_out).  K*^T K^-1 K* = Vt Vt^T with Vt = K*^T L^-T: two tile GEMMs.  M <= 16384. */
int gpb_predict_cov(gpb_handle_t h, const double* Xc, long long M, int noise_on_diagonal, double* mean, double* cov);

/* argmax with lowest-index tie-break, like np.argmax in aquisition_functions.py:68. v: (M). */
int gpb_argmax(gpb_handle_t h, const double* v, long long M, long long* idx, double* val);

/* Append m observed points to a factorised model without refactorising -- the in-place extension the reference
 * leaves as a TODO in GPSurrogate.add_training_data (custom_surrogate.py:122-134, ":132 TODO: Update Ky").
 * Per point O(N^2): l = L^-1 k(X, x), d^2 = sf^2 + sn^2 - |l|^2, new row of L^-1 is [-(L^-T l)^T/d, 1/d]; alpha, z and
 * log|K| follow.  Afterwards gpb_predict / gpb_marginal_variance see N + m training points at the same theta.
 * Returns N + 1 (LAPACK-style) when the extended matrix is not SPD; the model is then left unchanged for that
 * point.  When the padded capacity is exhausted the library re-uploads and refactorises internally.
 * Xnew: (m, d), ynew: (m, n_out). */
int gpb_append_train(gpb_handle_t h, const double* Xnew, const double* ynew, long long m);

/* Acquisition functions of profit/al/aquisition_functions.py evaluated on the device. */
enum {
  GPB_ACQ_VARIANCE = 0,         /* SimpleExploration.calculate_loss              (:107-119)  params: none            */
  GPB_ACQ_DISTANCE_PENALTY = 1, /* ExplorationWithDistancePenalty.calculate_loss (:146-158)  [weight, d, last[d]]    */
  GPB_ACQ_WEIGHTED = 2,         /* WeightedExploration.calculate_loss            (:183-198)  [weight]                */
  GPB_ACQ_POI = 3,              /* ProbabilityOfImprovement.calculate_loss       (:212-220)  [ymax[n_out]]           */
  GPB_ACQ_EI = 4,               /* ExpectedImprovement.calculate_loss/sigma_part (:253-276)  none                    */
  GPB_ACQ_EI2 = 5               /* ExpectedImprovement2.calculate_loss           (:303-315)  [xi, find_min, ymax[]]  */
};

/* The once-per-batch transform of the predictive mean (M, n_out) -> out (M, n_out):
 *   GPB_ACQ_WEIGHTED: AcquisitionFunction.normalize(mu) column-wise (:76-84, :200-201); no params.
 *   GPB_ACQ_EI:       ExpectedImprovement.mu_part (:262-269), params = [xi, find_min, ymax[n_out]]:
 *                     normalize(max(+-mu - ymax, 0)) - xi.
 * n_out <= 16.  out may alias mean.
 *
 * Sharded candidate sets: the normalisations are over ALL candidates.  stats_out (host, 2*n_out: [min, max] per column
 * of the quantity being normalised on this shard) and stats_in (host, same layout: the global values) let the caller
 * combine shards -- call once with stats_out, reduce over ranks, call again with stats_in.  Both may be NULL. */
int gpb_acq_prepare(gpb_handle_t h, int kind, const double* mean, long long M, int n_out, const double* params,
                    int n_params, double* out, const double* stats_in, double* stats_out);

/* Column-wise [min, max] of v (M, ncol) -> out (host, 2*ncol); transform 1 takes sqrt first.  The per-shard half of
 * AcquisitionFunction.normalize (:76-84) for gpb_acquire's stats_in. */
int gpb_acq_stats(gpb_handle_t h, const double* v, long long M, int ncol, int transform, double* out);

/* Loss of one acquisition function for every candidate, summed over outputs (axis=1), and the pick
 * idx = np.argmax(loss * mask) of AcquisitionFunction._find_next_candidates (:63-68): `visited` (host, n_visited
 * indices) are the rows whose mask is 0 -- they compete with value loss*0 exactly as in the reference; NaN ordering
 * and the lowest-index tie-break follow np.argmax.
 *   term : (M, n_out)  predictive mean (POI, EI2) | gpb_acq_prepare output (WEIGHTED, EI) | NULL otherwise
 *   var  : (M)         predictive or marginal variance
 *   extra: (M, n_out) standard-normal draws (POI) | (M, d) candidate coordinates (DISTANCE_PENALTY) | NULL otherwise
 *   loss_out: (M) or NULL (the unmasked loss, what calculate_loss returns); val may be NULL.
 * Inputs left on the device by gpb_predict never travel to the host: only (idx, val) come back.
 * stats_in (host, [min, max] or NULL): global statistics of var (sqrt(var) for EI) when this call sees one shard of
 * the candidate set (gpb_acq_stats on every shard, reduced by the caller); NULL = this call sees all candidates. */
int gpb_acquire(gpb_handle_t h, int kind, const double* term, const double* var, const double* extra, long long M,
                int n_out, const double* params, int n_params, const long long* visited, int n_visited, double* loss_out,
                long long* idx, double* val, const double* stats_in);

/* Dense kernel matrix and derivative tensor -- python_kernels.RBF (python_kernels.py:8-57) and the Matern-5/2
 * twin.  K: (na, nb); dK: (na, nb, n_ell + 2) or NULL. */
int gpb_kernel_matrix(gpb_handle_t h, int kind, const double* Xa, long long na, const double* Xb, long long nb, int d,
                      const double* ell, int n_ell, double sigma_f, double sigma_n, double* K, double* dK);

/* Linear-embedding kernel -- python_kernels.LinearEmbedding (python_kernels.py:60-114):
 * K = sf^2 exp(-1/2 |R (x - y)|^2) + sn^2 I with R (d, D) row-major; K: (na, nb); dK: (na, nb, d*D + 2) or NULL, ordered
 * [R.ravel()..., sf, sn].  dK/dR_ab = -K_pure (R delta)_a delta_b is the analytic derivative (upstream's einsum at
 * :108-110 is flagged "TODO: check" and only shape-consistent for d = 1). */
int gpb_linear_embedding(gpb_handle_t h, const double* Xa, long long na, const double* Xb, long long nb, int D,
                         const double* R, int d, double sigma_f, double sigma_n, double* K, double* dK);

/* Marginal variance of Osborne (2012) -- gp_functions.marginal_variance_BBQ (gp_functions.py:376-454):
 * out[c] = dm_c^T Hinv dm_c + pred_var[c] with dm_c = d(mean_c)/d(theta), theta = [length_scale..., sigma_f, sigma_n]
 * of the factorised model (gpb_factorize).  hess_inv: (P, P) row-major, P = n_ell + 2 (already zero-padded if sigma_n was
 * fixed); pred_var: (M) or NULL; out: (M).  Single-output models only. */
int gpb_marginal_variance(gpb_handle_t h, const double* Xc, long long M, const double* hess_inv, const double* pred_var,
                          double* out);

/* mean over all ordered pairs of |x_i - x_j| -- the data scale GaussianProcess.infer_hyperparameters derives the
 * default length scale from (gaussian_process.py:124-136) without its (N, N, D) host tensor.  X: (n, d). */
int gpb_mean_pairwise_distance(gpb_handle_t h, const double* X, long long n, int d, double* out);

/* L = cholesky(A) (lower, upper zeroed) -- np.linalg.cholesky at gp_functions.py:143,341. A: (n, n). */
int gpb_cholesky(gpb_handle_t h, const double* A, long long n, double* L);

/* K^-1 from its Cholesky factor -- gp_functions.invert_cholesky (gp_functions.py:304-322). */
int gpb_invert_cholesky(gpb_handle_t h, const double* L, long long n, double* Kinv);

/* x = L^T \ (L \ b), b: (n, nrhs) -- gp_functions.solve_cholesky (gp_functions.py:72-100). nrhs <= 128. */
int gpb_solve_cholesky(gpb_handle_t h, const double* L, long long n, const double* b, int nrhs, double* x);

/* Truncated eigen-decomposition exits of the reference (the branches taken when the Cholesky factorisation is not used).
 * Device solver: Lanczos with full re-orthogonalisation + implicit QL on the projected tridiagonal matrix, deterministic
 * start vector, Ritz pairs converged to ~1e-13 |lambda_max| (the reference's ARPACK call stops at tol = 1e-3 min|l| from a
 * random start; its values scatter by ~5e-7 relative).  Status -4 = the solver did not converge / broke down.
 *
 * gpb_eigsh: the k largest eigenpairs of Ky(theta) of the resident training set -- eigsh(Ky, neig, tol=clip_eig) in the
 * loop of gp_functions.negative_log_likelihood (gp_functions.py:273-275).  w_out: k eigenvalues ascending (eigsh's order,
 * so w_out[0] is what the loop compares with clip_eig, :280).  Calls with the same (kind, theta) continue one Lanczos run;
 * k_hint (0 = none) is the largest k the caller expects to request for this theta -- the reference's loop doubles neig up to
 * 0.05 N -- so that the Krylov space is built and diagonalised once.  steps_out (may be NULL) <- Lanczos steps used. */
int gpb_eigsh(gpb_handle_t h, int kind, const double* theta, int n_ell, int k, int k_hint, double* w_out, int* steps_out);

/* NLL and gradient on that exit (gp_functions.py:283-300) from the eigenpairs gpb_eigsh left resident:
 * w <- max(w, 1e-10), alpha = Q diag(1/w) Q^T y, nll = 1/2 (y^T alpha + sum log w + neig_term log 2 pi) with
 * neig_term = min(neig, N) of :286; grad = 1/2 tr((invert(Ky) - alpha alpha^T) dKy) on the linear scale, invert(Ky) by
 * Cholesky with the reference's single 1e-6 diagonal retry (:340-343) -- a second failure returns the pivot index (> 0),
 * as np.linalg.cholesky would raise there.  Single-output training sets only (the reference calls nll.item()). */
int gpb_nll_eig(gpb_handle_t h, int neig_term, int want_grad, double* nll, double* grad);

/* eigsh(K, neig, tol) of gp_functions.invert (gp_functions.py:351,359) for a caller-provided symmetric matrix A (n, n).
 * reuse != 0: A is the matrix of the previous call (its Lanczos run is continued; A is not read and may be NULL).
 * w_out: k eigenvalues ascending; Qt_out: (k, n), row i = eigenvector of w_out[i], or NULL. */
int gpb_eigsh_matrix(gpb_handle_t h, const double* A, long long n, int k, int reuse, double* w_out, double* Qt_out,
                     int* steps_out);

/* K^-1 = Q diag(1/w) Q^T (gp_functions.py:366-369) from the resident eigenvectors of gpb_eigsh_matrix and the caller's
 * eigenvalues w (host, k entries: the reference replaces slightly negative ones by 1e-2 first).  Kinv: (n, n). */
int gpb_eig_inverse(gpb_handle_t h, const double* w, double* Kinv);

/* Device milliseconds of each stage of the last gpb_nll / gpb_factorize / gpb_predict call (CUDA events on the
 * handle's stream).  out has GPB_NUM_STAGES entries. */
int gpb_stage_ms(gpb_handle_t h, double* out);

/* Number of kernels launched by this handle since creation (the host layer reports it as gpu_launches). */
long long gpb_launch_count(gpb_handle_t h);

/* Test hook, host only (no device is touched): the launch / tile-task lists the library would run for a matrix of
 * n_blocks 128-blocks.  which: 0 = Cholesky, 1 = full triangular inverse, 2 = K^-1 = U U^T, 10 + i = early phase i of
 * the triangular inverse, 20 = its late part.  tasks_out: 8 ints per task {a_row, a_k, b_row, b_k, nk, c_row, c_col, aux};
 * launches_out: 17 doubles per launch {type (0 = diagonal block, 1 = tile GEMM), k, cfg, matA, matB, matC, matCt, alpha,
 * beta, task_off, ntasks, epi, stream, wait_ev, wait_ev2, rec_ev, wait_ev3} with matrices 0 = L, 1 = U, 2 = W, 3 = Dinv.
 * counts (11 entries): {n_tasks, n_launches, n_phases, cut[0..3], event[0..3]} -- phase i may start once `event[i]`
 * of the Cholesky plan has fired (always filled; pass NULL / 0 for the arrays to size them).
 * tests/test_plans.py replays these lists with numpy and checks that they compute the factor and the inverses. */
int gpb_debug_plan(int n_blocks, int which, int* tasks_out, long long task_cap, double* launches_out, long long launch_cap,
                   long long* counts);

/* Test/benchmark hook: C(MxN) = alpha * A(MxK) * B(NxK)^T + beta * C with the tile-GEMM kernel, device pointers,
 * all dimensions multiples of 128.  cfg selects the tile configuration.  The product is applied `reps` times;
 * *ms receives the average device time of one application. */
int gpb_debug_gemm(gpb_handle_t h, const double* A, const double* B, double* C, long long M, long long N, long long K,
                   int cfg, double alpha, double beta, int reps, double* ms);

/* Test/benchmark hook for the serial stage of the Cholesky (np.linalg.cholesky, gp_functions.py:143): factor ONE
 * 128x128 SPD block A (row-major, host or device) with the diagonal-block kernel `mode` (3 = blocked DMMA kernel as it
 * runs on the panel chain: factor + 32x32 diagonal inverses, completed by complete_inverse_kernel after the timing;
 * 2 = blocked DMMA kernel with the full inverse; 1 = rank-4 register sweep, 0 = rank-1 register sweep), `reps` times.
 * L <- lower factor, W <- L^-1 (both 128x128 row-major, strict upper zero), *logdet <- sum_j log L_jj,
 * *ms <- mean device time of one factorisation.  Returns the LAPACK-style info like gpb_nll. */
int gpb_debug_diag(gpb_handle_t h, const double* A, double* L, double* W, double* logdet, int mode, int reps, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* GPB_H_ */
