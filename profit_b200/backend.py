"""Device-resident GP state: one `GPBackend` = one gpb handle = one GPU + one stream.

Thin, allocation-free wrapper over the C ABI; all numerics run in libgpb's CUDA kernels.
"""
import ctypes

import numpy as np

from . import _lib

KIND_RBF = 0
KIND_MATERN52 = 1

# acquisition kinds of gpb_acquire (include/gpb.h)
ACQ_VARIANCE, ACQ_DISTANCE_PENALTY, ACQ_WEIGHTED, ACQ_POI, ACQ_EI, ACQ_EI2 = range(6)


class GPBackend:
    def __init__(self, device=None):
        lib = _lib.load()
        if device is None:
            device = _default_device()
        self.device = int(device)
        self._h = ctypes.c_void_p()
        st = lib.gpb_create(self.device, ctypes.byref(self._h))
        if st != 0:
            raise RuntimeError(
                f"gpb_create(device={self.device}) failed with status {st}: a CUDA device is required "
                "(profit_b200 has no CPU fallback)"
            )
        self._lib = lib
        self.n = 0
        self.d = 0
        self.n_out = 0
        # bookkeeping for callers that share one backend (set by B200GPSurrogate): what is resident / factorised
        self.resident_key = None
        self.factor_key = None
        self._caller_stream = 0        # cudaStream_t last passed to gpb_set_stream (0 = legacy default stream)
        self.non_spd_events = []       # jitter retries of gp_functions.negative_log_likelihood since the last optimize()
        self.noise_override = None     # set by B200GPSurrogate when the resident factor carries the 1e-6 jitter retry

    def _order(self, *arrays):
        """Stream ordering for device arguments (include/gpb.h, "Stream semantics"): when any argument is a torch CUDA
        tensor, tell the library which torch stream is current so that the call is ordered after the work queued on
        it -- a producer on a non-default stream needs no manual synchronisation."""
        if not any(_lib.is_cuda_tensor(a) for a in arrays):
            return
        import torch

        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        if stream != self._caller_stream:
            _lib.check(self._lib.gpb_set_stream(self._h, ctypes.c_void_p(stream)), self._h, "set_stream")
            self._caller_stream = stream

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gpb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- training set ---------------------------------------------------------------------------------
    def set_train(self, X, y):
        X = _lib.as_f64(X)
        y = _lib.as_f64(y)
        if X.ndim != 2:
            raise ValueError(f"X should have shape (n, d) but has shape {tuple(X.shape)}")
        n, d = int(X.shape[0]), int(X.shape[1])
        if y.ndim == 1:
            n_out = 1
        elif y.ndim == 2:
            n_out = int(y.shape[1])
        else:
            raise ValueError(f"y should have shape (n,) or (n, D) but has shape {tuple(y.shape)}")
        if int(y.shape[0]) != n:
            raise ValueError(f"mismatched number of data points for X and y: {n} != {int(y.shape[0])}")
        self.resident_key = self.factor_key = self.noise_override = None
        self._order(X, y)
        _lib.check(self._lib.gpb_set_train(self._h, _lib.ptr(X), _lib.ptr(y), n, d, n_out), self._h, "set_train")
        self.n, self.d, self.n_out = n, d, n_out

    # -- likelihood -----------------------------------------------------------------------------------
    def nll(self, kind, theta, want_grad=False):
        """theta = [length_scale..., sigma_f, sigma_n] on the linear scale -> nll [, grad (P,)]."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).ravel()
        n_ell = theta.size - 2
        self.factor_key = self.noise_override = None
        self.nll_calls = getattr(self, "nll_calls", 0) + 1        # objective evaluations (multistart_optimize reports them)
        out = ctypes.c_double()
        grad = np.empty(theta.size) if want_grad else None
        st = self._lib.gpb_nll(self._h, int(kind), _lib.ptr(theta), n_ell, int(bool(want_grad)), ctypes.byref(out),
                               _lib.ptr(grad))
        _lib.check(st, self._h, "nll")
        return (out.value, grad) if want_grad else out.value

    def factorize(self, kind, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64).ravel()
        self.factor_key = self.noise_override = None
        _lib.check(self._lib.gpb_factorize(self._h, int(kind), _lib.ptr(theta), theta.size - 2), self._h, "factorize")

    def _check_candidates(self, Xc):
        Xc = _lib.as_f64(Xc)
        if Xc.ndim != 2 or int(Xc.shape[1]) != self.d:
            raise ValueError(f"Xpred should have shape (n, {self.d}) but has shape {tuple(Xc.shape)}")
        return Xc

    @staticmethod
    def _check_out(out, shape, name):
        """A caller-provided output buffer must hold exactly the result (trailing unit axes are tolerated)."""
        if out is None:
            return
        got = tuple(int(v) for v in out.shape)
        if int(np.prod(got)) != int(np.prod(shape)) or (got[:1] != tuple(shape[:1])):
            raise ValueError(f"{name} should have shape {tuple(shape)} but has shape {got}")

    def predict(self, Xc, add_data_variance=True, out_mean=None, out_var=None, want_mean=True, want_var=True):
        Xc = self._check_candidates(Xc)
        M = int(Xc.shape[0])
        self._check_out(out_mean, (M, self.n_out), "out_mean")
        self._check_out(out_var, (M,), "out_var")
        if want_mean and out_mean is None:
            out_mean = np.empty((M, self.n_out))
        if want_var and out_var is None:
            out_var = np.empty(M)
        self._order(Xc, out_mean, out_var)
        st = self._lib.gpb_predict(self._h, _lib.ptr(Xc), M, int(add_data_variance),
                                   _lib.ptr(out_mean) if want_mean else None, _lib.ptr(out_var) if want_var else None)
        _lib.check(st, self._h, "predict")
        return out_mean, out_var

    def predict_cov(self, Xc, noise_on_diagonal=True, want_mean=True, out_cov=None):
        """Full (M, M) predictive covariance [and the mean] -- predict_f(..., return_full_cov=True)."""
        Xc = self._check_candidates(Xc)
        M = int(Xc.shape[0])
        self._check_out(out_cov, (M, M), "out_cov")
        mean = np.empty((M, self.n_out)) if want_mean else None
        cov = np.empty((M, M)) if out_cov is None else out_cov
        self._order(Xc, cov)
        st = self._lib.gpb_predict_cov(self._h, _lib.ptr(Xc), M, int(bool(noise_on_diagonal)), _lib.ptr(mean), _lib.ptr(cov))
        _lib.check(st, self._h, "predict_cov")
        return mean, cov

    def marginal_variance(self, Xc, hess_inv, predictive_variance=None):
        """dm^T H^-1 dm + predictive variance per candidate (gp_functions.marginal_variance_BBQ) -> (M,) array."""
        Xc = self._check_candidates(Xc)
        H = np.ascontiguousarray(hess_inv, dtype=np.float64)
        if H.ndim != 2 or H.shape[0] != H.shape[1]:
            raise ValueError(f"hess_inv should be a square (P, P) matrix but has shape {H.shape}")
        M = int(Xc.shape[0])
        pv = None if predictive_variance is None else np.ascontiguousarray(predictive_variance, dtype=np.float64).ravel()
        out = np.empty(M)
        self._order(Xc)
        st = self._lib.gpb_marginal_variance(self._h, _lib.ptr(Xc), M, _lib.ptr(H), _lib.ptr(pv), _lib.ptr(out))
        _lib.check(st, self._h, "marginal_variance")
        return out

    def argmax(self, v):
        v = _lib.as_f64(v)
        idx = ctypes.c_longlong()
        val = ctypes.c_double()
        self._order(v)
        _lib.check(self._lib.gpb_argmax(self._h, _lib.ptr(v), int(v.shape[0]), ctypes.byref(idx), ctypes.byref(val)),
                   self._h, "argmax")
        return int(idx.value), float(val.value)

    # -- active learning ------------------------------------------------------------------------------
    def append_train(self, X_new, y_new):
        """Extend the factorised model by the rows (X_new, y_new) in O(N^2) per row (gpb_append_train)."""
        X_new = _lib.as_f64(X_new)
        y_new = _lib.as_f64(y_new)
        if X_new.ndim != 2 or int(X_new.shape[1]) != self.d:
            raise ValueError(f"X should have shape (m, {self.d}) but has shape {tuple(X_new.shape)}")
        m = int(X_new.shape[0])
        if y_new.ndim == 1:
            y_new = y_new.reshape(m, -1)
        if tuple(y_new.shape) != (m, self.n_out):
            raise ValueError(f"y should have shape ({m}, {self.n_out}) but has shape {tuple(y_new.shape)}")
        if m == 0:
            return
        self.resident_key = self.factor_key = None
        self._order(X_new, y_new)
        _lib.check(self._lib.gpb_append_train(self._h, _lib.ptr(X_new), _lib.ptr(y_new), m), self._h, "append_train")
        self.n += m

    def acq_prepare(self, kind, mean, params=(), out=None, stats_in=None, return_stats=False):
        """Once-per-batch transform of the predictive mean (M, n_out): normalize(mu) for ACQ_WEIGHTED, the
        ExpectedImprovement.mu_part for ACQ_EI.  ``mean`` / ``out`` may be numpy arrays or torch CUDA tensors.
        ``stats_in`` (n_out, 2) overrides the [min, max] of the normalisation (global values of a sharded candidate
        set); ``return_stats`` also returns this shard's own (n_out, 2)."""
        mean = _lib.as_f64(mean)
        M = int(mean.shape[0])
        n_out = 1 if mean.ndim == 1 else int(mean.shape[1])
        params = np.ascontiguousarray(params, dtype=np.float64).ravel()
        if out is None:
            out = np.empty((M, n_out))
        sin = None if stats_in is None else np.ascontiguousarray(stats_in, dtype=np.float64).reshape(n_out, 2)
        sout = np.empty((n_out, 2)) if return_stats else None
        self._order(mean, out)
        st = self._lib.gpb_acq_prepare(self._h, int(kind), _lib.ptr(mean), M, n_out,
                                       _lib.ptr(params) if params.size else None, int(params.size), _lib.ptr(out),
                                       _lib.ptr(sin), _lib.ptr(sout))
        _lib.check(st, self._h, "acq_prepare")
        return (out, sout) if return_stats else out

    def acq_stats(self, v, transform=0):
        """Column-wise [min, max] of v (M,) or (M, ncol) on the device -> (ncol, 2); ``transform=1`` takes sqrt first."""
        v = _lib.as_f64(v)
        M = int(v.shape[0])
        ncol = 1 if v.ndim == 1 else int(v.shape[1])
        out = np.empty((ncol, 2))
        self._order(v)
        _lib.check(self._lib.gpb_acq_stats(self._h, _lib.ptr(v), M, ncol, int(transform), _lib.ptr(out)), self._h, "acq_stats")
        return out

    def acquire(self, kind, var, term=None, extra=None, params=(), visited=(), n_out=1, loss_out=None, stats=None):
        """Loss of acquisition function ``kind`` over all candidates and ``np.argmax(loss * mask)`` (mask zero on the
        ``visited`` rows) -> (index, masked value).  ``loss_out`` (numpy (M,) or torch CUDA) receives the unmasked
        loss when given; device-resident inputs stay on the device and only the pick comes back.  ``stats`` = global
        [min, max] of var (sqrt(var) for ACQ_EI) when ``var`` is one shard of the candidate set."""
        var = _lib.as_f64(var)
        M = int(var.shape[0])
        self._check_out(var, (M,), "var")
        if term is not None:
            self._check_out(term, (M, int(n_out)), "term")
        if extra is not None and int(extra.shape[0]) != M:
            raise ValueError(f"extra should have {M} rows but has shape {tuple(extra.shape)}")
        self._check_out(loss_out, (M,), "loss_out")
        params = np.ascontiguousarray(params, dtype=np.float64).ravel()
        vis = (ctypes.c_longlong * max(1, len(visited)))(*[int(v) for v in visited])
        idx = ctypes.c_longlong()
        val = ctypes.c_double()
        sin = None if stats is None else np.ascontiguousarray(stats, dtype=np.float64).ravel()[:2].copy()
        self._order(var, term, extra, loss_out)
        st = self._lib.gpb_acquire(self._h, int(kind), _lib.ptr(None if term is None else _lib.as_f64(term)),
                                   _lib.ptr(var), _lib.ptr(None if extra is None else _lib.as_f64(extra)), M, int(n_out),
                                   _lib.ptr(params) if params.size else None, int(params.size), vis, len(visited),
                                   _lib.ptr(loss_out), ctypes.byref(idx), ctypes.byref(val), _lib.ptr(sin))
        _lib.check(st, self._h, "acquire")
        return int(idx.value), float(val.value)

    # -- dense twins ----------------------------------------------------------------------------------
    def kernel_matrix(self, kind, Xa, Xb, ell, sigma_f, sigma_n, eval_gradient=False, out_K=None, out_dK=None):
        """K (na, nb) [and dK (na, nb, P)].  ``out_K`` / ``out_dK`` may be preallocated numpy arrays or torch CUDA
        tensors; device outputs are written in place by the kernel (no staging copy)."""
        Xa = _lib.as_f64(Xa)
        Xb = _lib.as_f64(Xb)
        ell = np.ascontiguousarray(ell, dtype=np.float64).ravel()
        if Xa.ndim != 2 or Xb.ndim != 2 or int(Xa.shape[1]) != int(Xb.shape[1]):
            raise ValueError(f"X and Y should have shapes (n, d) and (m, d) but have {tuple(Xa.shape)} and {tuple(Xb.shape)}")
        na, nb, d = int(Xa.shape[0]), int(Xb.shape[0]), int(Xa.shape[1])
        if ell.size not in (1, d):
            raise ValueError(f"length_scale must be a scalar or have {d} entries, got {ell.size}")
        if na == 0 or nb == 0:     # nothing to evaluate: the reference returns empty arrays of these shapes
            K0, dK0 = np.empty((na, nb)), np.empty((na, nb, ell.size + 2))
            return (K0, dK0) if eval_gradient else K0
        K = np.empty((na, nb)) if out_K is None else out_K
        dK = (np.empty((na, nb, ell.size + 2)) if out_dK is None else out_dK) if eval_gradient else None
        self._order(Xa, Xb, K, dK)
        st = self._lib.gpb_kernel_matrix(self._h, int(kind), _lib.ptr(Xa), na, _lib.ptr(Xb), nb, d, _lib.ptr(ell),
                                         ell.size, float(sigma_f), float(sigma_n), _lib.ptr(K), _lib.ptr(dK))
        _lib.check(st, self._h, "kernel_matrix")
        return (K, dK) if eval_gradient else K

    def linear_embedding(self, Xa, Xb, R, sigma_f, sigma_n, eval_gradient=False):
        """python_kernels.LinearEmbedding on the device: K (na, nb) [and dK (na, nb, R.size + 2)], R (d, D)."""
        Xa, Xb, R = _lib.as_f64(Xa), _lib.as_f64(Xb), _lib.as_f64(R)
        if Xa.ndim != 2 or Xb.ndim != 2 or R.ndim != 2 or not (Xa.shape[1] == Xb.shape[1] == R.shape[1]):
            raise ValueError(f"X, Y and R should have shapes (n, D), (m, D) and (d, D) but have {tuple(Xa.shape)}, "
                             f"{tuple(Xb.shape)} and {tuple(R.shape)}")
        na, nb, D, d = int(Xa.shape[0]), int(Xb.shape[0]), int(Xa.shape[1]), int(R.shape[0])
        K = np.empty((na, nb))
        dK = np.empty((na, nb, d * D + 2)) if eval_gradient else None
        if na and nb:
            st = self._lib.gpb_linear_embedding(self._h, _lib.ptr(Xa), na, _lib.ptr(Xb), nb, D, _lib.ptr(R), d,
                                                float(sigma_f), float(sigma_n), _lib.ptr(K), _lib.ptr(dK))
            _lib.check(st, self._h, "linear_embedding")
        return (K, dK) if eval_gradient else K

    def mean_pairwise_distance(self, X):
        """mean_ij |x_i - x_j| on the device (gaussian_process.py:126-128 without the (N, N, D) tensor)."""
        X = _lib.as_f64(X)
        if X.ndim != 2:
            raise ValueError(f"X should have shape (n, d) but has shape {tuple(X.shape)}")
        out = ctypes.c_double()
        _lib.check(self._lib.gpb_mean_pairwise_distance(self._h, _lib.ptr(X), int(X.shape[0]), int(X.shape[1]),
                                                        ctypes.byref(out)), self._h, "mean_pairwise_distance")
        return float(out.value)

    @staticmethod
    def _check_square(A, name):
        if A.ndim != 2 or int(A.shape[0]) != int(A.shape[1]):
            raise ValueError(f"{name} should be a square matrix but has shape {tuple(A.shape)}")

    def cholesky(self, A):
        A = _lib.as_f64(A)
        self._check_square(A, "A")
        L = np.empty_like(A)
        _lib.check(self._lib.gpb_cholesky(self._h, _lib.ptr(A), int(A.shape[0]), _lib.ptr(L)), self._h, "cholesky")
        return L

    def invert_cholesky(self, L):
        L = _lib.as_f64(L)
        self._check_square(L, "L")
        out = np.empty_like(L)
        _lib.check(self._lib.gpb_invert_cholesky(self._h, _lib.ptr(L), int(L.shape[0]), _lib.ptr(out)), self._h,
                   "invert_cholesky")
        return out

    def solve_cholesky(self, L, b):
        L = _lib.as_f64(L)
        b = _lib.as_f64(b)
        self._check_square(L, "L")
        if b.ndim not in (1, 2) or int(b.shape[0]) != int(L.shape[0]):
            raise ValueError(f"b should have {int(L.shape[0])} rows but has shape {tuple(b.shape)}")
        nrhs = 1 if b.ndim == 1 else int(b.shape[1])
        x = np.empty_like(b)
        _lib.check(self._lib.gpb_solve_cholesky(self._h, _lib.ptr(L), int(L.shape[0]), _lib.ptr(b), nrhs, _lib.ptr(x)),
                   self._h, "solve_cholesky")
        return x

    # -- truncated eigen-decomposition exits (gp_functions.py:273-301, 350-373) -----------------------------------
    def eigsh(self, kind, theta, k, k_hint=0):
        """The k largest eigenvalues of Ky(theta) of the resident training set, ascending like
        ``scipy.sparse.linalg.eigsh(Ky, k)`` at gp_functions.py:273-275; the eigenvectors stay on the device.
        ``k_hint``: the largest k the caller expects to ask for with this theta (one Krylov space for all passes)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).ravel()
        k = int(min(k, self.n))
        w = np.empty(k)
        steps = ctypes.c_int()
        self.factor_key = self.noise_override = None
        _lib.check(self._lib.gpb_eigsh(self._h, int(kind), _lib.ptr(theta), theta.size - 2, k, int(min(k_hint, self.n)),
                                       _lib.ptr(w), ctypes.byref(steps)), self._h, "eigsh")
        self.last_eig_steps = steps.value
        return w

    def nll_eig(self, neig_term, want_grad=False, n_params=None):
        """NLL [, gradient on the linear scale] from the eigenpairs the last ``eigsh`` call left resident
        (gp_functions.py:283-300)."""
        out = ctypes.c_double()
        grad = np.empty(int(n_params)) if want_grad else None
        self.factor_key = self.noise_override = None
        _lib.check(self._lib.gpb_nll_eig(self._h, int(neig_term), int(bool(want_grad)), ctypes.byref(out), _lib.ptr(grad)),
                   self._h, "nll_eig")
        return (out.value, grad) if want_grad else out.value

    def eigsh_matrix(self, A, k, reuse=False, want_vectors=False):
        """``eigsh(A, k)`` of gp_functions.invert (:351,359): (w ascending, Q (n, k) or None)."""
        A = _lib.as_f64(A)
        self._check_square(A, "K")
        n = int(A.shape[0])
        k = int(min(k, n))
        w = np.empty(k)
        Qt = np.empty((k, n)) if want_vectors else None
        steps = ctypes.c_int()
        self._order(A)
        _lib.check(self._lib.gpb_eigsh_matrix(self._h, _lib.ptr(A), n, k, int(bool(reuse)), _lib.ptr(w), _lib.ptr(Qt),
                                              ctypes.byref(steps)), self._h, "eigsh_matrix")
        self.last_eig_steps = steps.value
        return w, (Qt.T if want_vectors else None)

    def eig_inverse(self, w, n):
        """Q diag(1/w) Q^T from the eigenvectors the last ``eigsh_matrix`` call left resident (gp_functions.py:366-369)."""
        w = np.ascontiguousarray(w, dtype=np.float64).ravel()
        out = np.empty((int(n), int(n)))
        _lib.check(self._lib.gpb_eig_inverse(self._h, _lib.ptr(w), _lib.ptr(out)), self._h, "eig_inverse")
        return out

    # -- introspection --------------------------------------------------------------------------------
    def stage_ms(self):
        out = np.zeros(_lib.NUM_STAGES)
        self._lib.gpb_stage_ms(self._h, _lib.ptr(out))
        return dict(zip(_lib.STAGE_NAMES, out.tolist()))

    def launch_count(self):
        return int(self._lib.gpb_launch_count(self._h))

    def debug_gemm(self, A, B, C, cfg=1, reps=5, alpha=1.0, beta=0.0):
        M, K = int(A.shape[0]), int(A.shape[1])
        N = int(B.shape[0])
        ms = ctypes.c_double()
        self._order(A, B, C)
        _lib.check(self._lib.gpb_debug_gemm(self._h, _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, int(cfg), float(alpha), float(beta),
                                            int(reps), ctypes.byref(ms)), self._h, "debug_gemm")
        return ms.value


    def debug_diag(self, A, mode=2, reps=1):
        """Factor one 128x128 SPD block with the diagonal-block kernel `mode` (include/gpb.h) -> (L, L^-1, sum log L_jj, ms)."""
        A = _lib.as_f64(A)
        if tuple(A.shape) != (128, 128):
            raise ValueError("A should have shape (128, 128)")
        L, W = np.empty((128, 128)), np.empty((128, 128))
        ld, ms = ctypes.c_double(), ctypes.c_double()
        _lib.check(self._lib.gpb_debug_diag(self._h, _lib.ptr(A), _lib.ptr(L), _lib.ptr(W), ctypes.byref(ld), int(mode),
                                            int(reps), ctypes.byref(ms)), self._h, "debug_diag")
        return L, W, ld.value, ms.value


def _default_device():
    """LOCAL_RANK under torchrun (one process per GPU), else the current torch device, else 0."""
    import os

    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


_shared = {}


def shared_backend(device=None):
    """Process-wide backend per device, used by the function-level API (kernels.py, gp_functions.py)."""
    if device is None:
        device = _default_device()
    be = _shared.get(device)
    if be is None:
        be = GPBackend(device)
        _shared[device] = be
    return be
