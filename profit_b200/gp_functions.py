"""Function-level twins of profit/sur/gp/backend/gp_functions.py, backed by libgpb (CUDA, sm_100a).

Same names, argument meaning and error behaviour as the reference:

* ``optimize``                          gp_functions.py:8-69    (SciPy BFGS on the host, objective on the GPU)
* ``solve_cholesky``                    gp_functions.py:72-100
* ``negative_log_likelihood_cholesky``  gp_functions.py:103-187
* ``negative_log_likelihood``           gp_functions.py:190-301 (Cholesky branch; see note)
* ``invert_cholesky`` / ``invert``      gp_functions.py:304-373
* ``marginal_variance_BBQ``             gp_functions.py:376-454
* ``predict_f``                         gp_functions.py:457-491

``negative_log_likelihood`` (the objective handed to SciPy) and the reference's eigen-decomposition exits.  The reference
runs eigsh passes with a doubling ``neig`` and leaves its loop on a TRUNCATED EIGEN-DECOMPOSITION either when the Cholesky
factorisation fails or -- without ever trying it -- when the ``neig``-th eigenvalue of Ky drops below
``clip_eig = max(1e-3 min|l|, 1e-10)`` inside the first 0.05 N eigenvalues (:254-301; only possible when
sigma_n^2 <= clip_eig).  That value is not the Cholesky NLL (tests/golden/reference_nonspd.json: +19.37 against -2517.99
on a perfectly SPD 1000 x 1000 matrix).  ``NON_SPD_POLICY`` (per call: ``on_non_spd``) selects what happens:

* ``"reference"`` (default): the reference's loop, branch for branch, with the eigenpairs from the device solver
  (``gpb_eigsh`` / ``gpb_nll_eig``: Lanczos with full re-orthogonalisation, deterministic).  The eigsh passes are skipped
  when they provably cannot change the outcome (sigma_n^2 > clip_eig: every eigenvalue of Ky is at least sigma_n^2), so the
  usual evaluation is exactly ``negative_log_likelihood_cholesky``.  On the eig exit the value agrees with the reference
  to the reference's own reproducibility (its ARPACK call starts from a random vector and stops at tol = clip_eig: three
  runs scatter by 5e-7 relative); a ``NonSPDWarning`` is emitted and the event is recorded on the backend.
* ``"jitter"``: whenever K does not factorise, retry on the device with a growing diagonal jitter (an inflated sigma_n),
  ``NonSPDWarning`` + recorded event; never takes the eig exit.
* ``"raise"``: ``np.linalg.LinAlgError`` propagates.
"""
import os
import warnings

import numpy as np

from .backend import shared_backend
from .kernels import kernel_kind


class NonSPDWarning(RuntimeWarning):
    """The covariance matrix of an NLL evaluation was not positive definite and a jittered one was used instead."""


# "reference" | "jitter" | "raise"; per call: negative_log_likelihood(..., on_non_spd=...); process-wide default from the
# environment (PROFIT_B200_NON_SPD=jitter gives the Cholesky NLL -- row a3 of SURVEY.md section 8 -- wherever K factorises)
NON_SPD_POLICY = os.environ.get("PROFIT_B200_NON_SPD", "reference")
JITTER_LADDER = (1e-10, 1e-8, 1e-6, 1e-4, 1e-2)   # relative to sigma_f^2


def _train_arrays(X, y):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    y = np.asarray(y, dtype=np.float64)
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def solve_cholesky(L, b):
    r"""alpha = L^T \ (L \ b)  (gp_functions.py:72-100)."""
    return shared_backend().solve_cholesky(np.asarray(L, dtype=np.float64), np.asarray(b, dtype=np.float64))


def invert_cholesky(L):
    r"""(L L^T)^{-1} from the lower Cholesky factor (gp_functions.py:304-322)."""
    return shared_backend().invert_cholesky(np.asarray(L, dtype=np.float64))


def invert(K, neig=0, tol=1e-10, eps=1e-6, max_iter=1000):
    """Inverse of a symmetric positive-definite matrix, gp_functions.py:325-373: through the Cholesky factorisation
    (retried once with ``eps`` on the diagonal, :340-343) and, for ``0 < neig <= 0.05 len(K)``, the reference's doubling
    eigsh loop that may leave on ``Q diag(1/w) Q^T`` (:350-369), with the eigenpairs from the device solver."""
    be = shared_backend()
    K = np.asarray(K, dtype=np.float64)
    try:
        L = be.cholesky(K)                                       # :341
    except np.linalg.LinAlgError:
        L = be.cholesky(K + eps * np.eye(K.shape[0], K.shape[1]))    # :343
    if neig <= 0 or neig > 0.05 * len(K):                        # :344-348
        return be.invert_cholesky(L)
    n = len(K)
    neig = int(neig)
    w, _ = be.eigsh_matrix(K, neig)                              # :351
    for _iteration in range(max_iter):                           # :352
        if neig > 0.05 * n:                                      # :353-357
            return be.invert_cholesky(L)
        neig = 2 * neig                                          # :358
        w, _ = be.eigsh_matrix(K, neig, reuse=True)              # :359
        if np.abs(w[0] - tol) <= tol or neig >= n:               # :362-363
            break
    w = np.array(w)
    w[(-1e-2 < w) & (w < 0)] = 1e-2                              # :366-367
    return be.eig_inverse(w, n)                                  # :368-369


def negative_log_likelihood_cholesky(hyp, X, y, kernel, eval_gradient=False, log_scale_hyp=False,
                                     fixed_sigma_n=False, backend=None):
    """NLL (and gradient) through the Cholesky factorisation (gp_functions.py:103-187).

    ``hyp`` is on the linear scale, ``[length_scale..., sigma_f, sigma_n]``.  Returns ``float`` or
    ``(float, ndarray)``; raises ``np.linalg.LinAlgError`` if the covariance matrix is not SPD.
    ``backend`` (optional) is a GPBackend that already holds (X, y); otherwise they are uploaded now.
    """
    hyp = np.asarray(hyp, dtype=np.float64)
    kind = kernel_kind(kernel)
    if backend is None:
        backend = shared_backend()
        Xc, yc = _train_arrays(X, y)
        backend.set_train(Xc, yc)
    if not eval_gradient:
        return backend.nll(kind, hyp, want_grad=False)
    nll, dnll = backend.nll(kind, hyp, want_grad=True)
    if fixed_sigma_n:                      # :181-183
        dnll = dnll[:-1]
        hyp = hyp[:-1]
    if log_scale_hyp:                      # :184-186, chain rule
        dnll = dnll * (hyp * np.log(10))
    return nll, dnll


def negative_log_likelihood(hyp, X, y, kernel, eval_gradient=False, log_scale_hyp=False,
                            fixed_sigma_n_value=None, neig=0, max_iter=1000, backend=None, on_non_spd=None):
    """Objective handed to the optimiser (gp_functions.py:190-301): ``hyp`` on the log10 scale if ``log_scale_hyp``; a
    fixed sigma_n is appended before evaluation; ``neig`` / ``max_iter`` steer the eigsh passes as in the reference.

    ``on_non_spd`` (default ``NON_SPD_POLICY``, see the module docstring): ``"reference"`` follows the reference's loop
    including its truncated-eigen-decomposition exit; ``"jitter"`` retries a failed factorisation with
    K + j sigma_f^2 I for j in ``JITTER_LADDER``; ``"raise"`` propagates ``np.linalg.LinAlgError``.  The eig exit and the
    jitter retry emit a ``NonSPDWarning`` and append an event to ``backend.non_spd_events``.
    """
    policy = on_non_spd or NON_SPD_POLICY
    if policy not in ("reference", "jitter", "raise"):
        raise ValueError(f"on_non_spd must be 'reference', 'jitter' or 'raise', got {policy!r}")
    hyp = np.asarray(hyp, dtype=np.float64)
    if log_scale_hyp:
        hyp = 10**hyp                      # :239-240
    fixed_sigma_n = fixed_sigma_n_value is not None
    if fixed_sigma_n:
        hyp = np.append(hyp, fixed_sigma_n_value)   # :241-243
    if backend is None:
        backend = shared_backend()
        Xc, yc = _train_arrays(X, y)
        backend.set_train(Xc, yc)
    if policy == "reference":
        return _nll_reference_loop(hyp, X, y, kernel, eval_gradient, log_scale_hyp, fixed_sigma_n, neig, max_iter, backend)
    try:
        return negative_log_likelihood_cholesky(hyp, X, y, kernel, eval_gradient=eval_gradient,
                                                log_scale_hyp=log_scale_hyp, fixed_sigma_n=fixed_sigma_n,
                                                backend=backend)
    except np.linalg.LinAlgError:
        if policy == "raise":
            raise
    sigma_f, sigma_n = float(hyp[-2]), float(hyp[-1])
    for jitter in JITTER_LADDER:
        inflated = np.array(hyp, dtype=np.float64)
        inflated[-1] = np.sqrt(sigma_n**2 + jitter * sigma_f**2)
        try:
            out = negative_log_likelihood_cholesky(inflated, X, y, kernel, eval_gradient=eval_gradient,
                                                   log_scale_hyp=False, fixed_sigma_n=False, backend=backend)
        except np.linalg.LinAlgError:
            continue
        backend.non_spd_events.append({"jitter": jitter, "hyp": hyp.copy()})
        warnings.warn(NonSPDWarning(
            f"covariance matrix not positive definite at hyp={hyp.tolist()}; NLL evaluated with diagonal jitter "
            f"{jitter:g} * sigma_f^2 (the reference would switch to a truncated eigen-decomposition here)"), stacklevel=2)
        if not eval_gradient:
            return out
        nll, dnll = out
        dnll = np.array(dnll)
        dnll[-1] *= sigma_n / inflated[-1]          # d sigma_n' / d sigma_n
        dnll[-2] += out[1][-1] * jitter * sigma_f / inflated[-1]   # d sigma_n' / d sigma_f
        if fixed_sigma_n:
            dnll, hyp = dnll[:-1], hyp[:-1]
        if log_scale_hyp:
            dnll = dnll * (hyp * np.log(10))
        return nll, dnll
    raise np.linalg.LinAlgError("covariance matrix is not positive definite even with diagonal jitter")


def _nll_reference_loop(hyp, X, y, kernel, eval_gradient, log_scale_hyp, fixed_sigma_n, neig, max_iter, backend):
    """The loop of gp_functions.py:254-301 on the linear-scale ``hyp`` (sigma_n already appended).

    eigsh -> ``backend.eigsh`` (device Lanczos; passes with growing ``neig`` continue one run), the Cholesky attempt ->
    ``negative_log_likelihood_cholesky``, the exit -> ``backend.nll_eig``.  An eigsh pass whose outcome is known in advance
    is not run: every eigenvalue of Ky = K + sigma_n^2 I is at least sigma_n^2 (K is positive semi-definite for the RBF and
    Matern kernels), so ``w[0] <= clip_eig`` (:280) cannot hold while sigma_n^2 exceeds clip_eig by more than the rounding
    of the spectrum -- until a factorisation has failed, which shows that the numerical spectrum is not what the exact
    one promises.
    """
    kind = kernel_kind(kernel)
    n = backend.n
    clip_eig = max(1e-3 * min(abs(hyp[:-2])), 1e-10)             # :245
    sf2, sn2 = float(hyp[-2]) ** 2, float(hyp[-1]) ** 2
    spectrum_rounding = 64 * np.finfo(float).eps * (n * sf2 + sn2)      # trace(Ky) bounds the largest eigenvalue
    eigenvalues_matter = not (sn2 - spectrum_rounding > clip_eig) or max_iter <= 0
    neig = max(int(neig), 1)                                     # :256
    iteration = 0
    failed = 0
    w = None
    k_hint = neig                                                # the largest pass before the first Cholesky attempt
    while 2 * k_hint <= 0.05 * n:
        k_hint *= 2
    while True:                                                  # :257
        if not neig or neig > 0.05 * n:                          # :258-272
            try:
                return negative_log_likelihood_cholesky(hyp, X, y, kernel, eval_gradient=eval_gradient,
                                                        log_scale_hyp=log_scale_hyp, fixed_sigma_n=fixed_sigma_n,
                                                        backend=backend)
            except np.linalg.LinAlgError:
                failed += 1                                      # the reference prints "Warning! Fallback to eig solver!"
                eigenvalues_matter = True
        k_used = min(neig, n)
        # Krylov-space hint: enough for the next two doublings, not for the largest pass at once -- the loop usually
        # leaves at a small neig (a start that wandered to huge length scales leaves at neig = 1), and the device run
        # is continued, not restarted, if a later pass needs more (N = 8192: 215 -> ~15 ms for such an evaluation)
        hint = min(k_hint, max(32, 4 * k_used))
        w = backend.eigsh(kind, hyp, k_used, hint) if eigenvalues_matter else None        # :273-275
        if iteration >= max_iter:                                # :276-278 (``iteration`` is never incremented upstream)
            break
        neig *= 2                                                # :279
        if (w is not None and w[0] <= clip_eig) or neig >= n:    # :280
            break
    if w is None:                                                # left by ``neig >= n`` with the passes skipped
        w = backend.eigsh(kind, hyp, k_used)
    backend.non_spd_events.append({"branch": "eig", "k": int(k_used), "neig": int(neig), "failed_cholesky": failed,
                                   "hyp": hyp.copy()})
    warnings.warn(NonSPDWarning(
        f"negative_log_likelihood left the Cholesky branch at hyp={hyp.tolist()}: NLL from the {k_used} largest "
        f"eigenpairs (gp_functions.py:283-301), {failed} failed factorisation(s), smallest kept eigenvalue {w[0]:.3e}"),
        stacklevel=3)
    neig_term = min(neig, n)                                     # :286
    if not eval_gradient:
        return backend.nll_eig(neig_term)
    try:
        nll, dnll = backend.nll_eig(neig_term, want_grad=True, n_params=hyp.size)   # :291-294
    except np.linalg.LinAlgError as exc:         # invert(Ky) failed even with the diagonal shifts (the reference raises too)
        raise np.linalg.LinAlgError(f"{exc}; eig-branch gradient at hyp={hyp.tolist()}") from None
    if fixed_sigma_n:                                            # :295-297
        dnll = dnll[:-1]
        hyp = hyp[:-1]
    if log_scale_hyp:                                            # :298-300
        dnll = dnll * (hyp * np.log(10))
    return nll, dnll


def optimize(xtrain, ytrain, a0, kernel, fixed_sigma_n=False, eval_gradient=False, return_hess_inv=False,
             backend=None, method="bfgs", bounds=None, options=None, return_result=False, resident_key=None):
    """Hyper-parameter optimisation, gp_functions.py:8-69: log10 transform, SciPy ``minimize(method="bfgs")``
    with the same arguments; the training set is uploaded once and stays resident across evaluations.

    ``method`` / ``bounds`` / ``options`` (extensions, defaults = the reference's call at :59-66) select another SciPy
    driver, e.g. ``method="L-BFGS-B"`` with log10 bounds as in the reference's own test (test_kernels.py:188-194);
    ``return_result`` returns the SciPy result object as the last element (with ``non_spd_evaluations`` = how many
    objective evaluations needed the jitter retry, see ``negative_log_likelihood``)."""
    from scipy.optimize import minimize

    a0 = np.asarray(a0, dtype=np.float64)
    if fixed_sigma_n:
        sigma_n = a0[-1]
        a0 = a0[:-1]
    else:
        sigma_n = None
    a0 = np.log10(a0)

    if backend is None:
        backend = shared_backend()
    Xc, yc = _train_arrays(xtrain, ytrain)
    # ``resident_key`` (set by B200GPSurrogate): the caller's fingerprint of (xtrain, ytrain); when the device already
    # holds exactly this training set the upload is skipped
    if resident_key is None or backend.resident_key != resident_key:
        backend.set_train(Xc, yc)
        backend.resident_key = resident_key
    backend.non_spd_events = []          # what this optimisation run sees (read by B200GPSurrogate.optimize)

    def objective(h):
        return negative_log_likelihood(h, Xc, yc, kernel, eval_gradient, True, sigma_n, backend=backend)

    opt_result = minimize(objective, a0, method=method, bounds=bounds, jac=eval_gradient, options=options)
    out = (10**opt_result.x,)
    if return_hess_inv:
        hess_inv = opt_result.hess_inv
        if hasattr(hess_inv, "todense"):          # L-BFGS-B returns a LinearOperator
            hess_inv = np.asarray(hess_inv.todense())
        out += (hess_inv,)
    if return_result:
        opt_result["non_spd_evaluations"] = len(backend.non_spd_events)
        out += (opt_result,)
    return out[0] if len(out) == 1 else out


def marginal_variance_BBQ(Xtrain, ytrain, Xpred, kernel, hyperparameters, hess_inv, fixed_sigma_n=False, alpha=None,
                          predictive_variance=0, backend=None):
    """Marginal variance for active learning after Osborne (2012), gp_functions.py:376-454:
    ``dm H^-1 dm^T`` (diagonal) + predictive variance, with ``dm`` the derivative of the predictive mean w.r.t.
    [length_scale..., sigma_f, sigma_n].  Returns an (M, 1) array.

    ``alpha`` is accepted for signature compatibility and ignored (the device factor supplies it).  With a vector
    length scale the derivative has one entry per length scale -- the reference indexes ``dKy[..., 0]`` and
    ``dKy[..., 1]`` as (l, sigma_f) (:439-448) and is only meaningful for a scalar length scale, where both agree.
    """
    ordered = [np.atleast_1d(np.asarray(hyperparameters[key], dtype=np.float64)).ravel()
               for key in ("length_scale", "sigma_f", "sigma_n")]
    predictive_variance = np.asarray(predictive_variance, dtype=np.float64)
    if hess_inv is None:                                   # :406-408
        return predictive_variance.reshape(-1, 1)
    theta = np.concatenate(ordered)
    P = theta.size
    hess_inv = np.asarray(hess_inv, dtype=np.float64)
    if fixed_sigma_n:                                      # :410-412, padded to the full parameter count
        padded = np.zeros((P, P))
        padded[:P - 1, :P - 1] = hess_inv
        hess_inv = padded
    Xp = np.asarray(Xpred, dtype=np.float64)
    if Xp.ndim == 1:
        Xp = Xp.reshape(-1, 1)
    be = backend
    if be is None:
        be = shared_backend()
        Xc, yc = _train_arrays(Xtrain, ytrain)
        be.set_train(Xc, yc)
    if be.factor_key is None:
        be.factorize(kernel_kind(kernel), theta)
    pv = np.broadcast_to(predictive_variance.reshape(-1), (Xp.shape[0],)) if predictive_variance.size else None
    mv = be.marginal_variance(np.ascontiguousarray(Xp), hess_inv, pv)
    return mv.reshape(-1, 1)


def predict_f(hyp, x, y, xtest, kernel, return_full_cov=False, neig=0):
    """Posterior mean and variance from a flat hyper-parameter array (gp_functions.py:457-491).

    The diagonal of K** carries sigma_n^2 (:484) and the clamp is applied after it (:488), element-wise on the whole
    matrix when ``return_full_cov`` is set.
    """
    hyp = np.asarray(hyp, dtype=np.float64)
    be = shared_backend()
    Xc, yc = _train_arrays(x, y)
    be.set_train(Xc, yc)
    be.factorize(kernel_kind(kernel), hyp)
    xt = np.asarray(xtest, dtype=np.float64)
    if xt.ndim == 1:
        xt = xt.reshape(-1, 1)
    if return_full_cov:
        mean, vstar = be.predict_cov(np.ascontiguousarray(xt), noise_on_diagonal=True)
    else:
        mean, vstar = be.predict(np.ascontiguousarray(xt), add_data_variance=2)   # sn^2 joins before the clamp
    if yc.ndim == 1:
        mean = mean.reshape(-1)
    return mean, vstar
