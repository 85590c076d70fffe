"""Kernel plug-ins in the reference's callable signature, evaluated on the GPU.

Mirror of profit/sur/gp/backend/python_kernels.py:8-57 (``RBF``) plus the Matern-5/2 twin the reference's
documentation promises (doc/surrogates.rst:103-106) but its Custom path never implemented.  Signature and
return shapes are the reference's:

    kernel(X, Y, length_scale=1, sigma_f=1, sigma_n=0, eval_gradient=False) -> K (n, m) [, dK (n, m, P)]

with P ordered ``[length_scale..., sigma_f, sigma_n]``.  ``kernel.gpb_kind`` selects the device kernel; the
NLL / predict paths never call these Python functions per element -- they pass the kind to libgpb.
"""
import numpy as np

from .backend import KIND_MATERN52, KIND_RBF, shared_backend


def _as_2d(a):
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(-1, 1) if a.ndim == 1 else a


def _scalar(v):
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


def _evaluate(kind, X, Y, length_scale, sigma_f, sigma_n, eval_gradient):
    X, Y = _as_2d(X), _as_2d(Y)
    ell = np.atleast_1d(np.asarray(length_scale, dtype=np.float64)).ravel()
    if ell.size not in (1, X.shape[1]):
        raise ValueError(f"length_scale must be a scalar or have {X.shape[1]} entries, got {ell.size}")
    return shared_backend().kernel_matrix(kind, X, Y, ell, _scalar(sigma_f), _scalar(sigma_n), eval_gradient)


def RBF(X, Y, length_scale=1, sigma_f=1, sigma_n=0, eval_gradient=False):
    r"""Squared-exponential kernel  K = sf^2 exp(-|x-y|^2 / (2 l^2)) + sn^2 I  (python_kernels.py:8-57)."""
    return _evaluate(KIND_RBF, X, Y, length_scale, sigma_f, sigma_n, eval_gradient)


def Matern52(X, Y, length_scale=1, sigma_f=1, sigma_n=0, eval_gradient=False):
    r"""Matern-5/2 kernel  K = sf^2 (1 + sqrt5 r + 5 r^2/3) exp(-sqrt5 r) + sn^2 I,  r = |(x-y)/l|."""
    return _evaluate(KIND_MATERN52, X, Y, length_scale, sigma_f, sigma_n, eval_gradient)


def LinearEmbedding(X, Y, R, sigma_f=1e-6, sigma_n=0, eval_gradient=False):
    r"""Linear-embedding kernel of Garnett (2014), python_kernels.py:60-114:
    K = sf^2 exp(-1/2 (x-y)^T R^T R (x-y)) + sn^2 I.  ``R`` is a (d, D) matrix or its flattening (reshaped like the
    reference does, :99); ``Y=None`` means ``Y = X`` (:101).

    ``eval_gradient=True`` returns dK of shape (n, m, R.size + 2), ordered ``[R.ravel()..., sigma_f, sigma_n]``.
    The sigma_f / sigma_n planes are the reference's (:111-112).  The R planes are the ANALYTIC derivative
    dK/dR_ab = -K_pure (R (x-y))_a (x-y)_b: the reference's ``einsum("ijk,kl", dX**2, R) * K_pure`` (:108-110) is marked
    "TODO: check" upstream, is only shape-consistent for d = 1 and is not the derivative of K there (wrong sign and
    (R delta)_a^2 R_ab in place of (R delta)_a delta_b); tests/test_gpu_parity.py checks this one by finite differences.
    """
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    R = np.asarray(R, dtype=np.float64)
    R = np.ascontiguousarray(R.reshape(R.size // X.shape[-1], X.shape[-1]))
    Yp = X if Y is None else np.atleast_2d(np.asarray(Y, dtype=np.float64))
    return shared_backend().linear_embedding(np.ascontiguousarray(X), np.ascontiguousarray(Yp), R, _scalar(sigma_f),
                                             _scalar(sigma_n), eval_gradient)


RBF.gpb_kind = KIND_RBF
Matern52.gpb_kind = KIND_MATERN52

_BY_NAME = {"RBF": RBF, "Matern52": Matern52, "Matern5/2": Matern52}   # kernels the NLL / predict paths run on device


def select_kernel(kernel):
    """'RBF' / 'Matern52' or a callable -> device-backed callable (custom_surrogate.py:213-238).

    A callable that is not one of ours is accepted if its ``__name__`` names a supported kernel (so the
    reference's own ``python_kernels.RBF`` maps to the device RBF).  Anything else raises, as the reference
    does for unknown names -- there is no host fallback.
    """
    if isinstance(kernel, str):
        name = kernel
    elif hasattr(kernel, "gpb_kind"):
        return kernel
    else:
        name = getattr(kernel, "__name__", None)
    try:
        return _BY_NAME[name]
    except KeyError:
        raise RuntimeError(f"Kernel {name} not implemented.")


def kernel_kind(kernel):
    return select_kernel(kernel).gpb_kind
