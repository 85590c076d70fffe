"""ctypes binding of the C ABI in include/gpb.h (libgpb.so, built in-tree by __graft_entry__.build()).

There is no CPU fallback: if the library is missing or no CUDA device is usable, every compute entry point
raises.  Status codes are mapped onto the exceptions the reference raises on this path
(np.linalg.LinAlgError for a non-SPD matrix, gp_functions.py:143/271; ValueError / RuntimeError otherwise).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpb.so")

_c_double_p = ctypes.c_void_p  # host (numpy) or device (torch.cuda) pointer, passed as a raw address
_ll = ctypes.c_longlong

# name -> (restype, argtypes); must list every symbol include/gpb.h declares (tests/test_abi.py checks this).
SIGNATURES = {
    "gpb_version": (ctypes.c_int, []),
    "gpb_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "gpb_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "gpb_error_string": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gpb_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "gpb_set_train": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _ll, ctypes.c_int, ctypes.c_int]),
    "gpb_nll": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int,
                               ctypes.POINTER(ctypes.c_double), _c_double_p]),
    "gpb_factorize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int]),
    "gpb_predict": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.c_int, _c_double_p, _c_double_p]),
    "gpb_predict_cov": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.c_int, _c_double_p, _c_double_p]),
    "gpb_argmax": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.POINTER(_ll),
                                  ctypes.POINTER(ctypes.c_double)]),
    "gpb_append_train": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _ll]),
    "gpb_acq_prepare": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, _ll, ctypes.c_int, _c_double_p,
                                       ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "gpb_acq_stats": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.c_int, ctypes.c_int, _c_double_p]),
    "gpb_acquire": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _ll,
                                   ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.POINTER(_ll), ctypes.c_int,
                                   _c_double_p, ctypes.POINTER(_ll), ctypes.POINTER(ctypes.c_double), _c_double_p]),
    "gpb_kernel_matrix": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, _ll, _c_double_p, _ll,
                                         ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                         _c_double_p, _c_double_p]),
    "gpb_linear_embedding": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, _c_double_p, _ll, ctypes.c_int, _c_double_p,
                                            ctypes.c_int, ctypes.c_double, ctypes.c_double, _c_double_p, _c_double_p]),
    "gpb_marginal_variance": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, _c_double_p, _c_double_p, _c_double_p]),
    "gpb_mean_pairwise_distance": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.c_int,
                                                  ctypes.POINTER(ctypes.c_double)]),
    "gpb_cholesky": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, _c_double_p]),
    "gpb_invert_cholesky": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, _c_double_p]),
    "gpb_solve_cholesky": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, _c_double_p, ctypes.c_int, _c_double_p]),
    "gpb_eigsh": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 _c_double_p, ctypes.POINTER(ctypes.c_int)]),
    "gpb_nll_eig": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                   _c_double_p]),
    "gpb_eigsh_matrix": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _ll, ctypes.c_int, ctypes.c_int, _c_double_p,
                                        _c_double_p, ctypes.POINTER(ctypes.c_int)]),
    "gpb_eig_inverse": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p]),
    "gpb_stage_ms": (ctypes.c_int, [ctypes.c_void_p, _c_double_p]),
    "gpb_launch_count": (_ll, [ctypes.c_void_p]),
    "gpb_debug_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, _ll, ctypes.c_void_p, _ll,
                                      ctypes.POINTER(_ll)]),
    "gpb_debug_gemm": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p, _ll, _ll, _ll,
                                      ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_double)]),
    "gpb_debug_diag": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p,
                                      ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_double)]),
}

NUM_STAGES = 8
STAGE_NAMES = ("assemble", "potrf", "trtri", "syrk", "grad", "cross", "var", "total")

_lib = None


def load():
    """Load libgpb.so (once) and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "profit_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(x):
    """Raw address of a C-contiguous float64 numpy array or torch tensor (CPU or CUDA); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a C-contiguous float64 array")
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch.Tensor
        import torch

        if x.dtype != torch.float64 or not x.is_contiguous():
            raise ValueError("expected a contiguous float64 tensor")
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError(f"cannot take the address of {type(x)!r}")


def is_cuda_tensor(x):
    return x is not None and not isinstance(x, np.ndarray) and hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def as_f64(a):
    """numpy view/copy in the layout the ABI wants (C-order float64); torch tensors pass through."""
    if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def check(status, handle=None, what=""):
    """Map a gpb status code onto the reference's exception types."""
    if status == 0:
        return
    msg = ""
    if handle is not None and _lib is not None:
        raw = _lib.gpb_error_string(handle)
        msg = raw.decode() if raw else ""
    if status > 0:
        raise np.linalg.LinAlgError(msg or f"{what}: matrix is not positive definite (info={status})")
    if status == -1:
        raise ValueError(msg or f"{what}: bad argument")
    if status == -4:   # scipy raises ArpackNoConvergence (a RuntimeError) when eigsh does not converge
        raise RuntimeError(msg or f"{what}: the device eigen-solver did not converge")
    raise RuntimeError(msg or f"{what}: CUDA backend failure (status {status})")
