"""profit_b200 -- B200-native (sm_100a) backend for proFit's Custom GP hot path.

Host-side mirror of the reference interface (profit.sur.gp: ``GPSurrogate``, ``gp_functions``,
``python_kernels``) over the C ABI in include/gpb.h.  All numerics run in hand-written CUDA kernels in
libgpb.so; there is no CPU fallback.
"""
from . import gp_functions, kernels  # noqa: F401
from .backend import GPBackend, KIND_MATERN52, KIND_RBF, shared_backend  # noqa: F401
from .kernels import RBF, LinearEmbedding, Matern52, select_kernel  # noqa: F401
from .surrogate import B200GPSurrogate, B200MultiOutputGPSurrogate  # noqa: F401

__all__ = ["GPBackend", "B200GPSurrogate", "B200MultiOutputGPSurrogate", "RBF", "Matern52", "LinearEmbedding", "select_kernel", "gp_functions", "kernels",
           "shared_backend", "KIND_RBF", "KIND_MATERN52"]
