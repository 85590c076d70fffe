"""The PyTorch-extension face of libgpb: ``torch.ops.profit_b200.*`` (profit_b200/csrc/torch_ext.cpp, built in-tree as
``_gpb_torch.so`` by ``__graft_entry__.build()``) and a small object wrapper for torch users.

The reference's plugin API is numpy (``GPSurrogate.train / predict``, ``gp_functions.*``), which ``profit_b200.surrogate``
and ``profit_b200.gp_functions`` mirror over the ctypes binding of the same C ABI.  This module is for callers that
already hold torch tensors -- candidate sets produced on the GPU, pinned host buffers -- and want operator semantics:
every op names torch's *current* CUDA stream to the library, so device tensors need no manual synchronisation, and the
outputs are torch tensors on the device of the inputs.  There is no CPU fallback: a missing extension raises.
"""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(_HERE, "_gpb_torch.so")
_loaded = False


def load():
    """Register ``torch.ops.profit_b200`` (once).  Raises if the extension has not been built."""
    global _loaded
    if _loaded:
        return torch.ops.profit_b200
    if not os.path.exists(EXT_PATH):
        raise RuntimeError(f"{EXT_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "profit_b200 has no CPU fallback.")
    torch.ops.load_library(EXT_PATH)
    _loaded = True
    return torch.ops.profit_b200


class TorchGP:
    """One device-resident GP (one gpb handle) driven through ``torch.ops.profit_b200``.

    ``theta = [length_scale..., sigma_f, sigma_n]`` is a float64 CPU tensor (read on the host, like the reference's
    ``hyp`` array); X, y and candidate sets may be CPU (ideally pinned) or CUDA tensors.
    """

    KIND = {"RBF": 0, "Matern52": 1}

    def __init__(self, device=0, kernel="RBF"):
        self.ops = load()
        self.kind = self.KIND[kernel]
        self.handle = int(self.ops.create(int(device)))
        self.n_out = 1

    def close(self):
        if getattr(self, "handle", 0):
            self.ops.destroy(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_train(self, X, y):
        self.n_out = 1 if y.dim() == 1 else int(y.shape[1])
        self.ops.set_train(self.handle, X.contiguous(), y.contiguous())

    def nll(self, theta, want_grad=True):
        """negative_log_likelihood_cholesky (gp_functions.py:103-187) on the linear scale -> (0-dim tensor, grad or None)."""
        nll, grad = self.ops.nll(self.handle, self.kind, theta.contiguous(), bool(want_grad))
        return (nll, grad) if want_grad else (nll, None)

    def factorize(self, theta):
        self.ops.factorize(self.handle, self.kind, theta.contiguous())

    def predict(self, Xpred, add_data_variance=True):
        """GPSurrogate.predict (custom_surrogate.py:95-120): (mean (M, n_out), variance (M,)) on Xpred's device."""
        return self.ops.predict(self.handle, Xpred.contiguous(), int(add_data_variance), self.n_out)

    def kernel_matrix(self, X, Y, length_scale, sigma_f, sigma_n=0.0, eval_gradient=False):
        """python_kernels.RBF (python_kernels.py:8-57) / the Matern-5/2 twin: K or (K, dK)."""
        K, dK = self.ops.kernel_matrix(self.handle, self.kind, X.contiguous(), Y.contiguous(), length_scale.contiguous(),
                                       float(sigma_f), float(sigma_n), bool(eval_gradient))
        return (K, dK) if eval_gradient else K

    def append_train(self, Xnew, ynew):
        self.ops.append_train(self.handle, Xnew.contiguous(), ynew.contiguous())

    def launch_count(self):
        return int(self.ops.launch_count(self.handle))
