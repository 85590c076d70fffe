"""The C-ABI library: builds, loads without a GPU, exports exactly what include/gpb.h declares, and fails loudly
(no CPU fallback) when no CUDA device is usable.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "gpb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = _header_symbols()
    for must in ("gpb_create", "gpb_destroy", "gpb_set_train", "gpb_nll", "gpb_factorize", "gpb_predict",
                 "gpb_kernel_matrix", "gpb_cholesky", "gpb_invert_cholesky", "gpb_solve_cholesky", "gpb_argmax",
                 "gpb_append_train", "gpb_acq_prepare", "gpb_acquire", "gpb_predict_cov", "gpb_marginal_variance"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} declared in gpb.h but not exported by libgpb.so"
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (gpb_[a-z_0-9]+)", out))
    assert exported == set(_header_symbols()), "exported C symbols and gpb.h differ"


def test_python_binding_covers_the_abi(built_lib):
    from profit_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _header_symbols()
    lib = _lib.load()
    assert lib.gpb_version() >= 100


def test_library_targets_sm100a_with_tma_and_dmma(built_lib):
    """SASS evidence that the hot kernels are Blackwell-native: bulk-tensor TMA loads and FP64 tensor MMAs."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass and "DMMA.8x8x4" in sass and "SYNCS" in sass


def test_no_cpu_fallback_without_a_device(built_lib):
    """Without a GPU the product path must raise, not silently compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from profit_b200 import GPBackend, RBF, gp_functions

    with pytest.raises(RuntimeError):
        GPBackend(0)
    X = np.random.default_rng(0).random((8, 2))
    with pytest.raises(RuntimeError):
        RBF(X, X, 1.0, 1.0, 0.1)
    with pytest.raises(RuntimeError):
        gp_functions.negative_log_likelihood_cholesky(np.array([1.0, 1.0, 0.1]), X, X[:, 0], RBF)


def test_product_code_never_imports_the_oracle():
    """Only tests/ (incl. tests/tools/), __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: the package,
    the developer tools and the examples must not."""
    for sub in ("profit_b200", "tools", "examples", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from\s+oracle\b|import\s+[^#\n]*\boracle\b)", text, flags=re.M), \
                        f"{sub}/{f} imports the oracle"
    # bench.py: only inside the CPU-baseline / reference-arm functions; __graft_entry__: only inside smoke()
    for name, allowed in (("bench.py", ("cpu_nll_grad_eval", "cpu_nll_grad_sample", "cpu_predict_sample", "run_reference")),
                          ("__graft_entry__.py", ("smoke",))):
        text = open(os.path.join(ROOT, name)).read()
        assert not re.search(r"^(import|from)\s+oracle\b", text, flags=re.M), f"{name} imports the oracle at module level"
        for m in re.finditer(r"^\s+(?:import|from)\s+oracle\b", text, flags=re.M):
            head = text[:m.start()]
            func = re.findall(r"^def\s+(\w+)", head, flags=re.M)[-1]
            assert func in allowed, f"{name}: oracle imported in {func}()"
