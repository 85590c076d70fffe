import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built_lib():
    """libgpb.so, built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry

    entry.build()
    return os.path.join(ROOT, "profit_b200", "libgpb.so")


@pytest.fixture(scope="session")
def backend(built_lib):
    from profit_b200 import GPBackend

    be = GPBackend(0)
    yield be
    be.close()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
