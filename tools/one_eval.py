"""Exactly one NLL+gradient evaluation at N=16384 (Matern-5/2, D=10) -- the target of single-kernel ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52

N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 10
X, y = synthetic_problem(N, D)
theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
be = GPBackend(0)
be.set_train(X, y)
nll, g = be.nll(KIND_MATERN52, theta, want_grad=True)
print("nll", nll, "launches", be.launch_count())
