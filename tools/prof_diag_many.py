import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from profit_b200 import GPBackend
be = GPBackend(0)
n = 8192
rng = np.random.default_rng(0)
L = np.tril(rng.standard_normal((n, n)) * 0.01) + np.eye(n) * 2.0
be.invert_cholesky(L)
print("done")
