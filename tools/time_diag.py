import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from profit_b200 import GPBackend
be = GPBackend(0)
rng = np.random.default_rng(0)
Q = rng.standard_normal((128, 128)); A = Q @ Q.T / 128 + np.eye(128)
L = np.linalg.cholesky(A)
for _ in range(5):
    be.cholesky(A); be.invert_cholesky(L)
print("done")
