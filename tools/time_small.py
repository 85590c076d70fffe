import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from profit_b200 import GPBackend, KIND_RBF
be = GPBackend(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
X, y = synthetic_problem(N, 8)
theta = np.concatenate([np.linspace(0.6, 1.2, 8), [1.5, 0.3]])
be.set_train(X, y)
for _ in range(3): be.nll(KIND_RBF, theta)
print(be.stage_ms())
