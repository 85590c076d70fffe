import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
from profit_b200 import GPBackend, KIND_RBF
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
be = GPBackend(0)
X, y = synthetic_problem(N, 8)
theta = np.concatenate([np.linspace(0.6, 1.2, 8), [1.5, 0.3]])
K = be.kernel_matrix(KIND_RBF, X, X, theta[:8], 1.5, 0.3)
Lt = torch.linalg.cholesky(torch.from_numpy(K).cuda()).cpu().numpy()
for rep in range(4):
    try:
        L = be.cholesky(K)
    except Exception as e:
        print("rep", rep, "FAILED:", e); continue
    err = np.abs(L - Lt)
    nb = N // 128
    te = err.reshape(nb, 128, nb, 128).max(axis=(1, 3))
    bad = np.argwhere(te > 1e-9)
    print("rep", rep, "max err", err.max(), "bad tiles", len(bad), "first", bad[:8].tolist(), flush=True)
