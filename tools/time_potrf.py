import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_RBF
be = GPBackend(0)
for (N, D) in ((1024, 8), (2048, 8), (4096, 8), (8192, 16), (16384, 10)):
    X, y = synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    be.set_train(X, y)
    be.nll(KIND_RBF, theta, want_grad=True)
    best = None
    for _ in range(3):
        be.nll(KIND_RBF, theta, want_grad=True)
        st = be.stage_ms()
        if best is None or st["total"] < best["total"]: best = st
    print(f"N={N:6d}: potrf {best['potrf']:7.3f} trtri {best['trtri']:7.3f} syrk {best['syrk']:7.3f} grad {best['grad']:6.3f} total {best['total']:8.3f} ms", flush=True)
