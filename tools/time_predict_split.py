"""Predict stage split (cross kernel + mean vs variance GEMM) over training-set sizes, Matern-5/2, D=10."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52

D = 10
be = GPBackend(0)
for N, M in ((256, 1 << 20), (512, 1 << 20), (1024, 1 << 19), (2048, 1 << 18), (4096, 1 << 17), (8192, 1 << 16)):
    X, y = synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    be.set_train(X, y)
    be.factorize(KIND_MATERN52, theta)
    Xc = torch.rand((M, D), dtype=torch.float64, device="cuda:0")
    mean = torch.empty((M, 1), dtype=torch.float64, device="cuda:0")
    var = torch.empty(M, dtype=torch.float64, device="cuda:0")
    best = None
    for _ in range(3):
        be.predict(Xc, out_mean=mean, out_var=var)
        st = be.stage_ms()
        if best is None or st["total"] < best["total"]:
            best = st
    Np = (N + 127) // 128 * 128
    gemm_tf = M * Np * (Np + 128) / (best["var"] * 1e-3) / 1e12     # lower-triangular W: ~Np^2 flop per candidate
    print(f"N={N:5d} M={M:8d}: cross {best['cross']:8.3f} ms, var {best['var']:8.3f} ms, total {best['total']:8.3f} ms "
          f"-> {M / best['total'] * 1e3:12.0f} cand/s, variance GEMM {gemm_tf:5.1f} TFLOP/s", flush=True)
