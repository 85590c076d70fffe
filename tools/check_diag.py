"""Diagonal-block kernels of the Cholesky on one 128x128 SPD block: correctness against numpy and device time of
mode 3 (blocked DMMA kernel as on the panel chain: factor + 32x32 diagonal inverses, completed off the chain), 2 (the
same with the full inverse), 1 (rank-4 register sweep) and 0 (rank-1 register sweep)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from profit_b200 import GPBackend

be = GPBackend(0)
out = []
rng = np.random.default_rng(3)
for case in ("wishart", "kernel", "identity"):
    if case == "wishart":
        M = rng.standard_normal((128, 128))
        A = M @ M.T + 128 * np.eye(128)
    elif case == "kernel":
        X = rng.random((128, 3))
        d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
        A = 2.25 * np.exp(-0.5 * d2 / 0.49) + 0.09 * np.eye(128)
    else:
        A = np.eye(128)
    Lref = np.linalg.cholesky(A)
    Wref = np.linalg.inv(Lref)
    for mode in (3, 2, 1, 0):
        L, W, ld, ms = be.debug_diag(np.tril(A), mode=mode, reps=20)
        rec = {"case": case, "mode": mode, "us": round(ms * 1e3, 2),
               "L_err": float(np.abs(L - Lref).max() / np.abs(Lref).max()),
               "W_err": float(np.abs(W - Wref).max() / np.abs(Wref).max()),
               "WL_minus_I": float(np.abs(W @ L - np.eye(128)).max()),
               "logdet_err": float(abs(ld - np.log(np.diag(Lref)).sum())),
               "upper_zero": bool(np.all(np.triu(L, 1) == 0) and np.all(np.triu(W, 1) == 0))}
        print(json.dumps(rec), flush=True)
        out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check_diag.json", "w"), indent=1)
