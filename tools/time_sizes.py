"""Stage times of one NLL+gradient evaluation over a range of sizes (best of 5 after warm-up), with the share of the FP64
peak that N^3 flop / total corresponds to -- the evidence behind DESIGN.md section 3.2 / 6b (mid sizes are chain-bound).
    python tools/time_sizes.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52, KIND_RBF

sizes = [int(a) for a in sys.argv[1:]] or [200, 512, 1024, 2048, 4096, 8192, 16384]
be = GPBackend(0)
D = int(os.environ.get("TS_D", "8"))                      # TS_D / TS_KIND: input dimension and kernel of the sweep
KIND = KIND_MATERN52 if os.environ.get("TS_KIND", "rbf").lower().startswith("m") else KIND_RBF
theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
print(f"{'N':>6} {'assemble':>9} {'potrf':>8} {'trtri':>8} {'syrk':>8} {'grad':>7} {'total':>9} {'TFLOP/s':>8} {'launches':>8}  (ms; {'Matern-5/2' if KIND == KIND_MATERN52 else 'RBF'} ARD, D={D})")
for N in sizes:
    X, y = synthetic_problem(N, D)
    be.set_train(X, y)
    be.nll(KIND, theta, want_grad=True)
    best = None
    for _ in range(5):
        l0 = be.launch_count()
        be.nll(KIND, theta, want_grad=True)
        st = be.stage_ms()
        st["launches"] = be.launch_count() - l0
        if best is None or st["total"] < best["total"]:
            best = st
    tf = float(N) ** 3 / (best["total"] * 1e-3) * 1e-12
    print(f"{N:6d} {best['assemble']:9.3f} {best['potrf']:8.3f} {best['trtri']:8.3f} {best['syrk']:8.3f} {best['grad']:7.3f} "
          f"{best['total']:9.3f} {tf:8.2f} {best['launches']:8d}", flush=True)
