"""Where the wall time of bench.py's configs[4] leg goes: per-method time and call counts of the backend during a few BFGS
restarts at N=8192, D=16 (one stream), and the non-SPD events the optimiser ran into.
    python tools/profile_c4.py [restarts]"""
import os
import sys
import time
import warnings
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from _data import synthetic_problem
from profit_b200 import RBF, GPBackend, gp_functions as gpf

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
X, y = synthetic_problem(bench.N_TRAIN_C4, bench.DIM_C4)
starts = bench.c4_starts(64)[:R]
be = GPBackend(0)
be.set_train(X, y)
be.nll(0, bench.theta_ref(bench.DIM_C4), want_grad=True)
stats = defaultdict(lambda: [0, 0.0])
for name in ("nll", "eigsh", "nll_eig", "factorize", "set_train", "predict"):
    fn = getattr(be, name)

    def wrap(*a, _fn=fn, _name=name, **k):
        t0 = time.perf_counter()
        try:
            return _fn(*a, **k)
        finally:
            s = stats[_name]
            s[0] += 1
            s[1] += time.perf_counter() - t0

    setattr(be, name, wrap)
warnings.simplefilter("ignore")
t0 = time.perf_counter()
events = 0
for r in range(R):
    t1 = time.perf_counter()
    th, res = gpf.optimize(X, y, starts[r], RBF, eval_gradient=True, backend=be, return_result=True)
    events += res["non_spd_evaluations"]
    print(f"restart {r}: {time.perf_counter() - t1:6.2f} s, nfev {res.nfev}, nit {res.nit}, non-SPD evaluations {res['non_spd_evaluations']}, "
          f"nll {res.fun:.4f}", flush=True)
tot = time.perf_counter() - t0
print(f"total {tot:.2f} s for {R} restarts")
for k, (n, t) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:10s}: {n:5d} calls, {t:7.2f} s, {1e3 * t / max(n, 1):8.2f} ms per call")
print(f"  host (SciPy, transforms): {tot - sum(t for _, t in stats.values()):.2f} s")
