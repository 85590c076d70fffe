"""BASELINE configs[1]: a complete hyper-parameter fit (RBF, N=4096, D=8, fp64) on one GPU -- the reference's BFGS
call and the L-BFGS-B driver the config names -- with evaluation counts and wall time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from _data import synthetic_problem
from profit_b200 import RBF, GPBackend, gp_functions as gpf

N, D = 4096, 8
X, y = synthetic_problem(N, D)
theta_ref = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
a0 = theta_ref * 10 ** (0.3 * (np.random.default_rng(5).random(D + 2) - 0.5))
be = GPBackend(0)
out = {"config": "RBF N=4096 D=8 fp64, start = reference point * 10^U(-0.15, 0.15)", "runs": []}
for method, bounds in (("bfgs", None), ("L-BFGS-B", [(-2.0, 2.0)] * (D + 1) + [(-4.0, 0.0)])):
    gpf.optimize(X[:256], y[:256], a0, RBF, eval_gradient=True, backend=be)      # warm-up (workspaces, SciPy import)
    t0 = time.perf_counter()
    hyp, res = gpf.optimize(X, y, a0, RBF, eval_gradient=True, backend=be, method=method, bounds=bounds,
                            return_result=True)
    dt = time.perf_counter() - t0
    nll_end = gpf.negative_log_likelihood_cholesky(hyp, X, y, RBF)
    nll_start = gpf.negative_log_likelihood_cholesky(a0, X, y, RBF)
    rec = {"method": method, "seconds": round(dt, 3), "nfev": int(res.nfev), "nit": int(res.nit),
           "ms_per_eval_incl_host": round(1e3 * dt / res.nfev, 3), "nll_start": nll_start, "nll_end": nll_end,
           "success": bool(res.success), "hyp": hyp.tolist()}
    print(json.dumps(rec), flush=True)
    out["runs"].append(rec)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/c1_fit.json", "w") as f:
    json.dump(out, f, indent=1)
