"""Cost of following the reference's eigsh passes (gp_functions.NON_SPD_POLICY = "reference") against the plain Cholesky
objective ("jitter") on the active-learning demo problem, warm (second fit of each policy): fit seconds, objective
evaluations, evaluations that left on the eig branch, final hyper-parameters."""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from profit_b200 import B200GPSurrogate, gp_functions as gpf

warnings.simplefilter("ignore")
n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(0)
X = rng.random((n0, 3))
y = (np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + 0.5 * X[:, 2]).reshape(-1, 1) + 0.02 * rng.standard_normal((n0, 1))
B200GPSurrogate().train(X, y, kernel="Matern52")          # context, workspaces, SciPy
for policy in ("reference", "jitter", "reference", "jitter"):
    gpf.NON_SPD_POLICY = policy
    gp = B200GPSurrogate()
    calls = {"n": 0, "eig": 0, "eig_s": 0.0}
    be = gp.backend
    orig_eigsh, orig_nll_eig, orig_nll = be.eigsh, be.nll_eig, be.nll

    def timed(fn, key):
        def wrapper(*a, **k):
            t = time.perf_counter()
            out = fn(*a, **k)
            calls["eig_s"] += time.perf_counter() - t
            calls[key] = calls.get(key, 0) + 1
            return out
        return wrapper

    be.eigsh, be.nll_eig = timed(orig_eigsh, "eigsh_calls"), timed(orig_nll_eig, "eig")
    t0 = time.perf_counter()
    gp.train(X, y, kernel="Matern52")
    dt = time.perf_counter() - t0
    print(f"{policy:9s}: fit {dt:6.3f} s, eig exits {calls['eig']}, eigsh calls {calls.get('eigsh_calls', 0)}, "
          f"{calls['eig_s']:.3f} s inside the eig path, non_spd_evaluations {gp.non_spd_evaluations}, "
          f"hyp { {k: np.round(v, 4).tolist() for k, v in gp.hyperparameters.items()} }", flush=True)
