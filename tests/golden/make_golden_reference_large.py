"""Golden values at BASELINE.json shapes from the UNMODIFIED reference itself (imported from /root/reference), for the
sizes it can still hold in this container's memory (SURVEY.md section 8d parity gates):

  c1: gp_functions.negative_log_likelihood_cholesky, RBF ARD, N=4096, D=8 (NLL + gradient), called directly;
  c3: GPSurrogate.predict (RBF ARD), N=8192, D=10, 4096 candidates in 4 chunks of 1024 (every call refactorises).

Run in the build container only (about 10 minutes of host time):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_reference_large.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from profit.sur import Surrogate  # noqa: E402
from profit.sur.gp.backend import gp_functions as G  # noqa: E402
from profit.sur.gp.backend.python_kernels import RBF  # noqa: E402

import oracle  # noqa: E402


def theta_ref(D):
    return np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])


out = {"generator": "tests/golden/make_golden_reference_large.py",
       "source": "unmodified profit.sur.gp (gp_functions.negative_log_likelihood_cholesky, GPSurrogate.predict)"}
t0 = time.time()

X, y = oracle.synthetic_problem(4096, 8)
th = theta_ref(8)
nll, g = G.negative_log_likelihood_cholesky(th, X, y, RBF, eval_gradient=True)
out["c1_nll"] = {"N": 4096, "D": 8, "kernel": "RBF", "theta": th.tolist(), "nll": float(nll), "grad": np.asarray(g).tolist()}
print("c1 done", round(time.time() - t0, 1), flush=True)

X, y = oracle.synthetic_problem(8192, 10)
th = theta_ref(10)
s = Surrogate["Custom"]()
s.Xtrain, s.ytrain, s.ndim, s.trained, s.kernel = X, y, 10, True, RBF
s.hyperparameters = {"length_scale": th[:-2].copy(), "sigma_f": th[-2:-1].copy(), "sigma_n": th[-1:].copy()}
Xc = np.random.default_rng(2).random((4096, 10))
means, variances = [], []
for c0 in range(0, 4096, 1024):
    m, v = s.predict(Xc[c0:c0 + 1024])
    means.append(m.ravel())
    variances.append(v.ravel())
    print("c3 chunk", c0, round(time.time() - t0, 1), flush=True)
out["c3_predict"] = {"N": 8192, "D": 10, "kernel": "RBF", "theta": th.tolist(), "seed": 2, "M": 4096,
                     "mean": np.concatenate(means).tolist(), "var": np.concatenate(variances).tolist()}

path = os.path.join(ROOT, "tests", "golden", "reference_baseline_shapes.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, round(time.time() - t0, 1))
