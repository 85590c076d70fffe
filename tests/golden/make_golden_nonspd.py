"""Golden values for the non-SPD / eig-branch contract, generated from
the UNMODIFIED reference in this container:

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nonspd.py

Case "quirk": 1-D, N=1000, l=1, sf=1.5, sn=0.02.  sn^2 = 4e-4 < clip_eig = 1e-3 * l and the spectrum of K decays inside the
first 0.05 N eigenvalues, so gp_functions.negative_log_likelihood (:254-301) leaves its loop on the truncated
eigen-decomposition WITHOUT ever trying the Cholesky factorisation and returns a different number than
negative_log_likelihood_cholesky on the same, perfectly SPD matrix.  eigsh starts from a random vector and stops at
tol = clip_eig, so the eig-branch value is only reproducible to about that tolerance (three runs are stored).
Case "singular": triplicated points without noise -- the Cholesky branch fails in the reference as well.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

from profit.sur.gp.backend.gp_functions import negative_log_likelihood, negative_log_likelihood_cholesky
from profit.sur.gp.backend.python_kernels import RBF

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {"generator": "tests/golden/make_golden_nonspd.py"}

X = np.linspace(0.0, 1.0, 1000).reshape(-1, 1)
y = np.sin(3 * X[:, 0]) + 0.02 * np.random.default_rng(7).standard_normal(1000)
hyp = np.array([1.0, 1.5, 0.02])
a3, g3 = negative_log_likelihood_cholesky(hyp, X, y, RBF, eval_gradient=True)
runs = []
for _ in range(3):
    buf = io.StringIO()
    with redirect_stdout(buf):
        a4 = negative_log_likelihood(np.log10(hyp), X, y, RBF, eval_gradient=False, log_scale_hyp=True)
    runs.append({"nll": float(a4), "fallback_warnings": buf.getvalue().count("Fallback")})
out["quirk"] = {"N": 1000, "hyp": hyp.tolist(), "noise_seed": 7, "a3_nll": float(a3), "a3_grad": g3.tolist(),
                "a4_eig_branch_runs": runs}

Xs = np.repeat(np.random.default_rng(0).random((10, 2)), 3, axis=0)
ys = np.sin(Xs[:, 0])
hs = np.array([0.5, 1.0, 1e-12])
try:
    negative_log_likelihood_cholesky(hs, Xs, ys, RBF)
    chol = "ok"
except np.linalg.LinAlgError:
    chol = "LinAlgError"
buf = io.StringIO()
with redirect_stdout(buf):
    a4s = negative_log_likelihood(np.log10(hs), Xs, ys, RBF, eval_gradient=False, log_scale_hyp=True)
out["singular"] = {"hyp": hs.tolist(), "a3": chol, "a4_eig_branch_nll": float(a4s),
                   "fallback_warnings": buf.getvalue().count("Fallback")}

path = os.path.join(ROOT, "tests", "golden", "reference_nonspd.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
