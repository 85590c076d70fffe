"""Golden values for the non-SPD / eig-branch contract, generated from
the UNMODIFIED reference in this container:

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nonspd.py

Case "quirk": 1-D, N=1000, l=1, sf=1.5, sn=0.02.  sn^2 = 4e-4 < clip_eig = 1e-3 * l and the spectrum of K decays inside the
first 0.05 N eigenvalues, so gp_functions.negative_log_likelihood (:254-301) leaves its loop on the truncated
eigen-decomposition WITHOUT ever trying the Cholesky factorisation and returns a different number than
negative_log_likelihood_cholesky on the same, perfectly SPD matrix.  eigsh starts from a random vector and stops at
tol = clip_eig, so the eig-branch value is only reproducible to about that tolerance (three runs are stored).
Case "singular": triplicated points without noise -- the Cholesky branch fails in the reference as well.
Both cases also store the gradient the reference returns on its eig exit (gp_functions.py:291-300).
Case "matern_quirk" has no reference kernel: the oracle's Matern-5/2 callable is plugged into
the reference's own negative_log_likelihood.
Case "invert_eig": gp_functions.invert(K, neig=1) on a matrix with a prescribed spectrum (1, 1/2, 1/4, 1e-10, ...), which
leaves the doubling loop on Q diag(1/w) Q^T with four eigenpairs (:350-369).
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

from profit.sur.gp.backend.gp_functions import invert, negative_log_likelihood, negative_log_likelihood_cholesky
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
with redirect_stdout(io.StringIO()):
    a4g = negative_log_likelihood(np.log10(hyp), X, y, RBF, eval_gradient=True, log_scale_hyp=True)
out["quirk"] = {"N": 1000, "hyp": hyp.tolist(), "noise_seed": 7, "a3_nll": float(a3), "a3_grad": g3.tolist(),
                "a4_eig_branch_runs": runs, "a4_eig_branch_grad_log10": a4g[1].tolist()}

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

with redirect_stdout(io.StringIO()):
    a4sg = negative_log_likelihood(hs, Xs, ys, RBF, eval_gradient=True, log_scale_hyp=False)
out["singular"]["a4_eig_branch_grad_linear"] = a4sg[1].tolist()

# Matern-5/2 ARD, 2-D, sigma_n^2 < clip_eig: the oracle callable inside the reference's loop
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

Xm = np.random.default_rng(3).random((400, 2))
ym = np.sin(3 * Xm[:, 0]) * np.cos(2 * Xm[:, 1]) + 0.01 * np.random.default_rng(4).standard_normal(400)
hm = np.array([6.0, 9.0, 1.2, 0.05])
runs = []
for _ in range(3):
    with redirect_stdout(io.StringIO()):
        v = negative_log_likelihood(hm, Xm, ym, oracle.matern52, eval_gradient=True, log_scale_hyp=False)
    runs.append({"nll": float(v[0]), "grad": v[1].tolist()})
out["matern_quirk"] = {"N": 400, "D": 2, "x_seed": 3, "noise_seed": 4, "hyp": hm.tolist(), "runs": runs,
                       "a3_nll": float(negative_log_likelihood_cholesky(hm, Xm, ym, oracle.matern52))}

# invert(K, neig=1): prescribed spectrum, exit on the eigen-decomposition with four pairs
n = 100
Qr, _ = np.linalg.qr(np.random.default_rng(11).standard_normal((n, n)))
lam = np.concatenate([[1.0, 0.5, 0.25, 1e-10], 1e-12 * np.linspace(1.0, 0.1, n - 4)])
Kp = (Qr * lam) @ Qr.T
Kp = 0.5 * (Kp + Kp.T)
with redirect_stdout(io.StringIO()):
    Kinv = invert(Kp, neig=1)
wk, Vk = np.linalg.eigh(Kinv)
out["invert_eig"] = {"n": n, "seed": 11, "spectrum_head": lam[:4].tolist(), "tail_scale": 1e-12, "neig": 1,
                     "top_eigenvalues_of_result": wk[-4:].tolist(), "rank_of_result": int((np.abs(wk) > 1e-3).sum()),
                     "frobenius_norm": float(np.linalg.norm(Kinv)),
                     "probe": (Kinv @ np.linspace(-1.0, 1.0, n)).tolist()}

path = os.path.join(ROOT, "tests", "golden", "reference_nonspd.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
