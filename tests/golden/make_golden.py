"""Generate tests/golden/reference_kats.json from the UNMODIFIED reference (imported from /root/reference).

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed JSON.  Every number below is produced by the
reference's own functions (python_kernels.RBF, gp_functions.*, GPSurrogate) on the fixtures of its own tests
(tests/unit_tests/sur/test_kernels.py, test_gp.py) and on seeded synthetic inputs.
Matern-5/2 has no reference kernel: its entries come from the reference's NLL / predict algebra with the
oracle's Matern callable plugged in, plus scikit-learn's Matern(nu=2.5) as an independent cross-check.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from profit.sur import Surrogate  # noqa: E402
from profit.sur.gp.backend import gp_functions as G  # noqa: E402
from profit.sur.gp.backend.python_kernels import RBF  # noqa: E402
from profit.util.halton import halton  # noqa: E402

import oracle  # noqa: E402


def L(a):
    return np.asarray(a, dtype=float).tolist()


out = {"generator": "tests/golden/make_golden.py", "numpy": np.__version__}

# ---- KAT1: test_kernels.py:86-140 (test_grad_nll fixture) ------------------------------------------------
xa = np.array([[0.0, 0.0], [0.0, 0.5], [-0.5, 0.1]])
ya = np.sin(xa[:, 0]) + np.cos(xa[:, 1])
hyp = np.hstack([[3.0, 5.0], [1.0, np.sqrt(0.1)]])
nll, g = G.negative_log_likelihood_cholesky(hyp, xa, ya, RBF, eval_gradient=True)
out["kat1"] = {"X": L(xa), "y": L(ya), "hyp": L(hyp), "nll": nll, "grad": L(g)}

# ---- KAT2: test_kernels.py:31-51 (test_build_K fixture) ---------------------------------------------------
xb = np.array([[0.0, 0.0], [1.4, 1.0]])
ell = np.array([4.0, 5.0])
K2, dK2 = RBF(xa, xb, ell, 1.0, 0.0, eval_gradient=True)
out["kat2"] = {"Xa": L(xa), "Xb": L(xb), "length_scale": L(ell), "K": L(K2), "dK": L(dK2)}

# ---- KAT3/KAT4: test_kernels.py:143-165 (128 Halton points) -----------------------------------------------
xt = halton(128, 2)
yt = np.sin(2 * xt[:, 0]) + np.cos(3 * xt[:, 1])
h3 = np.array([0.5, 0.5, 1.0, 1e-4])
n3, g3 = G.negative_log_likelihood_cholesky(h3, xt, yt, RBF, eval_gradient=True)
h4 = np.array([0.5, 1.0, 1e-2])
n4, g4 = G.negative_log_likelihood_cholesky(h4, xt, yt, RBF, eval_gradient=True)
out["halton128"] = {"X": L(xt), "y": L(yt),
                    "kat3": {"hyp": L(h3), "nll": n3, "grad": L(g3), "note": "cond(K)=7.7e9: use ~1e-6"},
                    "kat4": {"hyp": L(h4), "nll": n4, "grad": L(g4)}}

# ---- test_gp.py:21-38 (10-point sine) ---------------------------------------------------------------------
x = np.linspace(0, 5, 100).reshape(100, 1)
xtr = x[::10]
ytr = np.sin(xtr)
Ky = RBF(xtr, xtr, 1.0, 1.0)
Lc = np.linalg.cholesky(Ky)
alpha = G.solve_cholesky(Lc, ytr)
out["sine10"] = {"X": L(xtr), "y": L(ytr), "Ky": L(Ky), "L": L(Lc), "alpha": L(alpha),
                 "Kinv": L(G.invert_cholesky(Lc)), "ypred": L(RBF(x, xtr, 1.0, 1.0).dot(alpha)), "xtest": L(x)}

# ---- seeded synthetic problems (inputs regenerated from seeds by oracle.synthetic_problem) ----------------
cases = []
for (N, D, n_ell, kern_name) in ((200, 2, 2, "RBF"), (257, 3, 1, "RBF"), (512, 8, 8, "RBF"), (300, 5, 5, "Matern52"),
                                 (384, 10, 1, "Matern52")):
    X, y = oracle.synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D) if n_ell == D else [0.8], [1.5, 0.3]])
    kern = RBF if kern_name == "RBF" else oracle.matern52
    nll, g = G.negative_log_likelihood_cholesky(theta, X, y, kern, eval_gradient=True)
    nll_w, g_w = G.negative_log_likelihood(np.log10(theta), X, y, kern, eval_gradient=True, log_scale_hyp=True)
    nll_f, g_f = G.negative_log_likelihood(np.log10(theta[:-1]), X, y, kern, eval_gradient=True, log_scale_hyp=True,
                                           fixed_sigma_n_value=theta[-1])
    # predict through the unmodified GPSurrogate
    s = Surrogate["Custom"]()
    s.Xtrain, s.ytrain, s.ndim, s.trained, s.kernel = X, y, D, True, kern
    s.hyperparameters = {"length_scale": theta[:-2].copy(), "sigma_f": theta[-2:-1].copy(), "sigma_n": theta[-1:].copy()}
    Xc = np.random.default_rng(2).random((96, D))
    m1, v1 = s.predict(Xc)
    m0, v0 = s.predict(Xc, add_data_variance=False)
    case = {"N": N, "D": D, "n_ell": n_ell, "kernel": kern_name, "theta": L(theta), "nll": nll, "grad": L(g),
            "wrapper_log10": {"nll": nll_w, "grad": L(g_w)}, "wrapper_fixed_sn": {"nll": nll_f, "grad": L(g_f)},
            "predict": {"M": 96, "seed": 2, "mean": L(m1), "var": L(v1), "var_nodata": L(v0)}}
    if n_ell == 1 and kern_name == "RBF":
        f, v = G.predict_f(theta.copy(), X, y, Xc, RBF)
        case["predict_f"] = {"mean": L(f), "var": L(v)}
        fc, Vc = G.predict_f(theta.copy(), X, y, Xc[:40], RBF, return_full_cov=True)
        case["predict_f"]["full_cov_first40"] = L(Vc)
    cases.append(case)
out["synthetic"] = cases

# ---- multi-output NLL (gp_functions.py:154-164,173-178) ---------------------------------------------------
X, y = oracle.synthetic_problem(200, 3)
Y2 = np.hstack([y, np.cos(2 * X[:, :1]) + 0.1 * y])
theta = np.array([0.7, 0.9, 1.1, 1.5, 0.3])
nll, g = G.negative_log_likelihood_cholesky(theta, X, Y2, RBF, eval_gradient=True)
out["multi_output"] = {"N": 200, "D": 3, "theta": L(theta), "Y": L(Y2), "nll": nll, "grad": L(g)}

# ---- GPSurrogate.train on BASELINE configs[0] (N=200, 2-D, RBF) -------------------------------------------
X0 = halton(200, 2)
y0 = (np.sin(2 * X0[:, 0]) + np.cos(3 * X0[:, 1]) + 0.05 * np.random.default_rng(1).standard_normal(200)).reshape(-1, 1)
s = Surrogate["Custom"]()
s.train(X0, y0)
Xg = np.random.default_rng(2).random((64, 2))
mg, vg = s.predict(Xg)
hyp_opt = np.concatenate([s.hyperparameters[k] for k in ("length_scale", "sigma_f", "sigma_n")])
out["c0_train"] = {"X": L(X0), "y": L(y0), "hyp_opt": L(hyp_opt),
                   "nll_opt": G.negative_log_likelihood_cholesky(hyp_opt, X0, y0, RBF),
                   "predict": {"X": L(Xg), "mean": L(mg), "var": L(vg)}}

# ---- Matern-5/2 vs scikit-learn ---------------------------------------------------------------------------
from sklearn.gaussian_process.kernels import Matern  # noqa: E402

Xm = np.random.default_rng(5).random((40, 4))
ellm = np.array([0.5, 0.8, 1.1, 1.4])
Ks, dKs = Matern(length_scale=ellm, nu=2.5)(Xm, eval_gradient=True)   # gradient w.r.t. log(length_scale)
out["matern_sklearn"] = {"seed": 5, "n": 40, "D": 4, "length_scale": L(ellm), "K": L(Ks), "dK_dlogl": L(dKs)}

# ---- marginal_variance_BBQ (gp_functions.py:376-454), scalar length scale ----------------------------------
Xb, yb = oracle.synthetic_problem(80, 2)
Xcb = np.random.default_rng(3).random((37, 2))
hypb = {"length_scale": np.array([0.7]), "sigma_f": np.array([1.5]), "sigma_n": np.array([0.3])}
rngb = np.random.default_rng(11)
Ab = rngb.standard_normal((3, 3))
Hb = Ab @ Ab.T / 3 + 0.1 * np.eye(3)
pvb = rngb.random(37).reshape(-1, 1)
out["bbq"] = {"N": 80, "D": 2, "M": 37, "seed": 3, "hyp": [0.7, 1.5, 0.3], "hess_inv": L(Hb), "pred_var": L(pvb.ravel()),
              "mv": L(G.marginal_variance_BBQ(Xb, yb, Xcb, RBF, hypb, Hb, False, predictive_variance=pvb).ravel()),
              "mv_fixed_sn": L(G.marginal_variance_BBQ(Xb, yb, Xcb, RBF, hypb, Hb[:2, :2], True,
                                                       predictive_variance=pvb).ravel()),
              "note": "gp_functions.marginal_variance_BBQ, scalar length scale"}

# ---- LinearEmbedding kernel matrix (python_kernels.py:60-105; its gradient is unverified upstream) ---------
from profit.sur.gp.backend.python_kernels import LinearEmbedding  # noqa: E402

rngl = np.random.default_rng(4)
Xl, Yl, Rl = rngl.random((23, 5)), rngl.random((17, 5)), rngl.standard_normal((2, 5))
out["linear_embedding"] = {"seed": 4, "n": 23, "m": 17, "D": 5, "d": 2, "sigma_f": 1.3, "sigma_n": 0.2,
                           "K": L(LinearEmbedding(Xl, Yl, Rl.ravel(), 1.3, 0.2))}

path = os.path.join(ROOT, "tests", "golden", "reference_kats.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, os.path.getsize(path), "bytes")
