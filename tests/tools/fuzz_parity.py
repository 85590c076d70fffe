"""Randomised parity sweep against the oracle (developer tool; the fixed cases live in tests/).

For the worst variance and gradient cases the record carries cond(K), and -- for N <= 100, RBF, one output -- the error of
BOTH sides against a 50-digit mpmath evaluation, so that "conditioning-limited" is shown, not asserted: the oracle
restates the reference's explicit-inverse formulas, which are the noisier side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from profit_b200 import GPBackend

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 90.0
rng = np.random.default_rng(seed)
be = GPBackend(0)
t0 = time.time(); n_cases = 0; worst = {"nll": 0, "grad": 0, "mean": 0, "var": 0, "K": 0, "dK": 0}
worst_case = {}      # metric -> description of the case that set it
truth_rows = []      # small cases: error of the GPU and of the oracle against the 50-digit truth


def mp_truth(X, y, Xc, ell, sf, sn):
    """50-digit NLL, d NLL / d sigma_n and predictive variances (+ sn^2) of the isotropic RBF model."""
    import mpmath as mp

    mp.mp.dps = 50
    n = len(X)
    f = lambda v: mp.mpf(float(v))
    k = lambda a, b: f(sf) ** 2 * mp.exp(-sum(((f(p) - f(q)) / f(ell)) ** 2 for p, q in zip(a, b)) / 2)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(i + 1):
            K[i, j] = K[j, i] = k(X[i], X[j]) + (f(sn) ** 2 if i == j else 0)
    L = mp.cholesky(K)
    yv = mp.matrix([f(v) for v in y])
    z = mp.lu_solve(L, yv)
    nll = sum(v * v for v in z) / 2 + sum(mp.log(L[i, i]) for i in range(n)) + n * mp.log(2 * mp.pi) / 2
    Kinv = K ** -1
    alpha = Kinv * yv
    dsn = (sum(Kinv[i, i] for i in range(n)) - sum(a * a for a in alpha)) * f(sn)      # 0.5 tr((K^-1 - aa^T) 2 sn I)
    var = []
    for c in Xc:
        ks = mp.matrix([k(c, X[i]) for i in range(n)])
        v = mp.lu_solve(L, ks)
        var.append(float(f(sf) ** 2 - sum(x * x for x in v) + f(sn) ** 2))
    return float(nll), float(dsn), np.array(var)
while time.time() - t0 < budget:
    N = int(rng.choice([rng.integers(1, 140), rng.integers(120, 400), rng.integers(400, 1400)]))
    D = int(rng.integers(1, 33)); n_out = int(rng.choice([1, 1, 1, 2, 3]))
    ard = bool(rng.integers(0, 2)) and D > 1
    kind = int(rng.integers(0, 2)); kern = oracle.rbf if kind == 0 else oracle.matern52
    X = rng.random((N, D)); Y = rng.standard_normal((N, n_out)) * 0.5 + np.sin(3 * X[:, :1])
    ell = rng.uniform(0.4, 1.6, D) * np.sqrt(D) / 2 if ard else np.array([rng.uniform(0.4, 1.6) * np.sqrt(D) / 2])
    sf, sn = rng.uniform(0.5, 2.0), rng.uniform(0.1, 0.5)
    theta = np.concatenate([ell, [sf, sn]])
    y_in = Y if (n_out > 1 or rng.integers(0, 2)) else Y[:, 0]
    be.set_train(X, y_in)
    nll, g = be.nll(kind, theta, want_grad=True)
    rn, rg = oracle.nll_cholesky(theta, X, y_in, kern, eval_gradient=True)
    e_n = abs(nll - rn) / abs(rn); e_g = np.max(np.abs(g - rg) / np.maximum(np.abs(rg), 1e-10 * np.abs(rg).max()))
    M = int(rng.integers(1, 300)); Xc = rng.random((M, D))
    mean, var = be.predict(Xc, add_data_variance=int(rng.integers(0, 2)))
    # oracle predict with matching add_data_variance
    adv = None
    for a in (True, False):
        rm, rv = oracle.predict(X, Y, Xc, kern, ell if ard else ell[0], sf, sn, add_data_variance=a)
        if np.allclose(rv.ravel(), var, rtol=1e-6): adv = a; break
    assert adv is not None, "variance does not match either add_data_variance mode"
    e_m = np.max(np.abs(mean - rm)) / max(np.abs(rm).max(), 1e-300); e_v = np.max(np.abs(var - rv.ravel()) / rv.ravel())
    na, nb = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    K, dK = be.kernel_matrix(kind, X[:na] if na <= N else rng.random((na, D)), Xc[:nb] if nb <= M else rng.random((nb, D)), ell, sf, sn, eval_gradient=True)
    Xa = X[:na] if na <= N else None
    if Xa is not None and nb <= M:
        rK, rdK = kern(Xa, Xc[:nb], ell if ard else ell[0], sf, sn, eval_gradient=True)
        worst["K"] = max(worst["K"], np.abs(K - rK).max() / np.abs(rK).max()); worst["dK"] = max(worst["dK"], np.abs(dK - rdK).max() / max(np.abs(rdK).max(), 1e-300))
    cond = np.linalg.cond(kern(X, X, ell if ard else ell[0], sf, sn))
    for k, v in (("nll", e_n), ("grad", e_g), ("mean", e_m), ("var", e_v)):
        if v > worst[k]:
            worst[k] = v
            worst_case[k] = dict(N=N, D=D, n_out=n_out, ard=ard, kind=kind, M=M, sn=round(sn, 3), cond=float(f"{cond:.3g}"), err=float(f"{v:.3g}"))
    if N <= 100 and kind == 0 and not ard and n_out == 1 and len(truth_rows) < 12 and (e_v > 1e-12 or e_g > 1e-12):
        t_nll, t_dsn, t_var = mp_truth(X, Y[:, 0], Xc[:8], ell[0], sf, sn)
        base = sn**2 if adv else 0.0
        gv = var[:8] + (sn**2 - base)
        rv8 = rv.ravel()[:8] + (sn**2 - base)
        truth_rows.append(dict(N=N, D=D, cond=float(f"{cond:.3g}"),
                               var_err_gpu=float(f"{np.max(np.abs(gv - t_var) / t_var):.3g}"),
                               var_err_oracle=float(f"{np.max(np.abs(rv8 - t_var) / t_var):.3g}"),
                               dsn_err_gpu=float(f"{abs(g[-1] - t_dsn) / abs(t_dsn):.3g}"),
                               dsn_err_oracle=float(f"{abs(rg[-1] - t_dsn) / abs(t_dsn):.3g}"),
                               nll_err_gpu=float(f"{abs(nll - t_nll) / abs(t_nll):.3g}"),
                               nll_err_oracle=float(f"{abs(rn - t_nll) / abs(t_nll):.3g}")))
    # the oracle's explicit-inverse variance sf^2 - k*^T K^-1 k* carries an absolute error ~ cond * eps * sf^2
    v_tol = max(1e-8, 1e-13 * cond * sf**2 / float(np.min(rv)))
    if not (e_n < 1e-9 and e_g < max(1e-8, 1e-14 * cond) and e_m < 1e-8 and e_v < v_tol):
        print("FAIL", dict(N=N, D=D, n_out=n_out, ard=ard, kind=kind, M=M, cond=cond), e_n, e_g, e_m, e_v, flush=True)
        sys.exit(1)
    n_cases += 1
print(f"fuzz ok: {n_cases} cases in {time.time()-t0:.0f} s, worst relative errors {worst}")
for k in ("nll", "grad", "mean", "var"):
    print(f"  worst {k}: {worst_case.get(k)}")
print("  small cases against the 50-digit truth (relative errors; oracle = the reference's explicit-inverse formulas):")
for row in truth_rows:
    print("   ", row)
