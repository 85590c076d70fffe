"""Randomised sweep of the truncated-eigen-decomposition exits against the oracle's restatement of the reference loop
(oracle.nll_full, LAPACK eigenpairs) and of the device eigen-solver against LAPACK (developer tool; the fixed cases live
in tests/test_gpu_parity.py).  The exit the loop takes (branch, k, number of failed factorisations) must agree exactly;
the value is ill-conditioned by construction (the kept eigenvalue next to the sigma_n^2 cluster), so the record carries
the relative gap of that eigenvalue next to the error."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from profit_b200 import GPBackend, RBF, Matern52, gp_functions as gpf

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
rng = np.random.default_rng(seed)
be = GPBackend(0)
warnings.simplefilter("ignore", gpf.NonSPDWarning)
t0 = time.time()
n_eig = n_chol = n_solver = 0
worst = {"nll": (0, None), "grad": (0, None), "eigval": (0, None), "resid": (0, None), "orth": (0, None)}
while time.time() - t0 < budget:
    # ---- the NLL loop
    N = int(rng.integers(25, 700)); D = int(rng.integers(1, 4))
    kind = int(rng.integers(0, 2)); kern, okern = ((RBF, oracle.rbf), (Matern52, oracle.matern52))[kind]
    X = rng.random((N, D)); y = np.sin(3 * X[:, 0]) + 0.05 * rng.standard_normal(N)
    ell = rng.uniform(0.3, 8.0, D) if rng.integers(0, 2) else np.array([rng.uniform(0.3, 8.0)])
    sf = rng.uniform(0.5, 2.0)
    clip = max(1e-3 * ell.min(), 1e-10)
    sn = float(np.sqrt(clip) * rng.uniform(0.05, 1.5))            # mostly below the clip: the passes run
    hyp = np.concatenate([ell, [sf, sn]])
    be.set_train(X, y); be.non_spd_events = []
    tr = {}
    ref = oracle.nll_full(hyp, X, y, okern, True, False, trace=tr)
    got = gpf.negative_log_likelihood(hyp, X, y, kern, True, False, backend=be)
    if tr["branch"] == "eig":
        ev = be.non_spd_events[-1]
        assert (ev["k"], ev["neig"], ev["failed_cholesky"]) == (tr["k"], tr["neig"], tr["failed_cholesky"]), (ev, tr)
        Ky = okern(X, X, ell if ell.size > 1 else ell[0], sf, sn)
        w = np.linalg.eigvalsh(Ky)[::-1]
        k = tr["k"]
        gap = (w[k - 1] - w[k]) / w[k - 1] if k < N else 1.0
        e_n = abs(got[0] - ref[0]) / max(abs(ref[0]), 1.0)
        e_g = float(np.max(np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1e-6 * np.abs(ref[1]).max())))
        case = dict(N=N, D=D, kind=kind, k=k, sn=float(f"{sn:.3g}"), rel_gap=float(f"{gap:.2g}"))
        if e_n > worst["nll"][0]: worst["nll"] = (e_n, case)
        if e_g > worst["grad"][0]: worst["grad"] = (e_g, case)
        # the cut eigenvalue decides how well conditioned the exit value is: a relative gap g amplifies eps |lambda_max| / (g w_k)
        tol = max(1e-7, 1e-13 * w[0] / max(gap * w[k - 1], 1e-300))
        assert e_n <= min(tol, 1.0) or gap < 1e-9, ("nll", case, e_n, tol)
        n_eig += 1
    else:
        assert not be.non_spd_events, "Cholesky exit must not record an event"
        assert abs(got[0] - ref[0]) <= 1e-9 * abs(ref[0]) and np.allclose(got[1], ref[1], rtol=1e-7, atol=1e-9 * np.abs(ref[1]).max())
        n_chol += 1
    # ---- the solver itself
    n = int(rng.integers(8, 900)); k = int(rng.integers(1, max(2, n // 3)))
    Z = rng.random((n, 2))
    A = [oracle.rbf(Z, Z, rng.uniform(0.2, 3.0), 1.0, rng.uniform(1e-6, 0.3)), oracle.matern52(Z, Z, rng.uniform(0.2, 3.0), 1.0, 0.01),
         (lambda B: 0.5 * (B + B.T))(rng.standard_normal((n, n)))][int(rng.integers(0, 3))]
    wv, Q = be.eigsh_matrix(A, k, want_vectors=True)
    refw = np.linalg.eigvalsh(A); scale = np.abs(refw).max()
    errs = {"eigval": np.max(np.abs(wv - refw[n - k:])) / scale, "resid": np.max(np.abs(A @ Q - Q * wv)) / scale,
            "orth": np.max(np.abs(Q.T @ Q - np.eye(k)))}
    for name, v in errs.items():
        if v > worst[name][0]: worst[name] = (float(v), dict(n=n, k=k, steps=be.last_eig_steps))
    assert errs["eigval"] <= 1e-11 and errs["resid"] <= 1e-10 and errs["orth"] <= 1e-11, (errs, n, k)
    n_solver += 1
print(f"fuzz_eig ok (seed {seed}): {n_eig} eig exits + {n_chol} Cholesky exits of the NLL loop with identical branch decisions, "
      f"{n_solver} eigsh_matrix cases, {time.time() - t0:.0f} s")
for name, (v, case) in worst.items():
    print(f"  worst {name}: {v:.3g}  {case}")
