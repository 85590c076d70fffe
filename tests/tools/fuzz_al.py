"""Randomised parity sweep of the active-learning helpers against the oracle (developer tool): gpb_append_train vs a
fresh factorisation, every acquisition kind vs oracle/acq_oracle.py incl. the masked pick, and the full predictive
covariance."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
from oracle import acq_oracle as A
from profit_b200 import GPBackend
from profit_b200 import backend as B

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
rng = np.random.default_rng(seed)
be, fresh = GPBackend(0), GPBackend(0)
t0 = time.time()
n_cases = 0
worst = {"append_mean": 0.0, "append_var": 0.0, "loss": 0.0, "cov": 0.0}


def rel(a, b, floor):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


while time.time() - t0 < budget:
    N0 = int(rng.choice([rng.integers(2, 140), rng.integers(120, 400), rng.integers(400, 1100)]))
    m = int(rng.integers(1, 12))
    D = int(rng.integers(1, 17))
    n_out = int(rng.choice([1, 1, 2, 3]))
    kind = int(rng.integers(0, 2))
    kern = oracle.rbf if kind == 0 else oracle.matern52
    ard = bool(rng.integers(0, 2)) and D > 1
    X = rng.random((N0 + m, D))
    Y = rng.standard_normal((N0 + m, n_out)) * 0.3 + np.sin(3 * X[:, :1])
    ell = rng.uniform(0.4, 1.6, D) * np.sqrt(D) / 2 if ard else np.array([rng.uniform(0.4, 1.6) * np.sqrt(D) / 2])
    sf, sn = rng.uniform(0.5, 2.0), rng.uniform(0.1, 0.5)
    theta = np.concatenate([ell, [sf, sn]])
    M = int(rng.integers(1, 700))
    Xc = rng.random((M, D))
    # ---- append vs refactorisation
    be.set_train(X[:N0], Y[:N0])
    be.factorize(kind, theta)
    cut = int(rng.integers(0, m + 1))
    if cut:
        be.append_train(X[N0:N0 + cut], Y[N0:N0 + cut])
    if cut < m:
        be.append_train(X[N0 + cut:], Y[N0 + cut:])
    mean, var = be.predict(Xc)
    fresh.set_train(X, Y)
    fresh.factorize(kind, theta)
    rmean, rvar = fresh.predict(Xc)
    worst["append_mean"] = max(worst["append_mean"], float(np.abs(mean - rmean).max() / max(np.abs(rmean).max(), 1e-300)))
    worst["append_var"] = max(worst["append_var"], rel(var, rvar, 1e-300))
    # ---- acquisition kinds on the device-resident style inputs (host arrays here; torch path is in the tests)
    var2 = var.reshape(-1, 1)
    yobs = Y.copy()
    ymax = np.nanmax(yobs, axis=0)
    visited = sorted(set(int(v) for v in rng.integers(0, M, size=int(rng.integers(0, 4)))))
    w = float(rng.uniform(0, 1))
    xi, fm = float(rng.uniform(0, 0.1)), bool(rng.integers(0, 2))
    par = np.concatenate([[xi, float(fm)], ymax])
    noise = rng.standard_normal((M, n_out))
    last = rng.random(D)
    c1 = float(rng.uniform(1, 12))
    mun = be.acq_prepare(B.ACQ_WEIGHTED, mean)
    imp = be.acq_prepare(B.ACQ_EI, mean, params=par)
    cases = [
        (B.ACQ_VARIANCE, A.loss_variance(var2), {}),
        (B.ACQ_DISTANCE_PENALTY, A.loss_distance_penalty(var2, Xc, last, c1), dict(extra=Xc, params=np.concatenate([[c1, D], last]))),
        (B.ACQ_WEIGHTED, A.loss_weighted(A.normalize(mean), var2, w), dict(term=mun, params=[w])),
        (B.ACQ_POI, A.loss_poi(mean, var2, noise, ymax), dict(term=mean, extra=noise, params=ymax)),
        (B.ACQ_EI, A.loss_ei(A.ei_mu_part(mean.copy(), yobs, xi, fm), var2), dict(term=imp)),
        (B.ACQ_EI2, A.loss_ei2(mean.copy(), var2, yobs, xi, fm), dict(term=mean, params=par)),
    ]
    for k, ref, kw in cases:
        loss = np.empty(M)
        idx, val = be.acquire(k, var, n_out=n_out, visited=visited, loss_out=loss, **kw)
        scale = max(np.abs(ref).max(), 1e-300)
        err = float(np.abs(loss - ref).max() / scale)
        worst["loss"] = max(worst["loss"], err)
        masked = ref.copy()
        masked[visited] *= 0.0
        ok_pick = masked[idx] >= masked.max() - 1e-9 * scale
        if not (err < 1e-9 and ok_pick):
            print("FAIL acquisition", dict(kind=k, N0=N0, m=m, D=D, n_out=n_out, M=M, visited=visited), err, idx, int(np.argmax(masked)), flush=True)
            sys.exit(1)
    # ---- full covariance (single-output callable path)
    if M <= 300:
        _, cov = be.predict_cov(Xc, noise_on_diagonal=True, want_mean=False)
        _, rcov = oracle.predict_f_full_cov(np.concatenate([ell, [sf, sn]]), X, Y[:, :1], Xc, kern) if n_out == 1 else (None, None)
        if rcov is not None:
            worst["cov"] = max(worst["cov"], float(np.abs(cov - rcov).max() / np.abs(rcov).max()))
    cond = np.linalg.cond(kern(X, X, ell if ard else ell[0], sf, sn))
    v_tol = max(1e-8, 1e-13 * cond)
    if not (worst["append_mean"] < 1e-8 and worst["append_var"] < v_tol and worst["cov"] < v_tol):
        print("FAIL", dict(N0=N0, m=m, D=D, n_out=n_out, kind=kind, M=M, cond=cond), worst, flush=True)
        sys.exit(1)
    n_cases += 1
print(f"fuzz_al ok: {n_cases} cases in {time.time() - t0:.0f} s, worst relative errors {worst}")
