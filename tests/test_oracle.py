"""The CPU oracle against the reference's known answers (committed golden vectors) and, when the reference is
mounted, against the live reference.  No GPU."""
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import rel_err

KERN = {"RBF": oracle.rbf, "Matern52": oracle.matern52}


def test_kat1_grad_nll_fixture(golden):
    k = golden["kat1"]
    nll, g = oracle.nll_cholesky(np.array(k["hyp"]), np.array(k["X"]), np.array(k["y"]), oracle.rbf, eval_gradient=True)
    assert abs(nll - k["nll"]) <= 1e-12 * abs(k["nll"])
    assert rel_err(g, k["grad"]) <= 1e-12
    # the survey's frozen values
    assert abs(k["nll"] - 1.9722432670948513) < 1e-14
    assert np.allclose(k["grad"], [0.00636503697630928, -0.01083208294680081, 0.4185845267735342, 2.8176562446606948],
                       rtol=1e-13)


def test_kat2_build_K_fixture(golden):
    k = golden["kat2"]
    K, dK = oracle.rbf(np.array(k["Xa"]), np.array(k["Xb"]), np.array(k["length_scale"]), 1.0, 0.0, eval_gradient=True)
    assert np.allclose(K, k["K"], rtol=1e-15, atol=0)
    assert np.allclose(dK, k["dK"], rtol=1e-14, atol=1e-300)
    assert np.allclose(k["K"], [[1, 0.9219631718378984], [0.9950124791926823, 0.9358968608271978],
                                [0.9920195145156273, 0.8789629784396834]], rtol=1e-15)


def test_kat3_kat4_halton(golden):
    h = golden["halton128"]
    X, y = np.array(h["X"]), np.array(h["y"])
    n4, g4 = oracle.nll_cholesky(np.array(h["kat4"]["hyp"]), X, y, oracle.rbf, eval_gradient=True)
    assert abs(n4 - h["kat4"]["nll"]) <= 1e-11 * abs(h["kat4"]["nll"])
    assert rel_err(g4, h["kat4"]["grad"]) <= 1e-9
    n3, g3 = oracle.nll_cholesky(np.array(h["kat3"]["hyp"]), X, y, oracle.rbf, eval_gradient=True)
    assert abs(n3 - h["kat3"]["nll"]) <= 1e-6 * abs(h["kat3"]["nll"])   # cond(K) = 7.7e9
    assert rel_err(g3, h["kat3"]["grad"]) <= 1e-4


def test_sine10_linear_algebra(golden):
    s = golden["sine10"]
    X, y = np.array(s["X"]), np.array(s["y"])
    Ky = oracle.rbf(X, X, 1.0, 1.0)
    assert np.array_equal(Ky, Ky.T)                           # test_gp.py:24
    assert np.allclose(Ky, s["Ky"], rtol=1e-15)
    L = np.linalg.cholesky(Ky)
    alpha = oracle.solve_cholesky(L, y)
    assert np.allclose(alpha, s["alpha"], rtol=1e-12)
    assert np.allclose(alpha, np.linalg.solve(Ky, y), rtol=1e-10, atol=1e-10)   # test_gp.py:31
    assert np.allclose(oracle.invert_cholesky(L), s["Kinv"], rtol=1e-10)
    ypred = oracle.rbf(np.array(s["xtest"]), X, 1.0, 1.0).dot(alpha)
    assert np.allclose(np.array(y), ypred[::10], rtol=1e-14, atol=1e-14)        # test_gp.py:37


@pytest.mark.parametrize("idx", range(5))
def test_synthetic_cases(golden, idx):
    c = golden["synthetic"][idx]
    X, y = oracle.synthetic_problem(c["N"], c["D"])
    theta = np.array(c["theta"])
    kern = KERN[c["kernel"]]
    nll, g = oracle.nll_cholesky(theta, X, y, kern, eval_gradient=True, block=100)   # ragged row blocks
    assert abs(nll - c["nll"]) <= 1e-13 * abs(c["nll"])
    assert rel_err(g, c["grad"]) <= 1e-11
    nw, gw = oracle.nll(np.log10(theta), X, y, kern, eval_gradient=True, log_scale_hyp=True)
    assert abs(nw - c["wrapper_log10"]["nll"]) <= 1e-13 * abs(nw)
    assert rel_err(gw, c["wrapper_log10"]["grad"]) <= 1e-11
    nf, gf = oracle.nll(np.log10(theta[:-1]), X, y, kern, eval_gradient=True, log_scale_hyp=True,
                        fixed_sigma_n_value=theta[-1])
    assert abs(nf - c["wrapper_fixed_sn"]["nll"]) <= 1e-13 * abs(nf)
    assert len(gf) == len(theta) - 1 and rel_err(gf, c["wrapper_fixed_sn"]["grad"]) <= 1e-11
    Xc = np.random.default_rng(c["predict"]["seed"]).random((c["predict"]["M"], c["D"]))
    ell = theta[:-2] if c["n_ell"] > 1 else theta[0]
    m, v = oracle.predict(X, y, Xc, kern, ell, theta[-2], theta[-1], chunk=40)      # ragged chunks
    assert rel_err(m, c["predict"]["mean"]) <= 1e-10
    assert rel_err(v, c["predict"]["var"]) <= 1e-9
    _, v0 = oracle.predict(X, y, Xc, kern, ell, theta[-2], theta[-1], add_data_variance=False)
    assert rel_err(v0, c["predict"]["var_nodata"]) <= 1e-8
    if "predict_f" in c:
        f, vf = oracle.predict_f(theta, X, y, Xc, kern)
        assert rel_err(f, c["predict_f"]["mean"]) <= 1e-10
        assert rel_err(vf, c["predict_f"]["var"]) <= 1e-9


def test_multi_output(golden):
    c = golden["multi_output"]
    X, _ = oracle.synthetic_problem(c["N"], c["D"])
    nll, g = oracle.nll_cholesky(np.array(c["theta"]), X, np.array(c["Y"]), oracle.rbf, eval_gradient=True)
    assert abs(nll - c["nll"]) <= 1e-13 * abs(c["nll"])
    assert rel_err(g, c["grad"]) <= 1e-11


def test_marginal_variance_bbq(golden):
    c = golden["bbq"]
    X, y = oracle.synthetic_problem(c["N"], c["D"])
    Xc = np.random.default_rng(c["seed"]).random((c["M"], c["D"]))
    H = np.array(c["hess_inv"])
    mv = oracle.marginal_variance_bbq(X, y, Xc, oracle.rbf, c["hyp"][0], c["hyp"][1], c["hyp"][2], H, np.array(c["pred_var"]))
    assert rel_err(mv.ravel(), c["mv"]) <= 1e-12
    Hp = np.zeros((3, 3))
    Hp[:2, :2] = H[:2, :2]
    mv2 = oracle.marginal_variance_bbq(X, y, Xc, oracle.rbf, *c["hyp"], Hp, np.array(c["pred_var"]))
    assert rel_err(mv2.ravel(), c["mv_fixed_sn"]) <= 1e-12


def test_matern_against_sklearn_vectors(golden):
    c = golden["matern_sklearn"]
    X = np.random.default_rng(c["seed"]).random((c["n"], c["D"]))
    ell = np.array(c["length_scale"])
    K, dK = oracle.matern52(X, X, ell, 1.0, 0.0, eval_gradient=True)
    assert np.abs(K - np.array(c["K"])).max() <= 1e-15
    assert np.abs(dK[..., :-2] * ell - np.array(c["dK_dlogl"])).max() <= 2e-15
    assert np.all(dK[..., -1] == 0.0)


def test_matern_gradient_finite_difference():
    X, y = oracle.synthetic_problem(60, 3)
    theta = np.array([0.7, 0.9, 1.1, 1.5, 0.3])
    _, g = oracle.nll_cholesky(theta, X, y, oracle.matern52, eval_gradient=True)
    for p in range(len(theta)):
        e = np.zeros_like(theta)
        e[p] = 1e-6
        fd = (oracle.nll_cholesky(theta + e, X, y, oracle.matern52) - oracle.nll_cholesky(theta - e, X, y, oracle.matern52)) / 2e-6
        assert abs(fd - g[p]) <= 1e-6 * max(1.0, abs(g[p]))


def test_not_spd_raises():
    X = np.zeros((4, 2))   # identical points, no noise -> singular
    with pytest.raises(np.linalg.LinAlgError):
        oracle.nll_cholesky(np.array([1.0, 1.0, 1.0, 0.0]), X, np.ones(4), oracle.rbf)


@pytest.mark.skipif(not os.path.isdir("/root/reference/profit"), reason="reference tree not mounted")
def test_oracle_against_live_reference():
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from profit.sur.gp.backend import gp_functions as G
    from profit.sur.gp.backend.python_kernels import RBF

    X, y = oracle.synthetic_problem(333, 6)
    theta = np.concatenate([np.linspace(0.6, 1.2, 6), [1.5, 0.3]])
    for yy in (y, y.ravel()):
        a = G.negative_log_likelihood_cholesky(theta, X, yy, RBF, eval_gradient=True)
        b = oracle.nll_cholesky(theta, X, yy, oracle.rbf, eval_gradient=True, block=128)
        assert abs(a[0] - b[0]) <= 1e-13 * abs(a[0]) and rel_err(b[1], a[1]) <= 1e-12
    K1, dK1 = RBF(X[:50], X[:70], theta[:6], 1.5, 0.3, eval_gradient=True)
    K2, dK2 = oracle.rbf(X[:50], X[:70], theta[:6], 1.5, 0.3, eval_gradient=True)
    assert np.array_equal(K1, K2) and np.array_equal(dK1, dK2)
    # the reference's own NLL algebra with the oracle's Matern callable plugged in
    a = G.negative_log_likelihood_cholesky(theta, X, y, oracle.matern52, eval_gradient=True)
    b = oracle.nll_cholesky(theta, X, y, oracle.matern52, eval_gradient=True)
    assert abs(a[0] - b[0]) <= 1e-13 * abs(a[0]) and rel_err(b[1], a[1]) <= 1e-12


# ------------------------------------------------------------------------------------------ acquisition oracle
def test_acquisition_oracle_reproduces_reference_losses():
    """oracle/acq_oracle.py against the losses the unmodified reference classes produced (make_golden_acq.py)."""
    import json
    import os

    from oracle import acq_oracle as A

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "acquisition_kats.json")) as f:
        k = json.load(f)
    N, D, M, NS = k["N"], k["D"], k["M"], k["NS"]
    X, y = oracle.synthetic_problem(N, D)
    Xpred = np.random.default_rng(k["pred_seed"]).random((M, D))
    mu, var = np.array(k["mean"]).reshape(-1, 1), np.array(k["var"]).reshape(-1, 1)
    # the stored mean / variance are the reference surrogate's; the GP oracle agrees with them
    theta = np.array(k["theta"])
    m_o, v_o = oracle.predict(X, y, Xpred, oracle.rbf, theta[:D], theta[D], theta[D + 1])
    assert np.allclose(m_o, mu, rtol=1e-9, atol=1e-12) and np.allclose(v_o, var, rtol=1e-9)
    vout = np.full((NS, 1), np.nan)
    vout[:N] = y
    C = k["cases"]

    def same(ref, mine):
        assert np.allclose(mine, ref, rtol=1e-13, atol=0)

    same(C["simple_exploration"]["loss"], A.loss_variance(var))
    for key in ("exploration_with_distance_penalty_w10", "exploration_with_distance_penalty_w3.5"):
        same(C[key]["loss"], A.loss_distance_penalty(var, Xpred, np.array(C[key]["last_point"]), C[key]["weight"]))
        assert np.array_equal(C[key]["last_point"], X[N - 1])
    for key in ("weighted_exploration_w0.5", "weighted_exploration_w0.2"):
        same(C[key]["mu_normalized"], A.normalize(mu).ravel())
        same(C[key]["loss"], A.loss_weighted(A.normalize(mu), var, C[key]["weight"]))
    c = C["probability_of_improvement"]
    np.random.seed(c["np_random_seed"])
    same(c["loss"], A.loss_poi(mu, var, np.random.standard_normal(mu.shape), np.array(c["ymax"])))
    for key in ("expected_improvement_xi0.01_min0", "expected_improvement_xi0.05_min1"):
        imp = A.ei_mu_part(mu.copy(), vout, C[key]["xi"], C[key]["find_min"])
        same(C[key]["improvement"], imp.ravel())
        same(C[key]["loss"], A.loss_ei(imp, var))
    for key in ("expected_improvement_2_xi0.01_min0", "expected_improvement_2_xi0.05_min1"):
        same(C[key]["loss"], A.loss_ei2(mu.copy(), var, vout, C[key]["xi"], C[key]["find_min"]))
    for key, c in C.items():
        assert A.pick(np.array(c["loss"])) == c["argmax"]
    mp = k["masked_pick"]
    assert A.pick(np.array(C["simple_exploration"]["loss"]), mp["visited"]) == mp["expect"]


def test_oracle_full_covariance_matches_reference(golden):
    c = next(c for c in golden["synthetic"] if "predict_f" in c)
    X, y = oracle.synthetic_problem(c["N"], c["D"])
    Xc = np.random.default_rng(c["predict"]["seed"]).random((c["predict"]["M"], c["D"]))[:40]
    f, V = oracle.predict_f_full_cov(np.array(c["theta"]), X, y, Xc, oracle.rbf)
    ref = np.array(c["predict_f"]["full_cov_first40"])
    assert np.max(np.abs(V - ref)) <= 1e-10 * np.abs(ref).max()
    assert np.allclose(f.ravel(), np.array(c["predict_f"]["mean"]).ravel()[:40], rtol=1e-9, atol=1e-12)


def test_oracle_against_the_reference_at_c1_shape():
    """The blocked oracle NLL + gradient at BASELINE configs[1] (N=4096, D=8) against the unmodified reference's value
    (tests/golden/reference_baseline_shapes.json): this pins the large-N restatement the GPU tests lean on."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "reference_baseline_shapes.json")) as f:
        c = json.load(f)["c1_nll"]
    X, y = oracle.synthetic_problem(c["N"], c["D"])
    nll, g = oracle.nll_cholesky(np.array(c["theta"]), X, y, oracle.rbf, eval_gradient=True, block=512)
    assert abs(nll - c["nll"]) <= 1e-11 * abs(c["nll"])
    assert np.max(np.abs(g - np.array(c["grad"])) / np.abs(c["grad"])) <= 1e-9


def _nonspd_golden():
    import json

    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_nonspd.json")))


def nonspd_problems():
    """The inputs of tests/golden/make_golden_nonspd.py (shared with the GPU twin of this test)."""
    X = np.linspace(0.0, 1.0, 1000).reshape(-1, 1)
    y = np.sin(3 * X[:, 0]) + 0.02 * np.random.default_rng(7).standard_normal(1000)
    Xs = np.repeat(np.random.default_rng(0).random((10, 2)), 3, axis=0)
    ys = np.sin(Xs[:, 0])
    Xm = np.random.default_rng(3).random((400, 2))
    ym = np.sin(3 * Xm[:, 0]) * np.cos(2 * Xm[:, 1]) + 0.01 * np.random.default_rng(4).standard_normal(400)
    n = 100
    Qr, _ = np.linalg.qr(np.random.default_rng(11).standard_normal((n, n)))
    lam = np.concatenate([[1.0, 0.5, 0.25, 1e-10], 1e-12 * np.linspace(1.0, 0.1, n - 4)])
    Kp = (Qr * lam) @ Qr.T
    return {"quirk": (X, y), "singular": (Xs, ys), "matern_quirk": (Xm, ym), "invert_eig": 0.5 * (Kp + Kp.T)}


def test_nll_full_takes_the_reference_eig_exits():
    """oracle.nll_full / invert_full (the whole loops of gp_functions.py:254-301 and :350-373) against values of the
    unmodified reference.  Tolerances: the reference's eigsh starts from a random vector and stops at tol = 1e-3 min|l|
    (three stored runs scatter by ~5e-7 relative); the oracle takes the eigenpairs from LAPACK."""
    gold = _nonspd_golden()
    P = nonspd_problems()
    q = gold["quirk"]
    tr = {}
    nll, g = oracle.nll_full(np.log10(q["hyp"]), *P["quirk"], oracle.rbf, True, True, trace=tr)
    assert tr == {"branch": "eig", "neig": 16, "k": 8, "failed_cholesky": 0, "w_min": tr["w_min"]}
    for run in q["a4_eig_branch_runs"]:
        assert abs(nll - run["nll"]) <= 5e-6 * abs(run["nll"])
    assert rel_err(g, q["a4_eig_branch_grad_log10"]) <= 1e-6
    assert abs(nll - q["a3_nll"]) > 1000                      # and it is NOT the Cholesky NLL of the same SPD matrix
    s = gold["singular"]
    tr = {}
    nll, g = oracle.nll_full(np.array(s["hyp"]), *P["singular"], oracle.rbf, True, False, trace=tr)
    assert tr["branch"] == "eig" and tr["failed_cholesky"] == s["fallback_warnings"] == 4 and tr["k"] == 16
    assert abs(nll - s["a4_eig_branch_nll"]) <= 1e-10 * abs(nll)
    assert np.allclose(g, s["a4_eig_branch_grad_linear"], rtol=1e-6, atol=1e-10)
    m = gold["matern_quirk"]
    tr = {}
    nll, g = oracle.nll_full(np.array(m["hyp"]), *P["matern_quirk"], oracle.matern52, True, False, trace=tr)
    assert tr["branch"] == "eig" and tr["failed_cholesky"] == 0
    for run in m["runs"]:
        assert abs(nll - run["nll"]) <= 1e-8 * abs(run["nll"]) and rel_err(g, run["grad"]) <= 1e-6
    e = gold["invert_eig"]
    tr = {}
    Kinv = oracle.invert_full(P["invert_eig"], neig=e["neig"], trace=tr)
    assert tr == {"branch": "eig", "neig": 4}
    w = np.linalg.eigvalsh(Kinv)
    assert np.allclose(w[-4:], e["top_eigenvalues_of_result"], rtol=1e-6) and int((np.abs(w) > 1e-3).sum()) == 4
    assert np.allclose(Kinv @ np.linspace(-1.0, 1.0, e["n"]), e["probe"], rtol=1e-5, atol=1e-5 * e["frobenius_norm"])
    # neig = 0 (what every in-package caller passes) stays on the Cholesky branch
    tr = {}
    K = oracle.rbf(*[P["quirk"][0][:50]] * 2, 0.3, 1.0, 0.1)
    assert np.allclose(oracle.invert_full(K, trace=tr) @ K, np.eye(50), atol=1e-8) and tr["branch"] == "cholesky"
