"""Active-learning helpers on the device: the acquisition losses and the masked pick against the
reference classes' golden vectors (tests/golden/acquisition_kats.json, made by make_golden_acq.py from the unmodified
``profit/al/aquisition_functions.py``) and the numpy oracle; the in-place factor extension (``gpb_append_train``)
against a full refactorisation."""
import json
import os
import types

import numpy as np
import pytest

import oracle
from oracle import acq_oracle as A

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS_RTOL = 1e-8      # inherits the predictive-variance tolerance; absolute floor relative to the loss scale


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "acquisition_kats.json")) as f:
        return json.load(f)


def make_surrogate(X, y, theta, kernel="RBF"):
    from profit_b200 import B200GPSurrogate

    s = B200GPSurrogate()
    s.Xtrain, s.ytrain, s.ndim, s.trained = X.copy(), y.copy(), X.shape[1], True
    s.kernel = s.select_kernel(kernel)
    s.hyperparameters = {"length_scale": np.array(theta[:-2], dtype=float), "sigma_f": np.array(theta[-2:-1], dtype=float),
                         "sigma_n": np.array(theta[-1:], dtype=float)}
    return s


@pytest.fixture(scope="module")
def problem(kats):
    N, D, M, NS = kats["N"], kats["D"], kats["M"], kats["NS"]
    X, y = oracle.synthetic_problem(N, D)
    Xpred = np.random.default_rng(kats["pred_seed"]).random((M, D))
    vin, vout = np.full((NS, D), np.nan), np.full((NS, 1), np.nan)
    vin[:N], vout[:N] = X, y
    return {"X": X, "y": y, "Xpred": Xpred, "theta": np.array(kats["theta"]),
            "variables": types.SimpleNamespace(input=vin, output=vout),
            "variables_full": types.SimpleNamespace(input=X.copy(), output=y.copy())}


def assert_loss(loss, ref, argmax):
    ref = np.asarray(ref)
    scale = np.abs(ref).max()
    assert loss.shape == ref.shape
    assert np.all(np.abs(loss - ref) <= LOSS_RTOL * np.abs(ref) + 1e-10 * scale), np.abs(loss - ref).max()
    top2 = np.sort(ref)[-2:]
    if top2[1] - top2[0] > 1e-7 * scale:          # a clear winner must be the same candidate
        assert int(np.argmax(loss)) == argmax


# ------------------------------------------------------------------------------------------ reference golden vectors
def test_simple_exploration_golden(kats, problem):
    from profit_b200.acquisition import SimpleExploration

    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = SimpleExploration(problem["Xpred"], s, problem["variables"])
    c = kats["cases"]["simple_exploration"]
    loss = af.calculate_loss()
    assert_loss(loss, c["loss"], c["argmax"])
    idx, val = af._acquire((), None)
    assert idx == c["argmax"] and abs(val - c["loss"][idx]) <= LOSS_RTOL * abs(val)
    # masked pick: visited rows compete with loss * 0  (aquisition_functions.py:66-67)
    mp = kats["masked_pick"]
    idx2, _ = af._acquire(tuple(mp["visited"]), None)
    assert idx2 == mp["expect"]


@pytest.mark.parametrize("key", ["exploration_with_distance_penalty_w10", "exploration_with_distance_penalty_w3.5"])
def test_distance_penalty_golden(kats, problem, key):
    from profit_b200.acquisition import ExplorationWithDistancePenalty

    c = kats["cases"][key]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = ExplorationWithDistancePenalty(problem["Xpred"], s, problem["variables"], weight=c["weight"])
    assert_loss(af.calculate_loss(), c["loss"], c["argmax"])


@pytest.mark.parametrize("key", ["weighted_exploration_w0.5", "weighted_exploration_w0.2"])
def test_weighted_exploration_golden(kats, problem, key):
    from profit_b200.acquisition import WeightedExploration

    c = kats["cases"][key]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = WeightedExploration(problem["Xpred"], s, problem["variables"], weight=c["weight"])
    mun = af.normalized_mean()
    assert np.allclose(mun.cpu().numpy().ravel(), c["mu_normalized"], rtol=1e-8, atol=1e-10)
    assert_loss(af.calculate_loss(mun), c["loss"], c["argmax"])


def test_probability_of_improvement_golden(kats, problem):
    from profit_b200.acquisition import ProbabilityOfImprovement

    c = kats["cases"]["probability_of_improvement"]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = ProbabilityOfImprovement(problem["Xpred"], s, problem["variables_full"])
    mu = af.predicted_mean()
    np.random.seed(c["np_random_seed"])          # the reference draws from numpy's global generator (:215)
    assert_loss(af.calculate_loss(mu), c["loss"], c["argmax"])
    # NaN rows in the bookkeeping array make the plain max NaN in the reference (:216); np.argmax then returns 0
    af_nan = ProbabilityOfImprovement(problem["Xpred"], s, problem["variables"])
    loss = af_nan.calculate_loss(mu)
    assert np.all(np.isnan(loss))
    assert af_nan._acquire((), None, mu)[0] == 0


@pytest.mark.parametrize("key", ["expected_improvement_xi0.01_min0", "expected_improvement_xi0.05_min1"])
def test_expected_improvement_golden(kats, problem, key):
    from profit_b200.acquisition import ExpectedImprovement

    c = kats["cases"][key]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = ExpectedImprovement(problem["Xpred"], s, problem["variables"], exploration_factor=c["xi"], find_min=c["find_min"])
    imp = af.mu_part()
    assert np.allclose(imp.cpu().numpy().ravel(), c["improvement"], rtol=1e-8, atol=1e-10)
    assert_loss(af.calculate_loss(imp), c["loss"], c["argmax"])


@pytest.mark.parametrize("key", ["expected_improvement_2_xi0.01_min0", "expected_improvement_2_xi0.05_min1"])
def test_expected_improvement_2_golden(kats, problem, key):
    from profit_b200.acquisition import ExpectedImprovement2

    c = kats["cases"][key]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = ExpectedImprovement2(problem["Xpred"], s, problem["variables"], exploration_factor=c["xi"], find_min=c["find_min"])
    assert_loss(af.calculate_loss(), c["loss"], c["argmax"])


def test_batch_loop_extends_the_factor_in_place(kats, problem):
    """_find_next_candidates (:63-74) with optimize() disabled on both sides: identical picks, identical model
    afterwards, and exactly one factorisation (every added point is a gpb_append_train)."""
    from profit_b200.acquisition import SimpleExploration

    g = kats["batch_simple_exploration"]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    s.optimize = lambda *a, **k: None
    calls = {"n": 0}
    be = s.backend
    orig = be.factorize

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    be.factorize = counting
    af = SimpleExploration(problem["Xpred"], s, problem["variables"])
    cand = af.find_next_candidates(g["batch_size"])
    assert np.array_equal(cand, np.array(g["candidates"]))
    N = kats["N"]
    assert s.Xtrain.shape[0] == N + g["batch_size"] and be.n == N + g["batch_size"]
    assert np.allclose(s.ytrain[N:].ravel(), g["ytrain_tail"], rtol=1e-8, atol=1e-10)
    m, v = s.predict(problem["Xpred"][:50])
    assert calls["n"] == 1
    assert np.allclose(m.ravel(), g["mean_after"], rtol=1e-8, atol=1e-10)
    assert np.allclose(v.ravel(), g["var_after"], rtol=1e-8)


def test_expected_improvement_2_batch_runs_without_refactorising(problem):
    from profit_b200.acquisition import ExpectedImprovement2

    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    af = ExpectedImprovement2(problem["Xpred"], s, problem["variables"])
    cand = af.find_next_candidates(4)
    assert cand.shape == (4, 2) and len({tuple(c) for c in cand}) == 4
    # the extended device model equals a fresh one on the same data
    fresh = make_surrogate(s.Xtrain, s.ytrain, problem["theta"])
    m0, v0 = s.predict(problem["Xpred"][:64])
    m1, v1 = fresh.predict(problem["Xpred"][:64])
    assert np.allclose(m0, m1, rtol=1e-8, atol=1e-10) and np.allclose(v0, v1, rtol=1e-8)


def test_alternating_and_registry(problem):
    from profit_b200 import acquisition as acq

    for label in ("simple_exploration", "exploration_with_distance_penalty", "weighted_exploration",
                  "probability_of_improvement", "expected_improvement", "expected_improvement_2",
                  "alternating_exploration"):
        assert acq._RefBase.labels["b200_" + label] is acq._TWINS[label]
    s = make_surrogate(problem["X"], problem["y"], problem["theta"])
    s.optimize = lambda *a, **k: None
    af = acq.AlternatingAF(problem["Xpred"], s, problem["variables"], alternating_freq=1)
    first = af.find_next_candidates(1)
    assert af.current_af is af.expected_improvement
    af.set_al_parameters(krun=1)
    second = af.find_next_candidates(1)
    assert af.current_af is af.exploration and first.shape == second.shape == (1, 2)
    with pytest.raises(TypeError):
        acq.SimpleExploration(problem["Xpred"], object(), problem["variables"]).calculate_loss()


# ------------------------------------------------------------------------------------------ oracle, other shapes
@pytest.mark.parametrize("n_out", [1, 3])
def test_losses_match_oracle_multi_output_device_inputs(backend, n_out):
    """Raw gpb_acquire / gpb_acq_prepare on torch CUDA tensors, several output columns (the axis=1 sums)."""
    import torch

    from profit_b200 import backend as B

    rng = np.random.default_rng(7)
    M = 5000
    mu = rng.standard_normal((M, n_out))
    var = rng.random((M, 1)) + 0.05
    yobs = rng.standard_normal((40, n_out))
    yobs[30:] = np.nan
    dev = torch.device("cuda", 0)
    mu_d, var_d = torch.from_numpy(mu).to(dev), torch.from_numpy(var.ravel().copy()).to(dev)
    loss_d = torch.empty(M, dtype=torch.float64, device=dev)

    def run(kind, ref, rtol=1e-12, **kw):
        idx, val = backend.acquire(kind, var_d, n_out=n_out, loss_out=loss_d, **kw)
        loss = loss_d.cpu().numpy()
        assert np.allclose(loss, ref, rtol=rtol, atol=rtol * np.abs(ref).max()), np.abs(loss - ref).max()
        assert idx == int(np.argmax(ref)) and abs(val - ref[idx]) <= 10 * rtol * abs(ref[idx])

    mun = backend.acq_prepare(B.ACQ_WEIGHTED, mu_d, out=torch.empty_like(mu_d))
    assert np.allclose(mun.cpu().numpy(), A.normalize(mu), rtol=1e-13, atol=1e-15)
    run(B.ACQ_WEIGHTED, A.loss_weighted(A.normalize(mu), var, 0.3), term=mun, params=[0.3])
    for xi, fm in ((0.01, False), (0.1, True)):
        par = np.concatenate([[xi, float(fm)], np.nanmax(yobs, axis=0)])
        imp = backend.acq_prepare(B.ACQ_EI, mu_d, params=par, out=torch.empty_like(mu_d))
        imp_ref = A.ei_mu_part(mu.copy(), yobs, xi, fm)
        assert np.allclose(imp.cpu().numpy(), imp_ref, rtol=1e-13, atol=1e-15)
        run(B.ACQ_EI, A.loss_ei(imp_ref, var), rtol=1e-10, term=imp)
        run(B.ACQ_EI2, A.loss_ei2(mu.copy(), var, yobs, xi, fm), rtol=1e-10, term=mu_d, params=par)
    noise = rng.standard_normal((M, n_out))
    ymax = np.nanmax(yobs, axis=0)
    run(B.ACQ_POI, A.loss_poi(mu, var, noise, ymax), term=mu_d, extra=noise, params=ymax)
    Xc = rng.random((M, 4))
    last = rng.random(4)
    run(B.ACQ_DISTANCE_PENALTY, A.loss_distance_penalty(var, Xc, last, 7.0), extra=torch.from_numpy(Xc).to(dev),
        params=np.concatenate([[7.0, 4.0], last]))
    # host inputs take the staging path and give the same answer
    loss_h = np.empty(M)
    idx, _ = backend.acquire(B.ACQ_WEIGHTED, var.ravel().copy(), term=A.normalize(mu), params=[0.3], n_out=n_out,
                             loss_out=loss_h)
    assert np.allclose(loss_h, A.loss_weighted(A.normalize(mu), var, 0.3), rtol=1e-12)


def test_argmax_semantics_large(backend):
    """np.argmax order on 2^22 values: lowest index among ties, the first NaN wins, visited rows count as 0."""
    rng = np.random.default_rng(3)
    v = rng.random(1 << 22)
    assert backend.argmax(v) == (int(np.argmax(v)), float(v.max()))
    v[[17, 4_000_000, 123_456]] = 2.0
    assert backend.argmax(v)[0] == 17
    from profit_b200 import backend as B

    assert backend.acquire(B.ACQ_VARIANCE, v, visited=[17])[0] == 123_456
    neg = -1.0 - rng.random(1000)
    assert backend.acquire(B.ACQ_VARIANCE, neg, visited=[500])[0] == 500      # -x * 0 = -0.0 beats every negative loss
    v[3_000_000] = np.nan
    v[2_000_000] = np.nan
    assert backend.argmax(v)[0] == 2_000_000
    with pytest.raises(ValueError):
        backend.acquire(B.ACQ_VARIANCE, v, visited=[1 << 22])
    with pytest.raises(ValueError):
        backend.acquire(B.ACQ_WEIGHTED, v)                                     # term missing


# ------------------------------------------------------------------------------------------ gpb_append_train
@pytest.mark.parametrize("N0,m,D,n_out,kind", [(60, 5, 2, 1, 0), (126, 5, 3, 1, 1), (128, 3, 4, 2, 0), (250, 10, 5, 3, 1),
                                               (1000, 4, 8, 1, 0)])
def test_append_matches_refactorisation(backend, N0, m, D, n_out, kind):
    """Extending the factor point by point (also across the 128 padding boundary, where the library regrows and
    refactorises) gives the model of a fresh factorisation on the concatenated data."""
    from profit_b200 import GPBackend

    X, y = oracle.synthetic_problem(N0 + m, D)
    Y = np.hstack([y * (1 + 0.3 * o) + 0.1 * o * X[:, :1] for o in range(n_out)])
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    Xc = np.random.default_rng(2).random((300, D))
    be = GPBackend(0)
    try:
        be.set_train(X[:N0], Y[:N0])
        be.factorize(kind, theta)
        be.append_train(X[N0:N0 + 1], Y[N0:N0 + 1])            # one point
        be.append_train(X[N0 + 1:], Y[N0 + 1:])                # then the rest in one call
        assert be.n == N0 + m
        m_app, v_app = be.predict(Xc)
    finally:
        be.close()
    backend.set_train(X, Y)
    backend.factorize(kind, theta)
    m_ref, v_ref = backend.predict(Xc)
    assert np.allclose(m_app, m_ref, rtol=1e-9, atol=1e-11)
    assert np.allclose(v_app, v_ref, rtol=1e-9)
    kern = oracle.rbf if kind == 0 else oracle.matern52
    m_o, v_o = oracle.predict(X, Y, Xc, kern, theta[:D], theta[D], theta[D + 1])
    assert np.allclose(m_app, m_o, rtol=1e-8, atol=1e-10) and np.allclose(v_app, v_o.ravel(), rtol=1e-8)


def test_append_then_marginal_variance_and_nll(backend):
    """The appended model feeds gpb_marginal_variance, and a later NLL sees the extended training set."""
    from profit_b200 import GPBackend

    X, y = oracle.synthetic_problem(90, 2)
    theta = np.array([0.7, 1.5, 0.3])
    Xc = np.random.default_rng(3).random((40, 2))
    H = np.diag([0.2, 0.1, 0.05])
    be = GPBackend(0)
    try:
        be.set_train(X[:80], y[:80])
        be.factorize(0, theta)
        be.append_train(X[80:], y[80:])
        mv = be.marginal_variance(Xc, H)
        nll = be.nll(0, theta)
    finally:
        be.close()
    backend.set_train(X, y)
    backend.factorize(0, theta)
    assert np.allclose(mv, backend.marginal_variance(Xc, H), rtol=1e-8)
    assert abs(nll - oracle.nll_cholesky(theta, X, y, oracle.rbf)) <= 1e-9 * abs(nll)


def test_append_argument_errors(backend):
    from profit_b200 import GPBackend

    X, y = oracle.synthetic_problem(50, 3)
    be = GPBackend(0)
    try:
        be.set_train(X, y)
        with pytest.raises(RuntimeError):
            be.append_train(X[:1], y[:1])          # no factor yet
        be.factorize(0, np.array([0.8, 1.5, 0.3]))
        with pytest.raises(ValueError):
            be.append_train(X[:1, :2], y[:1])
        with pytest.raises(ValueError):
            be.append_train(X[:2], y[:1])
    finally:
        be.close()


def test_predict_variance_only_and_mean_only(backend):
    X, y = oracle.synthetic_problem(200, 4)
    theta = np.concatenate([np.linspace(0.6, 1.2, 4), [1.5, 0.3]])
    Xc = np.random.default_rng(2).random((333, 4))
    backend.set_train(X, y)
    backend.factorize(1, theta)
    m, v = backend.predict(Xc)
    m_only, none_v = backend.predict(Xc, want_var=False)
    none_m, v_only = backend.predict(Xc, want_mean=False)
    assert none_v is None and none_m is None
    assert np.array_equal(m, m_only) and np.array_equal(v, v_only)


# ------------------------------------------------------------------------------------------ sharded over ranks
_SHARD_WORKER = r'''
import json, os, sys, types
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
import oracle
from profit_b200 import B200GPSurrogate
from profit_b200 import acquisition as acq

k = json.load(open(os.path.join(sys.argv[1], "tests", "golden", "acquisition_kats.json")))
N, D, M, NS = k["N"], k["D"], k["M"], k["NS"]
X, y = oracle.synthetic_problem(N, D)
Xpred = np.random.default_rng(k["pred_seed"]).random((M, D))
vin, vout = np.full((NS, D), np.nan), np.full((NS, 1), np.nan)
vin[:N], vout[:N] = X, y
variables = types.SimpleNamespace(input=vin, output=vout)
full = types.SimpleNamespace(input=X.copy(), output=y.copy())
theta = np.array(k["theta"])

def surrogate():
    s = B200GPSurrogate(0)                      # every rank on cuda:0 here; one GPU per rank in production
    s.Xtrain, s.ytrain, s.ndim, s.trained = X.copy(), y.copy(), D, True
    s.kernel = s.select_kernel("RBF")
    s.hyperparameters = {"length_scale": theta[:-2].copy(), "sigma_f": theta[-2:-1].copy(), "sigma_n": theta[-1:].copy()}
    return s

def check(loss, c):
    ref = np.array(c["loss"]); scale = np.abs(ref).max()
    assert loss.shape == ref.shape, (loss.shape, ref.shape)
    assert np.all(np.abs(loss - ref) <= 1e-8 * np.abs(ref) + 1e-10 * scale), np.abs(loss - ref).max()

C = k["cases"]
s = surrogate()
af = acq.SimpleExploration(Xpred, s, variables)
assert af._buffers()["sharded"] and af._buffers()["M"] < M
check(af.calculate_loss(), C["simple_exploration"])
assert af._acquire((), None)[0] == C["simple_exploration"]["argmax"]
assert af._acquire(tuple(k["masked_pick"]["visited"]), None)[0] == k["masked_pick"]["expect"]
for key in ("exploration_with_distance_penalty_w10", "exploration_with_distance_penalty_w3.5"):
    a = acq.ExplorationWithDistancePenalty(Xpred, s, variables, weight=C[key]["weight"])
    check(a.calculate_loss(), C[key]); assert a._acquire((), None)[0] == C[key]["argmax"]
for key in ("weighted_exploration_w0.5", "weighted_exploration_w0.2"):
    a = acq.WeightedExploration(Xpred, s, variables, weight=C[key]["weight"])
    check(a.calculate_loss(a.normalized_mean()), C[key])
    check(a.calculate_loss(np.array(C[key]["mu_normalized"]).reshape(-1, 1)), C[key])    # host argument over all candidates
    assert a._acquire((), None, a.normalized_mean())[0] == C[key]["argmax"]
c = C["probability_of_improvement"]
a = acq.ProbabilityOfImprovement(Xpred, s, full)
mu = a.predicted_mean()
np.random.seed(c["np_random_seed"]); check(a.calculate_loss(mu), c)
for key in ("expected_improvement_xi0.01_min0", "expected_improvement_xi0.05_min1"):
    a = acq.ExpectedImprovement(Xpred, s, variables, exploration_factor=C[key]["xi"], find_min=C[key]["find_min"])
    imp = a.mu_part()
    check(a.calculate_loss(imp), C[key]); assert a._acquire((), None, imp)[0] == C[key]["argmax"]
for key in ("expected_improvement_2_xi0.01_min0", "expected_improvement_2_xi0.05_min1"):
    a = acq.ExpectedImprovement2(Xpred, s, variables, exploration_factor=C[key]["xi"], find_min=C[key]["find_min"])
    check(a.calculate_loss(), C[key]); assert a._acquire((), None)[0] == C[key]["argmax"]
# the batch loop: every rank picks the same points and ends with the same model as the single-process reference run
g = k["batch_simple_exploration"]
s2 = surrogate(); s2.optimize = lambda *a, **kw: None
cand = acq.SimpleExploration(Xpred, s2, variables).find_next_candidates(g["batch_size"])
assert np.array_equal(cand, np.array(g["candidates"]))
m, v = s2.predict(Xpred[:50])
assert np.allclose(m.ravel(), g["mean_after"], rtol=1e-8, atol=1e-10) and np.allclose(v.ravel(), g["var_after"], rtol=1e-8)
dist.barrier(); dist.destroy_process_group()
print("shard worker", sys.argv[3], "ok")
'''


@pytest.mark.parametrize("world_size", [2, 3])
def test_acquisition_sharded_over_ranks_matches_the_reference(built_lib, tmp_path, world_size):
    """Candidates sharded over ranks (gloo here, all ranks on cuda:0): global min / max for the normalisations and the
    (value, index) gather reproduce the reference's losses and picks of the unsharded run."""
    import socket
    import subprocess
    import sys

    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = str(sock.getsockname()[1])
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(world_size)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world_size)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"shard worker {r} ok" in o, o[-3000:]
