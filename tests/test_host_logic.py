"""Host-side logic of the drop-in (no GPU): kernel selection, the log10 / fixed-sigma_n handling around the
device objective, surrogate plumbing, sharding helpers, and the world_size-2 gloo path."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import rel_err
from profit_b200 import gp_functions as gpf
from profit_b200 import kernels
from profit_b200.distributed import gather_argmax, shard_bounds
from profit_b200.surrogate import B200GPSurrogate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackedFake:
    """Stands in for GPBackend so the host logic can be exercised on CPU: same call surface, numbers from the
    oracle.  (Test infrastructure only -- the product never does this.)"""

    def __init__(self):
        self.calls = 0
        self.resident_key = self.factor_key = self.noise_override = None
        self.non_spd_events = []

    def set_train(self, X, y):
        self.noise_override = None
        self.X, self.y = np.array(X), np.array(y)
        self.n_out = 1 if self.y.ndim == 1 else self.y.shape[1]
        self.n = len(self.X)

    def nll(self, kind, theta, want_grad=False):
        self.calls += 1
        kern = oracle.rbf if kind == 0 else oracle.matern52
        return oracle.nll_cholesky(np.asarray(theta), self.X, self.y, kern, eval_gradient=want_grad)

    # -- the truncated eigen-decomposition exit (GPBackend.eigsh / nll_eig)
    def eigsh(self, kind, theta, k, k_hint=0):
        self.eig_calls = getattr(self, "eig_calls", 0) + 1
        kern = oracle.rbf if kind == 0 else oracle.matern52
        theta = np.asarray(theta)
        self._eig_kern, self._eig_theta = kern, theta
        Ky = kern(self.X, self.X, theta[:-2], theta[-2], theta[-1])
        self._eig_w, self._eig_Q = oracle.top_eigenpairs(Ky, k)
        return self._eig_w.copy()

    def nll_eig(self, neig_term, want_grad=False, n_params=None):
        w = np.maximum(self._eig_w, 1e-10)
        y = self.y.reshape(len(self.X))
        alpha = self._eig_Q @ ((self._eig_Q.T @ y) / w)
        nll = 0.5 * (y @ alpha + np.sum(np.log(w)) + neig_term * np.log(2 * np.pi))
        if not want_grad:
            return float(nll)
        th = self._eig_theta
        Ky, dKy = self._eig_kern(self.X, self.X, th[:-2], th[-2], th[-1], eval_gradient=True)
        A = oracle.invert_full(Ky) - np.outer(alpha, alpha)
        return float(nll), 0.5 * np.einsum("ij,jip->p", A, dKy)

    def mean_pairwise_distance(self, X):
        X = np.asarray(X, dtype=float)
        return float(np.linalg.norm(X[:, None, :] - X[None, :, :], axis=-1).mean())

    # -- what the predict / add_training_data plumbing touches
    def factorize(self, kind, theta):
        self.factorizations = getattr(self, "factorizations", 0) + 1
        self.noise_override = None
        self.kind, self.theta = kind, np.array(theta)
        if getattr(self, "fail_below_sn", None) is not None and self.theta[-1] < self.fail_below_sn:
            raise np.linalg.LinAlgError("not positive definite (test double)")

    def append_train(self, X, y):
        self.appends = getattr(self, "appends", 0) + 1
        self.X, self.y = np.vstack([self.X, X]), np.vstack([self.y.reshape(len(self.X), -1), y])
        self.resident_key = self.factor_key = None        # as GPBackend.append_train does; the caller re-keys

    def predict(self, Xc, add_data_variance=True, **_):
        kern = oracle.rbf if self.kind == 0 else oracle.matern52
        m, v = oracle.predict(self.X, self.y.reshape(len(self.X), -1), np.asarray(Xc), kern, self.theta[:-2],
                              self.theta[-2], self.theta[-1], add_data_variance=bool(add_data_variance))
        return m, v.ravel()


def test_select_kernel():
    assert kernels.select_kernel("RBF") is kernels.RBF
    assert kernels.select_kernel("Matern52") is kernels.Matern52
    assert kernels.select_kernel(kernels.RBF) is kernels.RBF
    assert kernels.select_kernel(oracle.rbf) is kernels.RBF          # any callable named "RBF" (e.g. the reference's)
    assert kernels.RBF.__name__ == "RBF" and kernels.Matern52.__name__ == "Matern52"
    with pytest.raises(RuntimeError, match="not implemented"):       # custom_surrogate.py:238
        kernels.select_kernel("LinearEmbedding")


def test_reference_loop_sizes_the_krylov_space_for_the_next_two_doublings():
    """The hint handed to backend.eigsh: never more than the largest pass before the first Cholesky attempt (the largest
    power of two <= 0.05 n), and for the early passes only enough for the next two doublings -- a loop that leaves at
    neig = 1 (an optimiser start at a huge length scale: lambda_1 <= 1e-3 l) must not pay for the 256-pair pass of an
    N = 8192 problem."""
    import warnings

    class Recorder(OracleBackedFake):
        def eigsh(self, kind, theta, k, k_hint=0):
            self.hints = getattr(self, "hints", []) + [(int(k), int(k_hint))]
            return super().eigsh(kind, theta, k, k_hint)

    # (1) leaves at the first pass: length scale 5e3 -> clip_eig = 5 > lambda_1 = trace bound n sf^2 + sn^2 ~ 3
    be = Recorder()
    X, y = oracle.synthetic_problem(1400, 2)
    be.set_train(X, y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", gpf.NonSPDWarning)
        gpf.negative_log_likelihood(np.array([5e3, 5e3, 0.045, 0.3]), X, y, kernels.RBF, False, False, backend=be)
    assert be.hints == [(1, 32)]                                   # 0.05 n = 70 -> largest pass 64, but only 32 asked for
    # (2) the quirk case of the test below leaves at k = 8 of n = 300 (largest pass 8): the cap is the largest pass
    be = Recorder()
    X = np.linspace(0.0, 1.0, 300).reshape(-1, 1)
    y = np.sin(3 * X[:, 0]) + 0.02 * np.random.default_rng(7).standard_normal(300)
    be.set_train(X, y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", gpf.NonSPDWarning)
        gpf.negative_log_likelihood(np.log10([1.0, 1.5, 0.02]), X, y, kernels.RBF, False, True, backend=be)
    assert [k for k, _ in be.hints] == [1, 2, 4, 8] and all(h == 8 for _, h in be.hints)


def test_reference_loop_takes_the_same_exits_as_the_reference():
    """negative_log_likelihood(on_non_spd="reference") against oracle.nll_full (the restated loop of
    gp_functions.py:254-301, itself pinned to the unmodified reference in test_oracle.py): the quirk exit on an SPD matrix,
    the exit after failed factorisations, and the usual Cholesky exit with the eigsh passes skipped."""
    import warnings

    be = OracleBackedFake()
    # (1) quirk: sigma_n^2 < clip_eig and a fast-decaying spectrum -> eig exit at k = 8 without any factorisation
    X = np.linspace(0.0, 1.0, 300).reshape(-1, 1)
    y = np.sin(3 * X[:, 0]) + 0.02 * np.random.default_rng(7).standard_normal(300)
    hyp = np.array([1.0, 1.5, 0.02])
    be.set_train(X, y)
    tr = {}
    ref = oracle.nll_full(np.log10(hyp), X, y, oracle.rbf, True, True, trace=tr)
    assert tr["branch"] == "eig" and tr["failed_cholesky"] == 0
    with pytest.warns(gpf.NonSPDWarning, match="left the Cholesky branch"):
        got = gpf.negative_log_likelihood(np.log10(hyp), X, y, kernels.RBF, True, True, backend=be)
    assert rel_err(got[0], ref[0]) < 1e-12 and rel_err(got[1], ref[1]) < 1e-9
    assert be.non_spd_events[-1]["k"] == tr["k"] and be.non_spd_events[-1]["neig"] == tr["neig"] and be.calls == 0
    # fixed sigma_n, no gradient
    ref = oracle.nll_full(np.log10(hyp[:-1]), X, y, oracle.rbf, False, True, fixed_sigma_n_value=0.02)
    with pytest.warns(gpf.NonSPDWarning):
        got = gpf.negative_log_likelihood(np.log10(hyp[:-1]), X, y, kernels.RBF, False, True, 0.02, backend=be)
    assert rel_err(got, ref) < 1e-12
    # (2) triplicated points without noise: every factorisation fails, exit on the clipped eigenvalues
    Xs = np.repeat(np.random.default_rng(0).random((10, 2)), 3, axis=0)
    ys = np.sin(Xs[:, 0])
    hs = np.array([0.5, 1.0, 1e-12])
    be.set_train(Xs, ys)
    tr = {}
    ref = oracle.nll_full(hs, Xs, ys, oracle.rbf, True, False, trace=tr)
    assert tr["branch"] == "eig" and tr["failed_cholesky"] == 4
    with pytest.warns(gpf.NonSPDWarning):
        got = gpf.negative_log_likelihood(hs, Xs, ys, kernels.RBF, True, False, backend=be)
    assert rel_err(got[0], ref[0]) < 1e-10 and rel_err(got[1], ref[1]) < 1e-6
    assert be.non_spd_events[-1]["failed_cholesky"] == 4
    # (3) the usual case: sigma_n^2 > clip_eig -> no eigsh pass is run at all and the value is the Cholesky NLL
    X3, y3 = oracle.synthetic_problem(150, 3)
    th = np.array([0.7, 0.9, 1.1, 1.5, 0.3])
    be.set_train(X3, y3)
    be.eig_calls = 0
    with warnings.catch_warnings():
        warnings.simplefilter("error", gpf.NonSPDWarning)
        got = gpf.negative_log_likelihood(th, X3, y3, kernels.Matern52, True, False, backend=be)
    tr = {}
    ref = oracle.nll_full(th, X3, y3, oracle.matern52, True, False, trace=tr)
    assert tr["branch"] == "cholesky" and be.eig_calls == 0
    assert rel_err(got[0], ref[0]) < 1e-12 and rel_err(got[1], ref[1]) < 1e-10
    # (4) sigma_n^2 < clip_eig but a slowly decaying spectrum: the passes run, nothing drops below clip_eig, Cholesky exit
    th4 = np.array([0.05, 0.05, 0.05, 1.5, 0.005])
    be.eig_calls = 0
    tr = {}
    ref = oracle.nll_full(th4, X3, y3, oracle.matern52, False, False, trace=tr)
    got = gpf.negative_log_likelihood(th4, X3, y3, kernels.Matern52, False, False, backend=be)
    assert tr["branch"] == "cholesky" and be.eig_calls == 3 and rel_err(got, ref) < 1e-12      # neig = 1, 2, 4; 8 > 0.05 N


@pytest.mark.parametrize("idx", [0, 3])
def test_objective_wrappers_match_reference_semantics(golden, idx):
    c = golden["synthetic"][idx]
    X, y = oracle.synthetic_problem(c["N"], c["D"])
    theta = np.array(c["theta"])
    kern = kernels.select_kernel(c["kernel"])
    be = OracleBackedFake()
    be.set_train(X, y)
    nll = gpf.negative_log_likelihood_cholesky(theta, X, y, kern, backend=be)
    assert isinstance(nll, float) and abs(nll - c["nll"]) <= 1e-12 * abs(c["nll"])
    nll, g = gpf.negative_log_likelihood_cholesky(theta, X, y, kern, eval_gradient=True, backend=be)
    assert rel_err(g, c["grad"]) <= 1e-10
    nw, gw = gpf.negative_log_likelihood(np.log10(theta), X, y, kern, eval_gradient=True, log_scale_hyp=True, backend=be)
    assert abs(nw - c["wrapper_log10"]["nll"]) <= 1e-12 * abs(nw) and rel_err(gw, c["wrapper_log10"]["grad"]) <= 1e-10
    nf, gf = gpf.negative_log_likelihood(np.log10(theta[:-1]), X, y, kern, eval_gradient=True, log_scale_hyp=True,
                                         fixed_sigma_n_value=theta[-1], backend=be)
    assert len(gf) == len(theta) - 1
    assert abs(nf - c["wrapper_fixed_sn"]["nll"]) <= 1e-12 * abs(nf) and rel_err(gf, c["wrapper_fixed_sn"]["grad"]) <= 1e-10


def test_optimize_driver_matches_reference_optimum(golden):
    """gp_functions.optimize: same SciPy BFGS call as the reference -> same optimum on BASELINE configs[0]."""
    c = golden["c0_train"]
    X, y = np.array(c["X"]), np.array(c["y"])
    s = B200GPSurrogate()
    s._backend = OracleBackedFake()
    s.train(X, y)
    assert s.trained and s.kernel is kernels.RBF
    hyp = np.concatenate([s.hyperparameters[k] for k in ("length_scale", "sigma_f", "sigma_n")])
    assert all(v.ndim == 1 and v.dtype == np.float64 for v in s.hyperparameters.values())
    assert np.allclose(hyp, c["hyp_opt"], rtol=1e-4)
    assert abs(oracle.nll_cholesky(hyp, X, y, oracle.rbf) - c["nll_opt"]) <= 1e-8 * abs(c["nll_opt"])
    # fixed sigma_n: one parameter fewer is optimised, sigma_n stays put
    s2 = B200GPSurrogate()
    s2._backend = OracleBackedFake()
    s2.train(X, y, hyperparameters={"length_scale": np.array([0.3]), "sigma_f": np.array([1.0]),
                                    "sigma_n": np.array([0.05])}, fixed_sigma_n=True)
    assert s2.hyperparameters["sigma_n"][0] == 0.05 and s2.hyperparameters["length_scale"].shape == (1,)
    s3 = B200GPSurrogate()
    s3._backend = OracleBackedFake()
    s3.train(X, y, return_hess_inv=True)
    assert s3.hess_inv.shape == (3, 3)


def test_surrogate_plumbing_and_errors(tmp_path):
    rng = np.random.default_rng(0)
    X, y = rng.random((30, 2)), rng.random(30)
    s = B200GPSurrogate()
    s._backend = OracleBackedFake()
    with pytest.raises(RuntimeError, match="train"):                  # sur.py:197-198
        s.predict(X)
    with pytest.raises(ValueError):
        s.pre_train(rng.random((3, 2, 2)), y)
    with pytest.raises(ValueError, match="mismatched"):
        s.pre_train(X, y[:-1])
    s.pre_train(X, y, kernel="RBF")
    assert s.Xtrain.shape == (30, 2) and s.ytrain.shape == (30, 1) and s.ndim == 2
    # inferred defaults, gaussian_process.py:124-136: l = 0.5*mean|x-x'| (scalar), sf = std(y), sn = 0.01*range(y)
    dist = np.linalg.norm(X[:, None, :] - X[None, :, :], axis=-1)
    assert np.allclose(s.hyperparameters["length_scale"], [0.5 * dist.mean()], rtol=1e-12)
    assert np.allclose(s.hyperparameters["sigma_f"], [np.std(y)])
    assert np.allclose(s.hyperparameters["sigma_n"], [1e-2 * (y.max() - y.min())])
    # add_training_data / set_ytrain (test_units.py:28-45)
    s.add_training_data(np.array([[0.1, 0.2]]), np.array([[0.3]]))
    assert s.Xtrain.shape == (31, 2) and s.ytrain.shape == (31, 1) and s.ytrain[-1, 0] == 0.3
    s.set_ytrain(np.arange(31.0).reshape(-1, 1))
    assert s.ytrain[5, 0] == 5.0
    s.trained = True
    with pytest.raises(ValueError):
        s.pre_predict(rng.random((4, 3)))
    # save / load round trip keeps the reference's attribute set
    path = str(tmp_path / "model.npz")
    s.save_model(path)
    t = B200GPSurrogate.load_model(path)
    assert t.trained and t.ndim == 2 and t.kernel is kernels.RBF and t.kernel.__name__ == "RBF"
    assert np.array_equal(t.Xtrain, s.Xtrain) and np.array_equal(t.ytrain, s.ytrain)
    for k in ("length_scale", "sigma_f", "sigma_n"):
        assert np.array_equal(t.hyperparameters[k], s.hyperparameters[k])
    with pytest.raises(RuntimeError, match="not implemented"):
        s.select_kernel("Wendland4")


def test_multi_output_wrapper_trains_one_child_per_column(golden):
    """custom_surrogate.py:285-328: independent children, merged hyper-parameters (norm / max)."""
    from profit_b200 import B200MultiOutputGPSurrogate

    c = golden["c0_train"]
    X, y = np.array(c["X"])[:60], np.array(c["y"])[:60]
    Y = np.hstack([y, 2.0 * y + 0.1])
    s = B200MultiOutputGPSurrogate()
    s._shared_backend = OracleBackedFake()
    s.train(X, Y)
    assert s.trained and s.output_ndim == 2 and len(s.models) == 2
    assert all(m._backend is s._shared_backend for m in s.models)
    # custom_surrogate.py:317-330 appends the children's values to the PARENT's own initial (inferred) ones before the
    # norm / max reduction -- mirrored, quirk included
    dist = s._shared_backend.mean_pairwise_distance(X)
    ls = np.concatenate([[0.5 * dist]] + [m.hyperparameters["length_scale"] for m in s.models])
    assert np.allclose(s.hyperparameters["length_scale"], [np.linalg.norm(ls)])
    sf0 = np.mean(np.std(Y, axis=0))
    assert np.allclose(s.hyperparameters["sigma_f"], [max([sf0] + [m.hyperparameters["sigma_f"][0] for m in s.models])])
    # the second output is an affine image of the first: same length scale, doubled signal scale
    assert np.allclose(s.models[1].hyperparameters["length_scale"], s.models[0].hyperparameters["length_scale"], rtol=5e-2)
    s.add_training_data(np.array([[0.3, 0.3]]), np.array([[0.5, 1.1]]))
    assert s.models[1].ytrain.shape == (61, 1) and s.models[1].ytrain[-1, 0] == 1.1
    s.set_ytrain(np.zeros((61, 2)))
    assert s.models[0].ytrain.shape == (61, 1)


def test_predict_retries_a_failed_factorisation_with_jitter(golden):
    """_ensure_factor retries a non-SPD Ky once with 1e-6 on the diagonal (gp_functions.invert, :343-346) and
    predict keeps reporting the model's own sigma_n^2 as data variance."""
    from profit_b200 import B200GPSurrogate

    c = golden["c0_train"]
    X, y = np.array(c["X"])[:40], np.array(c["y"])[:40]
    s = B200GPSurrogate()
    fake = s._backend = OracleBackedFake()
    s.Xtrain, s.ytrain, s.ndim, s.trained = X, y, 2, True
    s.kernel = s.select_kernel("RBF")
    s.hyperparameters = {"length_scale": np.array([0.4]), "sigma_f": np.array([1.0]), "sigma_n": np.array([1e-5])}
    fake.fail_below_sn = 5e-4                       # Ky(sigma_n = 1e-5) "fails", Ky + 1e-6 I (sigma_n' = 1.0000e-3) works
    Xq = np.random.default_rng(3).random((7, 2))
    m, v = s.predict(Xq)
    assert fake.factorizations == 2 and abs(fake.theta[-1] - np.sqrt(1e-10 + 1e-6)) < 1e-15
    assert abs(fake.noise_override - 1e-10) < 1e-24
    m0, v0 = s.predict(Xq, add_data_variance=False)
    assert fake.factorizations == 2                  # cached
    assert np.allclose(v - v0, 1e-10, rtol=0, atol=1e-15) and np.array_equal(m, m0)
    ref_m, ref_v = oracle.predict(X, y, Xq, oracle.rbf, np.array([0.4]), 1.0, np.sqrt(1e-10 + 1e-6), add_data_variance=False)
    assert np.allclose(m, ref_m) and np.allclose(v0, ref_v)


def test_shard_bounds_cover_everything_once():
    for M in (0, 1, 7, 8, 1000, 2**22):
        for G in (1, 2, 3, 4, 8):
            spans = [shard_bounds(M, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-M // G) if M else True
    assert gather_argmax(3.5, 17) == (17, 3.5)      # single process: identity


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from profit_b200.distributed import gather_argmax, gather_argmin_rows, gather_concat, gather_minmax, shard_bounds, world
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, ws = world()
v = np.sin(np.arange(1001) * 0.37); v[[100, 900]] = 2.0           # a tie across the two shards
lo, hi = shard_bounds(len(v), ws, rank)
li = int(np.argmax(v[lo:hi]))
idx, val = gather_argmax(v[lo:hi][li], lo + li)
assert (idx, val) == (int(np.argmax(v)), 2.0) and idx == 100, (idx, val)
idx, val = gather_argmax(-np.inf, -1) if rank == 1 else gather_argmax(1.0, 5)   # an empty shard
assert (idx, val) == (5, 1.0)
rows = np.array([[r, 10.0 - (r % 3), 0.1 * r, 0.2] for r in range(rank, 5, ws)])   # restarts r -> rank r mod G
best, table = gather_argmin_rows(rows)
assert best[0] == 2 and best[1] == 8.0 and table.shape == (5, 4) and list(table[:, 0]) == [0, 1, 2, 3, 4]
# np.argmax order across shards: the first NaN wins over any value
w = v.copy(); w[[700, 650]] = np.nan
lw = w[lo:hi]; li = int(np.argmax(lw))
idx, val = gather_argmax(lw[li], lo + li)
assert idx == int(np.argmax(w)) == 650 and np.isnan(val), (idx, val)
# per-shard [min, max] -> global (the normalisations of the acquisition functions), empty shard = [+inf, -inf]
cols = np.stack([v, -2.0 * v], axis=1)
local = np.stack([cols[lo:hi].min(axis=0), cols[lo:hi].max(axis=0)], axis=1)
g = gather_minmax(local)
assert np.array_equal(g, np.stack([cols.min(axis=0), cols.max(axis=0)], axis=1))
g1 = gather_minmax([[np.inf, -np.inf]] if rank == 1 else [[0.5, 0.7]])
assert np.array_equal(g1, [[0.5, 0.7]])
assert np.array_equal(gather_concat(v[lo:hi], len(v)), v)         # ragged last shard (1001 = 501 + 500)
# the restart queue of multistart_optimize: a counter in the group's store hands every index to exactly one worker
import threading, time
from profit_b200.distributed import restart_queue
nxt, mode = restart_queue(37)
assert mode == "store", mode
got = []
def pull():
    while True:
        i = nxt()
        if i is None:
            return
        got.append(i); time.sleep(0.001 * (1 + rank))             # unequal speeds: the faster rank takes more
ts = [threading.Thread(target=pull) for _ in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]
mine = np.full(37, -1.0); mine[:len(got)] = sorted(got)
both = [torch.empty(37, dtype=torch.float64) for _ in range(2)]
dist.all_gather(both, torch.from_numpy(mine))
taken = sorted(int(i) for t in both for i in t.tolist() if i >= 0)
assert taken == list(range(37)), taken
nxt2, _ = restart_queue(3)                                         # a second queue is independent of the first
mine2 = np.full(3, -1.0); got2 = list(iter(nxt2, None)); mine2[:len(got2)] = got2
both2 = [torch.empty(3, dtype=torch.float64) for _ in range(2)]
dist.all_gather(both2, torch.from_numpy(mine2))
assert sorted(int(i) for t in both2 for i in t.tolist() if i >= 0) == [0, 1, 2]
dist.barrier(); dist.destroy_process_group()
print("worker", rank, "ok")
'''


def test_world_size_2_gloo_gathers(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"worker {r} ok" in o, o


@pytest.mark.skipif(not os.path.isdir("/root/reference/profit"), reason="reference tree not mounted")
def _profit_path():
    """The unmodified reference: the pip-installed copy that travels with the repo, else the mounted tree."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "profit")):
            return p
    return None


@pytest.mark.skipif(_profit_path() is None, reason="proFit itself is not available")
def test_registers_with_profit_through_its_include_mechanism():
    """`include: profit_b200_include.py` as proFit processes it (config.py:309-320 -> util.load_includes, which runs the
    file as a stand-alone module): the surrogates and the acquisition twins land in proFit's registries."""
    shim = os.path.join(ROOT, "profit_b200_include.py")
    code = f"""
from profit.util import load_includes
load_includes([{shim!r}])
from profit.sur import Surrogate
from profit.sur.gp import GaussianProcess
from profit.al.aquisition_functions import AcquisitionFunction
import profit_b200.surrogate as s, profit_b200.acquisition as a
assert Surrogate['B200'] is s.B200GPSurrogate and issubclass(Surrogate['B200'], GaussianProcess)
assert Surrogate['B200MultiOutputGP'] is s.B200MultiOutputGPSurrogate
assert AcquisitionFunction['b200_simple_exploration'] is a.SimpleExploration
assert issubclass(a.ExpectedImprovement, AcquisitionFunction)
assert Surrogate['Custom'] is not s.B200GPSurrogate          # the reference's own labels are untouched by default
print('registered')
"""
    env = dict(os.environ, PYTHONPATH=_profit_path(), PYTHONDONTWRITEBYTECODE="1")
    env.pop("PROFIT_B200_TAKE_OVER", None)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert out.returncode == 0 and "registered" in out.stdout, out.stderr[-3000:]
    env["PROFIT_B200_TAKE_OVER"] = "1"
    code2 = code.replace("assert Surrogate['Custom'] is not s.B200GPSurrogate", "assert Surrogate['Custom'] is s.B200GPSurrogate and "
                         "AcquisitionFunction['simple_exploration'] is a.SimpleExploration")
    out = subprocess.run([sys.executable, "-c", code2], env=env, capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert out.returncode == 0 and "registered" in out.stdout, out.stderr[-3000:]


def test_add_training_data_extends_the_factor_only_when_it_is_current():
    """custom_surrogate.py:122-134 with the in-place extension: append when the device holds the factor of exactly
    this (data, hyper-parameters, kernel); otherwise leave it to the next predict, which re-uploads and refactorises."""
    rng = np.random.default_rng(1)
    X, y = rng.random((40, 2)), rng.random((40, 1))
    s = B200GPSurrogate()
    fake = OracleBackedFake()
    s._backend = fake
    s.Xtrain, s.ytrain, s.ndim, s.trained = X[:30], y[:30], 2, True
    s.kernel = s.select_kernel("RBF")
    s.hyperparameters = {"length_scale": np.array([0.7]), "sigma_f": np.array([1.2]), "sigma_n": np.array([0.2])}
    Xc = rng.random((5, 2))
    s.add_training_data(X[30:31], y[30:31])                      # nothing resident yet: no append, no factorisation
    assert getattr(fake, "appends", 0) == 0 and getattr(fake, "factorizations", 0) == 0
    m0, v0 = s.predict(Xc)
    assert fake.factorizations == 1 and len(fake.X) == 31
    s.add_training_data(X[31:33], y[31:33])                      # factor is current: extended in place
    assert fake.appends == 1 and len(fake.X) == 33 and s.Xtrain.shape[0] == 33
    m1, v1 = s.predict(Xc)
    assert fake.factorizations == 1                              # ... and the next predict does not refactorise
    rm, rv = oracle.predict(X[:33], y[:33], Xc, oracle.rbf, np.array([0.7]), 1.2, 0.2)
    assert np.allclose(m1, rm) and np.allclose(v1.ravel(), rv.ravel())
    s.hyperparameters["sigma_f"] = np.array([1.5])               # stale factor: no append, refactorise on demand
    s.add_training_data(X[33:34], y[33:34])
    assert fake.appends == 1
    s.predict(Xc)
    assert fake.factorizations == 2 and len(fake.X) == 34
    s.set_ytrain(s.ytrain * 2.0)                                 # new y: a plain re-upload + refactorisation
    s.add_training_data(X[34:35], y[34:35])
    assert fake.appends == 1
    s.predict(Xc)
    assert fake.factorizations == 3 and len(fake.X) == 35


def test_acquisition_twins_mirror_the_reference_registry():
    from profit_b200 import acquisition as acq

    labels = ("simple_exploration", "exploration_with_distance_penalty", "weighted_exploration",
              "probability_of_improvement", "expected_improvement", "expected_improvement_2", "alternating_exploration")
    for label in labels:
        cls = acq._RefBase.labels["b200_" + label]
        assert cls is acq._TWINS[label] and issubclass(cls, acq.DeviceAcquisitionFunction)
    # constructor defaults = profit/defaults.py:93-126
    Xp = np.zeros((4, 2))
    assert acq.SimpleExploration(Xp, None, None).parameters == {"use_marginal_variance": False}
    assert acq.ExplorationWithDistancePenalty(Xp, None, None).parameters == {"use_marginal_variance": False, "weight": 10}
    assert acq.WeightedExploration(Xp, None, None).parameters == {"weight": 0.5, "use_marginal_variance": False}
    assert acq.ExpectedImprovement(Xp, None, None).parameters == {"exploration_factor": 0.01, "find_min": False}
    assert acq.ExpectedImprovement2(Xp, None, None).parameters == {"exploration_factor": 0.01, "find_min": False}
    alt = acq.AlternatingAF(Xp, None, None)
    assert alt.parameters == {"alternating_freq": 1} and alt.current_af is alt.exploration
    alt.set_al_parameters(krun=3, unknown=1)
    assert alt.al_parameters["krun"] == 3
    alt.set_al_parameters(krun=0)
    with pytest.raises(TypeError, match="B200GPSurrogate"):      # no host path: a foreign surrogate is refused
        acq.SimpleExploration(Xp, object(), None).calculate_loss()


def test_empty_kernel_matrix_needs_no_device():
    """python_kernels.RBF on an empty operand returns empty arrays (n, 0) / (n, 0, P); so does the twin, without a call
    into the library."""
    from profit_b200.backend import GPBackend

    be = GPBackend.__new__(GPBackend)              # no handle: the early return must not touch the device
    K, dK = GPBackend.kernel_matrix(be, 0, np.zeros((5, 3)), np.zeros((0, 3)), [1.0, 2.0, 3.0], 1.0, 0.1, eval_gradient=True)
    assert K.shape == (5, 0) and dK.shape == (5, 0, 5)
    assert GPBackend.kernel_matrix(be, 1, np.zeros((0, 2)), np.zeros((4, 2)), [1.0], 1.0, 0.1).shape == (0, 4)
    with pytest.raises(ValueError):
        GPBackend.kernel_matrix(be, 0, np.zeros((5, 3)), np.zeros((4, 2)), [1.0], 1.0, 0.1)


def test_reference_unit_tests_of_the_surrogate_plumbing():
    """tests/unit_tests/sur/test_units.py: default hyper-parameters (:21-26), add_training_data (:29-46),
    select_kernel (:49-53) and default_Xpred (:62-69) on the drop-in class."""
    Xtrain = np.array([1, 2, 3, 4]).reshape(-1, 1)
    ytrain = np.array([5, 6, 7, 8]).reshape(-1, 1)
    sur = B200GPSurrogate()
    sur._backend = OracleBackedFake()
    sur.pre_train(Xtrain, ytrain)
    assert np.allclose(sur.hyperparameters["length_scale"], [0.625], rtol=1e-15)
    assert np.allclose(sur.hyperparameters["sigma_n"], [0.03], rtol=1e-15)
    assert sur.hyperparameters["sigma_f"] == np.std(ytrain, axis=0)
    Xadd, yadd = np.array([11, 12]).reshape(-1, 1), np.array([9, 10]).reshape(-1, 1)
    sur = B200GPSurrogate()
    sur.Xtrain, sur.ytrain = Xtrain, ytrain
    sur.add_training_data(Xadd, yadd)
    assert np.all(sur.Xtrain == np.concatenate([Xtrain, Xadd])) and np.all(sur.ytrain == np.concatenate([ytrain, yadd]))
    assert sur.select_kernel("RBF") is kernels.RBF and sur.select_kernel("RBF").__name__ == "RBF"
    sur = B200GPSurrogate()
    sur.Xtrain, sur.ndim = Xtrain, 1
    assert np.all(sur.default_Xpred() == np.linspace(Xtrain[0], Xtrain[-1], 50).reshape(-1, 1))
