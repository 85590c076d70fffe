"""Drop-in check with the reference's own code on the GPU box: proFit (installed, unmodified, into the git-ignored
baseline/_ref by `pip install --no-deps --target baseline/_ref <reference>`) drives `Surrogate["B200"]` through its
registry and its active-learning acquisition loop, next to its own `Surrogate["Custom"]` on the same inputs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

_SCRIPT = r'''
import numpy as np
from profit.sur import Surrogate
from profit.sur.gp import GaussianProcess
from profit.al.aquisition_functions import SimpleExploration
from profit.util import load_includes
load_includes([SHIM])                                    # what `include: profit_b200_include.py` does (config.py:309-320)
import profit_b200.surrogate as ps                      # registered as Surrogate["B200"] by the include

assert Surrogate["B200"] is ps.B200GPSurrogate and issubclass(Surrogate["B200"], GaussianProcess)
rng = np.random.default_rng(0)
X = rng.random((150, 2)); y = (np.sin(3 * X[:, 0]) + np.cos(2 * X[:, 1]) + 0.05 * rng.standard_normal(150)).reshape(-1, 1)
Xc = rng.random((400, 2))

ref = Surrogate["Custom"](); ref.train(X, y)
gpu = Surrogate["B200"]();   gpu.train(X, y)
hr = np.concatenate([ref.hyperparameters[k] for k in ("length_scale", "sigma_f", "sigma_n")])
hg = np.concatenate([gpu.hyperparameters[k] for k in ("length_scale", "sigma_f", "sigma_n")])
assert np.allclose(hr, hg, rtol=1e-4), (hr, hg)
assert gpu.kernel.__name__ == "RBF"                      # asserted by the reference's own integration tests
mr, vr = ref.predict(Xc); mg, vg = gpu.predict(Xc)
assert mg.shape == mr.shape and vg.shape == vr.shape
assert np.allclose(mr, mg, rtol=1e-4, atol=1e-6) and np.allclose(vr, vg, rtol=1e-3)

# same hyper-parameters -> the two backends must agree to the north-star tolerance
gpu.hyperparameters = {k: v.copy() for k, v in ref.hyperparameters.items()}
mg, vg = gpu.predict(Xc)
assert np.max(np.abs(mg - mr)) <= 1e-8 * np.abs(mr).max() and np.max(np.abs(vg - vr) / vr) <= 1e-8

# the reference's acquisition loop (aquisition_functions.py:63-74) drives predict / add_training_data / optimize
picks = {}
for name, sur in (("ref", ref), ("gpu", gpu)):
    af = SimpleExploration(Xc, sur, None)
    picks[name] = af.find_next_candidates(2)
    assert sur.Xtrain.shape[0] == 152
assert np.allclose(picks["ref"][0], picks["gpu"][0]), picks   # identical first pick (same model, same candidates)

# device twins of the acquisition functions: registered in the reference's registry, same losses as the reference
# classes on the reference surrogate when both models hold the same hyper-parameters
import types
from profit.al.aquisition_functions import AcquisitionFunction
import profit_b200.acquisition as pa

assert AcquisitionFunction["b200_expected_improvement"] is pa.ExpectedImprovement
assert issubclass(pa.SimpleExploration, AcquisitionFunction)
ref2 = Surrogate["Custom"](); gpu2 = Surrogate["B200"]()
for s2 in (ref2, gpu2):
    s2.Xtrain, s2.ytrain, s2.ndim, s2.trained = X.copy(), y.copy(), 2, True
    s2.kernel = s2.select_kernel("RBF")
    s2.hyperparameters = {"length_scale": np.array([0.4]), "sigma_f": np.array([1.2]), "sigma_n": np.array([0.1])}
vin = np.full((160, 2), np.nan); vout = np.full((160, 1), np.nan); vin[:150] = X; vout[:150] = y
variables = types.SimpleNamespace(input=vin, output=vout)
pairs = [("simple_exploration", {}, ()), ("exploration_with_distance_penalty", {"weight": 5}, ()),
         ("weighted_exploration", {"weight": 0.4}, "mu"), ("expected_improvement", {"exploration_factor": 0.02}, "imp"),
         ("expected_improvement_2", {"find_min": True}, ())]
for label, kw, arg in pairs:
    ra = AcquisitionFunction[label](Xc, ref2, variables, **kw)
    ga = AcquisitionFunction["b200_" + label](Xc, gpu2, variables, **kw)
    if arg == "mu":
        rl, gl = ra.calculate_loss(ra.normalize(ref2.predict(Xc)[0])), ga.calculate_loss(ga.normalized_mean())
    elif arg == "imp":
        rl, gl = ra.calculate_loss(ra.mu_part()), ga.calculate_loss(ga.mu_part())
    else:
        rl, gl = ra.calculate_loss(), ga.calculate_loss()
    assert np.all(np.abs(rl - gl) <= 1e-8 * np.abs(rl) + 1e-10 * np.abs(rl).max()), (label, np.abs(rl - gl).max())
    assert int(np.argmax(rl)) == int(np.argmax(gl)), label

# the configuration route: profit.yaml with `include:` + `fit: {surrogate: B200}` as proFit's own Config reads it
# (config.py:309-320, :397-402), then from_config -> train -> save_model -> load_model -> predict
open("profit.yaml", "w").write("""
ntrain: 10
include: %s
variables:
    u: Uniform(0, 1)
    v: Uniform(0, 1)
    f: Output
run:
    worker: {class: function, entry_point: f}
fit:
    surrogate: B200
    kernel: Matern52
    save: ./model_b200.npz
active_learning:
    algorithm:
        class: simple
        acquisition_function: {class: b200_expected_improvement, exploration_factor: 0.05}
""" % SHIM)
from profit.config import BaseConfig
cfg = BaseConfig.from_file("profit.yaml")
sur = Surrogate[cfg["fit"]["surrogate"]].from_config(cfg["fit"], cfg)
assert type(sur) is ps.B200GPSurrogate and sur.kernel == "Matern52" and sur.hyperparameters["length_scale"] is None
sur.train(X, y)
assert sur.trained and sur.kernel.__name__ == "Matern52"
sur.save_model(cfg["fit"]["save"])
again = Surrogate["B200"].load_model(cfg["fit"]["save"])
m1, v1 = sur.predict(Xc); m2, v2 = again.predict(Xc)
assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
assert cfg["active_learning"]["algorithm"]["acquisition_function"]["class"] == "b200_expected_improvement"

# proFit's own SimpleAL (al/simple_al.py: warmup -> learn -> find_next_candidates -> update_run -> set_ytrain ->
# optimize) driven end to end, once with its own surrogate + acquisition function and once with the B200 twins; the
# run system is replaced by an in-process runner that evaluates the test function
from profit.al import ActiveLearning

def run_al(surrogate_label, acq_label):
    open("profit_al.yaml", "w").write("""
ntrain: 12
include: %s
variables:
    u: ActiveLearning(0, 1)
    v: ActiveLearning(0, 1)
    f: Output
run:
    worker: {class: function, entry_point: f}
fit:
    surrogate: %s
    kernel: RBF
active_learning:
    nwarmup: 6
    batch_size: 2
    nsearch: 15
    algorithm:
        class: simple
        acquisition_function: {class: %s}
""" % (SHIM, surrogate_label, acq_label))
    c = BaseConfig.from_file("profit_al.yaml")
    n = c["ntrain"]
    ins = [v.name for v in c.variable_group.input_list]
    outs = [v.name for v in c.variable_group.output_list]

    class Interface:
        def __init__(self):
            self.input = np.zeros(n, dtype=[(k, float) for k in ins])
            self.output = np.full(n, np.nan, dtype=[(k, float) for k in outs])

        def resize(self, m):
            pass

    class Runner:
        def __init__(self):
            self.interface, self.next = Interface(), 0
        input_data = property(lambda self: self.interface.input)
        output_data = property(lambda self: self.interface.output)

        def spawn_array(self, params_array, wait=True, **kw):
            for p in params_array:
                for k, val in p.items():
                    self.interface.input[k][self.next] = val
                u, v = self.interface.input["u"][self.next], self.interface.input["v"][self.next]
                self.interface.output["f"][self.next] = np.sin(3 * u) + np.cos(2 * v)
                self.next += 1

    al = ActiveLearning.from_config(Runner(), c.variable_group, c["active_learning"], c)
    al.warmup(save_intermediate=None)
    al.learn(save_intermediate=None)
    return al

al_ref = run_al("Custom", "simple_exploration")
al_gpu = run_al("B200", "b200_simple_exploration")
assert type(al_gpu.surrogate) is ps.B200GPSurrogate and type(al_gpu.acquisition_function) is pa.SimpleExploration
assert al_gpu.surrogate.Xtrain.shape == al_ref.surrogate.Xtrain.shape == (12, 2)
Xr, Xg = al_ref.variables.input, al_gpu.variables.input
assert np.array_equal(Xr[:8], Xg[:8]), (Xr, Xg)            # warm-up points and the first learned batch are identical
same = sum(bool(np.array_equal(a, b)) for a, b in zip(Xr, Xg))
assert same >= 10, (same, Xr, Xg)                          # later picks may differ only through the 1e-4 spread of the optimiser
assert np.all(np.isfinite(al_gpu.variables.output))

# `profit fit` (main.py:172-193) step by step on the files SimpleAL's variables produce: Surrogate.from_config on the
# base class (adds the configured encoders), FileHandler.load of the text files, train, save_model, and a second
# configuration that loads the saved model instead of training (sur.py:265-266)
from profit.util.file_handler import FileHandler
FileHandler.save("input.txt", al_gpu.variables.named_input)
FileHandler.save("output.txt", al_gpu.variables.formatted_output)
c = BaseConfig.from_file("profit_al.yaml")
fit_sur = Surrogate.from_config(c["fit"], c)
assert type(fit_sur) is ps.B200GPSurrogate and not fit_sur.trained
xin, yout = FileHandler.load("input.txt"), FileHandler.load("output.txt")
xin = np.hstack([xin[k] for k in xin.dtype.names]); yout = np.hstack([yout[k] for k in yout.dtype.names])
fit_sur.train(xin, yout)
fit_sur.save_model("model_fit.npz")
c.fit.load = "model_fit.npz"                             # c["fit"] hands out copies
loaded = Surrogate.from_config(c["fit"], c)
assert type(loaded) is ps.B200GPSurrogate and loaded.trained
ma, va = fit_sur.predict(Xc[:20]); mb, vb = loaded.predict(Xc[:20])
# fit_sur carries the configured input encoders (encode in pre_predict, decode in predict: an identity up to round-off,
# custom_surrogate.py:96-99); the loaded model has none, exactly as with the reference's own save / load
assert np.allclose(ma, mb, rtol=1e-10, atol=1e-12) and np.allclose(va, vb, rtol=1e-10)

# fixed_sigma_n (custom_surrogate.py:47-56,240-282; gp_functions.py:181-183,241-243): the noise level stays put and the
# remaining hyper-parameters land where the reference's do
hyp0 = {"length_scale": np.array([0.4]), "sigma_f": np.array([1.0]), "sigma_n": np.array([0.07])}
rf, gf = Surrogate["Custom"](), Surrogate["B200"]()
rf.train(X, y, hyperparameters={k: v.copy() for k, v in hyp0.items()}, fixed_sigma_n=True)
gf.train(X, y, hyperparameters={k: v.copy() for k, v in hyp0.items()}, fixed_sigma_n=True)
assert gf.hyperparameters["sigma_n"][0] == 0.07 == rf.hyperparameters["sigma_n"][0]
for k in ("length_scale", "sigma_f"):
    assert np.allclose(gf.hyperparameters[k], rf.hyperparameters[k], rtol=1e-3), (k, gf.hyperparameters[k], rf.hyperparameters[k])

# marginal variance (custom_surrogate.py:144-163 -> gp_functions.marginal_variance_BBQ): both surrogates re-optimise
# from the same hyper-parameters and add the hyper-parameter uncertainty of the mean
for s3 in (ref, gpu):
    s3.hyperparameters = {"length_scale": np.array([0.45]), "sigma_f": np.array([1.1]), "sigma_n": np.array([0.06])}
mv_r = ref.get_marginal_variance(Xc[:64])
mv_g = gpu.get_marginal_variance(Xc[:64])
assert mv_g.shape == mv_r.shape == (64, 1)
assert np.allclose(mv_g, mv_r, rtol=2e-2), np.max(np.abs(mv_g - mv_r) / mv_r)

pa.register_as_default()                                  # last: from here on the reference's labels select the twins
assert AcquisitionFunction["simple_exploration"] is pa.SimpleExploration
print("profit integration ok", hg)
'''


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "profit")), reason="reference not installed in baseline/_ref")
def test_profit_drives_the_b200_surrogate(built_lib, tmp_path):
    script = tmp_path / "integration.py"
    script.write_text(_SCRIPT.replace("SHIM", repr(os.path.join(ROOT, "profit_b200_include.py"))))
    env = dict(os.environ, PYTHONPATH=REF, PYTHONDONTWRITEBYTECODE="1")     # the include shim adds the repository itself
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert out.returncode == 0 and "profit integration ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
