"""Generate tests/golden/acquisition_kats.json from the UNMODIFIED reference acquisition classes
(profit/al/aquisition_functions.py, imported from /root/reference) driving the unmodified ``GPSurrogate``.

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_acq.py
It also asserts that oracle/acq_oracle.py reproduces every stored loss from the reference's mean / variance
(the oracle is pinned here, the GPU tests then check the device path against both).
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from profit.al.aquisition_functions import AcquisitionFunction  # noqa: E402
from profit.sur import Surrogate  # noqa: E402
from profit.sur.gp.backend.python_kernels import RBF  # noqa: E402

import oracle  # noqa: E402
from oracle import acq_oracle as A  # noqa: E402


def L(a):
    return np.asarray(a, dtype=float).tolist()


def make_surrogate(X, y, theta):
    s = Surrogate["Custom"]()
    s.Xtrain, s.ytrain, s.ndim, s.trained, s.kernel = X.copy(), y.copy(), X.shape[1], True, RBF
    s.hyperparameters = {"length_scale": theta[:-2].copy(), "sigma_f": theta[-2:-1].copy(), "sigma_n": theta[-1:].copy()}
    return s


N, D, M, NS = 60, 2, 700, 80
X, y = oracle.synthetic_problem(N, D)
theta = np.array([0.6, 0.9, 1.5, 0.3])
Xpred = np.random.default_rng(2).random((M, D))
# the AL bookkeeping arrays: NS rows, the first N filled (simple_al.py / variable.py keep NaN in the unfilled rows)
vin = np.full((NS, D), np.nan)
vout = np.full((NS, 1), np.nan)
vin[:N], vout[:N] = X, y
variables = types.SimpleNamespace(input=vin, output=vout)
variables_full = types.SimpleNamespace(input=X.copy(), output=y.copy())    # no NaN rows: plain max is finite

s = make_surrogate(X, y, theta)
mu, var = s.predict(Xpred)
out = {"generator": "tests/golden/make_golden_acq.py", "N": N, "D": D, "M": M, "NS": NS, "theta": L(theta),
       "pred_seed": 2, "mean": L(mu.ravel()), "var": L(var.ravel()), "cases": {}}
C = out["cases"]


def check(name, ref, mine):
    err = np.max(np.abs(ref - mine) / np.maximum(np.abs(ref), 1e-300))
    assert err < 1e-13, (name, err)


# simple_exploration
af = AcquisitionFunction["simple_exploration"](Xpred, s, variables)
loss = af.calculate_loss()
check("simple", loss, A.loss_variance(var))
C["simple_exploration"] = {"loss": L(loss), "argmax": int(np.argmax(loss))}

# exploration_with_distance_penalty (weight 10 default and 3.5)
for w in (10, 3.5):
    af = AcquisitionFunction["exploration_with_distance_penalty"](Xpred, s, variables, weight=w)
    loss = af.calculate_loss()
    check("edp", loss, A.loss_distance_penalty(var, Xpred, X[N - 1], w))
    C[f"exploration_with_distance_penalty_w{w}"] = {"weight": w, "last_point": L(X[N - 1]), "loss": L(loss),
                                                    "argmax": int(np.argmax(loss))}

# weighted_exploration
for w in (0.5, 0.2):
    af = AcquisitionFunction["weighted_exploration"](Xpred, s, variables, weight=w)
    mun = af.normalize(s.predict(Xpred)[0])
    loss = af.calculate_loss(mun)
    check("weighted", loss, A.loss_weighted(A.normalize(mu), var, w))
    C[f"weighted_exploration_w{w}"] = {"weight": w, "mu_normalized": L(mun.ravel()), "loss": L(loss),
                                       "argmax": int(np.argmax(loss))}

# probability_of_improvement (global numpy generator, seeded)
af = AcquisitionFunction["probability_of_improvement"](Xpred, s, variables_full)
np.random.seed(1234)
loss = af.calculate_loss(mu)
np.random.seed(1234)
noise = np.random.standard_normal(mu.shape)
check("poi", loss, A.loss_poi(mu, var, noise, y.max(axis=0)))
C["probability_of_improvement"] = {"np_random_seed": 1234, "ymax": L(y.max(axis=0)), "loss": L(loss),
                                   "argmax": int(np.argmax(loss))}

# expected_improvement
for xi, fm in ((0.01, False), (0.05, True)):
    af = AcquisitionFunction["expected_improvement"](Xpred, s, variables, exploration_factor=xi, find_min=fm)
    imp = af.mu_part()
    loss = af.calculate_loss(imp)
    imp_o = A.ei_mu_part(mu.copy(), vout, xi, fm)
    check("ei_mu", imp, imp_o)
    check("ei", loss, A.loss_ei(imp_o, var))
    C[f"expected_improvement_xi{xi}_min{int(fm)}"] = {"xi": xi, "find_min": fm, "improvement": L(imp.ravel()),
                                                       "loss": L(loss), "argmax": int(np.argmax(loss))}

# expected_improvement_2
for xi, fm in ((0.01, False), (0.05, True)):
    af = AcquisitionFunction["expected_improvement_2"](Xpred, s, variables, exploration_factor=xi, find_min=fm)
    loss = af.calculate_loss()
    check("ei2", loss, A.loss_ei2(mu.copy(), var, vout, xi, fm))
    C[f"expected_improvement_2_xi{xi}_min{int(fm)}"] = {"xi": xi, "find_min": fm, "loss": L(loss),
                                                         "argmax": int(np.argmax(loss))}

# ---- the batch loop of _find_next_candidates (:63-74) with the reference surrogate, optimise() disabled so that the
# picks depend only on predict / add_training_data (what the in-place factor extension must reproduce) -------------
s2 = make_surrogate(X, y, theta)
s2.optimize = lambda *a, **k: None
af = AcquisitionFunction["simple_exploration"](Xpred, s2, variables)
cand = af.find_next_candidates(6)
m_after, v_after = s2.predict(Xpred[:50])
out["batch_simple_exploration"] = {"batch_size": 6, "candidates": L(cand), "ytrain_tail": L(s2.ytrain[N:].ravel()),
                                   "mean_after": L(m_after.ravel()), "var_after": L(v_after.ravel())}

# ---- masked pick semantics: visited rows compete with loss*0 ------------------------------------------------
lossm = C["simple_exploration"]["loss"]
order = np.argsort(lossm)[::-1][:3].tolist()
out["masked_pick"] = {"visited": order[:2], "expect": int(A.pick(np.asarray(lossm), order[:2]))}
assert out["masked_pick"]["expect"] == order[2]

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "acquisition_kats.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, os.path.getsize(path), "bytes")
