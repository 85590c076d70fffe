"""Stand-alone active-learning loop on one B200 (no proFit needed): fit a GP, pick the next points with the device
acquisition functions, and let every accepted point extend the factor in place.

    python examples/active_learning_demo.py [n_start] [n_candidates] [n_picks]
"""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from profit_b200 import B200GPSurrogate
from profit_b200.acquisition import ExpectedImprovement2, SimpleExploration


def f(X):
    return (np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + 0.5 * X[:, 2]).reshape(-1, 1)


n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 400
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
picks = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rng = np.random.default_rng(0)
X = rng.random((n0, 3))
y = f(X) + 0.02 * rng.standard_normal((n0, 1))
Xpred = rng.random((M, 3))

gp = B200GPSurrogate()
t0 = time.perf_counter()
gp.train(X, y, kernel="Matern52")
print(f"trained on {n0} points in {time.perf_counter() - t0:.2f} s:",
      {k: np.round(v, 4).tolist() for k, v in gp.hyperparameters.items()})

# the bookkeeping arrays proFit's AL keeps (rows not yet evaluated are NaN)
vin, vout = np.full((n0 + picks, 3), np.nan), np.full((n0 + picks, 1), np.nan)
vin[:n0], vout[:n0] = X, y
variables = types.SimpleNamespace(input=vin, output=vout)

# (1) pure exploration, hyper-parameters frozen inside the batch: every pick is predict -> loss -> arg-max on the
#     device (16 bytes back) and an O(N^2) extension of the factor
gp.optimize = lambda *a, **k: None
t0 = time.perf_counter()
cand = SimpleExploration(Xpred, gp, variables).find_next_candidates(picks // 2)
dt = time.perf_counter() - t0
print(f"{picks // 2} exploration picks over {M} candidates: {1e3 * dt / (picks // 2):.1f} ms per pick; "
      f"training set is now {gp.Xtrain.shape[0]} points")

# (2) one expected-improvement point followed by variance points (ExpectedImprovement2)
t0 = time.perf_counter()
cand2 = ExpectedImprovement2(Xpred, gp, variables, exploration_factor=0.01).find_next_candidates(picks - picks // 2)
dt = time.perf_counter() - t0
print(f"{picks - picks // 2} EI-2 picks: {1e3 * dt / (picks - picks // 2):.1f} ms per pick; first pick {np.round(cand2[0], 3).tolist()}, "
      f"predicted {gp.predict(cand2[:1])[0].ravel()[0]:.3f}, true {f(cand2[:1]).ravel()[0]:.3f}")
