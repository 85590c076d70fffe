"""Golden values at the BASELINE.json shapes (configs[2], [3], [4]) from the CPU oracle (oracle/gp_oracle.py, itself
pinned against the reference by tests/test_oracle.py).  The unmodified reference cannot run these sizes in 62 GB.
Takes ~10 minutes of host time:   python tests/golden/make_golden_large.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle  # noqa: E402

out = {"generator": "tests/golden/make_golden_large.py", "source": "oracle.nll_cholesky / oracle.predict"}


def theta_ref(D):
    return np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])


t0 = time.time()
# configs[3]: predict vs N=8192 training points (D=10, Matern-5/2), 4096 of the candidates
X, y = oracle.synthetic_problem(8192, 10)
th = theta_ref(10)
Xc = np.random.default_rng(2).random((4096, 10))
m, v = oracle.predict(X, y, Xc, oracle.matern52, th[:-2], th[-2], th[-1], chunk=1024)
out["c3_predict"] = {"N": 8192, "D": 10, "kernel": "Matern52", "theta": th.tolist(), "seed": 2, "M": 4096,
                     "mean": m.ravel().tolist(), "var": v.ravel().tolist()}
print("c3 done", time.time() - t0, flush=True)

# configs[4]: N=8192, D=16, RBF ARD NLL+grad
X, y = oracle.synthetic_problem(8192, 16)
th = theta_ref(16)
nll, g = oracle.nll_cholesky(th, X, y, oracle.rbf, eval_gradient=True, block=256)
out["c4_nll"] = {"N": 8192, "D": 16, "kernel": "RBF", "theta": th.tolist(), "nll": nll, "grad": g.tolist()}
print("c4 done", time.time() - t0, flush=True)

# configs[1]: N=4096, D=8, RBF ARD
X, y = oracle.synthetic_problem(4096, 8)
th = theta_ref(8)
nll, g = oracle.nll_cholesky(th, X, y, oracle.rbf, eval_gradient=True, block=256)
out["c1_nll"] = {"N": 4096, "D": 8, "kernel": "RBF", "theta": th.tolist(), "nll": nll, "grad": g.tolist()}
print("c1 done", time.time() - t0, flush=True)

# configs[2]: N=16384, D=10, Matern-5/2 ARD NLL+grad (the headline workload)
X, y = oracle.synthetic_problem(16384, 10)
th = theta_ref(10)
nll, g = oracle.nll_cholesky(th, X, y, oracle.matern52, eval_gradient=True, block=256)
out["c2_nll"] = {"N": 16384, "D": 10, "kernel": "Matern52", "theta": th.tolist(), "nll": nll, "grad": g.tolist()}
print("c2 done", time.time() - t0, flush=True)

path = os.path.join(ROOT, "tests", "golden", "oracle_baseline_shapes.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path)
