"""A/B sweep of the Cholesky scheduling switches at the BASELINE sizes: every configuration runs
in its own process (the switches are read once in gpb_create) and prints the stage times of one NLL+gradient."""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, os, json
sys.path.insert(0, %r)
import numpy as np
sys.path.insert(0, os.path.join(%r, "tools"))
from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_RBF
N, D = int(sys.argv[1]), 8
X, y = synthetic_problem(N, D)
th = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
be = GPBackend(0); be.set_train(X, y)
best = None
for _ in range(6):
    be.nll(KIND_RBF, th, want_grad=True)
    st = be.stage_ms()
    if best is None or st["total"] < best["total"]:
        best = st
print(json.dumps({k: round(best[k], 3) for k in ("potrf", "trtri", "syrk", "total")}))
""" % (ROOT, ROOT)

sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
grid = {"GPB_OUTER_BLOCKS": ["0", "1", "2"], "GPB_TAIL_BLOCKS": ["16", "32", "64"], "GPB_PDL": ["1", "0"]}
for N in sizes:
    for combo in itertools.product(*grid.values()):
        env = dict(os.environ, **dict(zip(grid.keys(), combo)))
        out = subprocess.run([sys.executable, "-c", CHILD, str(N)], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-200:]
        print(N, dict(zip(grid.keys(), combo)), line, flush=True)
