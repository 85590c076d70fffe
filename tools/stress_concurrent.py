"""Two device handles driven from two host threads at once (what B200MultiOutputGPSurrogate and multistart_optimize do):
every NLL+gradient must be bit-identical to the same evaluation done alone."""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_RBF

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
D = 2
X, y = synthetic_problem(N, D)
rng = np.random.default_rng(0)
thetas = [np.array([0.3, 1.0, 0.05]) * 10 ** (0.5 * (rng.random(3) - 0.5)) for _ in range(reps)]
solo = GPBackend(0)
solo.set_train(X, y)
ref = [solo.nll(KIND_RBF, th, want_grad=True) for th in thetas]
ref_plain = [solo.nll(KIND_RBF, th) for th in thetas]
bad = []


def worker(tag, ys):
    be = GPBackend(0)
    be.set_train(X, ys)
    for i, th in enumerate(thetas):
        nll, g = be.nll(KIND_RBF, th, want_grad=True)
        if tag == 0 and (nll != ref[i][0] or not np.array_equal(g, ref[i][1])):
            bad.append((i, nll, ref[i][0]))
        n2 = be.nll(KIND_RBF, th)
        if tag == 0 and n2 != ref_plain[i]:
            bad.append((i, "plain", n2, ref_plain[i]))
    be.close()


ts = [threading.Thread(target=worker, args=(0, y)), threading.Thread(target=worker, args=(1, 2 * y + 1))]
for t in ts:
    t.start()
for t in ts:
    t.join()
print(f"N={N}: {reps} evaluations next to a second handle, mismatches: {len(bad)}", bad[:5])
sys.exit(1 if bad else 0)
