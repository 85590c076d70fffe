"""Repeat the finite-difference BFGS fit of two concurrent children (what B200MultiOutputGPSurrogate.train does) and
compare the evaluation traces bit for bit; on the first divergence re-evaluate that theta alone."""
import os
import sys
import threading
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from profit_b200 import GPBackend, RBF, gp_functions as gpf

warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
X = rng.random((150, 2))
y = (np.sin(3 * X[:, 0]) + np.cos(2 * X[:, 1]) + 0.05 * rng.standard_normal(150)).reshape(-1, 1)
Xm = np.column_stack([2 * X[:, 0], 4 * X[:, 1] - 1])
Ym = np.column_stack([30 + 5 * np.sin(3 * X[:, 0]) + 2 * np.cos(2 * X[:, 1]), -200 + 40 * X[:, 0] * X[:, 1]])
Ym = Ym + np.array([0.25, 1.0]) * rng.standard_normal(Ym.shape)
Xe = (Xm - Xm.mean(0)) / Xm.std(0)
Ye = (Ym - Ym.mean(0)) / Ym.std(0)
a0 = np.array([0.5 * np.linalg.norm(Xe[:, None] - Xe[None], axis=-1).mean(), np.mean(np.std(Ye, axis=0)),
               1e-2 * np.mean(np.abs(Ye.max(0) - Ye.min(0)))])
traces = []


def fit(dim, out):
    be = GPBackend(0)
    trace = []
    orig = be.nll

    def rec(kind, theta, want_grad=False):
        try:
            r = orig(kind, theta, want_grad)
        except np.linalg.LinAlgError:
            trace.append((np.array(theta), "LinAlgError"))
            raise
        trace.append((np.array(theta), r))
        return r

    be.nll = rec
    hyp = gpf.optimize(Xe, Ye[:, [dim]], a0, RBF, eval_gradient=False, backend=be)
    out[dim] = (hyp, trace)
    be.close()


for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    out = {}
    ts = [threading.Thread(target=fit, args=(d, out)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    traces.append(out)
    print(rep, [out[d][0].tolist() for d in (0, 1)], [len(out[d][1]) for d in (0, 1)], flush=True)
base = traces[0]
solo = GPBackend(0)
for rep, out in enumerate(traces[1:], 1):
    for d in (0, 1):
        for i, (a, b) in enumerate(zip(base[d][1], out[d][1])):
            same_theta = np.array_equal(a[0], b[0])
            if not same_theta or a[1] != b[1]:
                solo.set_train(Xe, Ye[:, [d]])
                try:
                    alone = solo.nll(0, a[0])
                except np.linalg.LinAlgError:
                    alone = "LinAlgError"
                print(f"rep {rep} child {d}: first divergence at evaluation {i}: same theta {same_theta}, "
                      f"theta {a[0].tolist()} -> run0 {a[1]!r} / this run {b[1]!r} / alone {alone!r}")
                break
