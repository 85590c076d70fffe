"""Throughput of independent NLL+grad evaluations with several concurrent restart streams on one GPU."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52

def run(N, D, streams, evals):
    X, y = synthetic_problem(N, D)
    th0 = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    bes = [GPBackend(0) for _ in range(streams)]
    for be in bes:
        be.set_train(X, y); be.nll(KIND_MATERN52, th0, want_grad=True)
    out = [None] * streams
    def work(i):
        for k in range(evals):
            out[i] = bes[i].nll(KIND_MATERN52, th0 * (1 + 1e-3 * (i * 17 + k)), want_grad=True)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(streams)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0
    for be in bes: be.close()
    return streams * evals / dt

for (N, D) in ((16384, 10), (8192, 16), (4096, 8)):
    for s in (1, 2, 3):
        ev = 6 if N == 16384 else 20
        print(f"N={N} D={D} streams={s}: {run(N, D, s, ev):.2f} evals/s", flush=True)
