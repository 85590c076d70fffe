import os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from profit_b200 import B200GPSurrogate
X0 = bench.halton(200, 2)
y0 = (np.sin(2 * X0[:, 0]) + np.cos(3 * X0[:, 1]) + 0.05 * np.random.default_rng(1).standard_normal(200)).reshape(-1, 1)
g = np.linspace(0, 1, 50)
Xg = np.stack(np.meshgrid(g, g), axis=-1).reshape(-1, 2)
B200GPSurrogate().train(X0, y0)
for rep in range(2):
    s = B200GPSurrogate()
    with warnings.catch_warnings(record=True) as wl:
        warnings.simplefilter("always")
        t0 = time.perf_counter(); s.train(X0, y0); t1 = time.perf_counter()
        n0 = s.backend.nll_calls if hasattr(s.backend, "nll_calls") else -1
        m, v = s.predict(Xg); t2 = time.perf_counter()
        m, v = s.predict(Xg); t3 = time.perf_counter()
    print(f"train {1e3*(t1-t0):.2f} ms, predict {1e3*(t2-t1):.2f} ms, again {1e3*(t3-t2):.2f} ms, hyp {s.hyperparameters}, warnings {len(wl)}, nll_calls {n0}")
    for w in wl[:3]: print("   ", str(w.message)[:200])
