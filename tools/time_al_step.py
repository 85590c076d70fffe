"""Active-learning step timings on one GPU: (1) a pick over M candidates with everything kept on the device
(gpb_predict into device buffers + gpb_acquire, 16 bytes back) against the host route the reference takes (predict
to host arrays, np.argmax there); (2) adding one observed point by gpb_append_train against a refactorisation."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52
from profit_b200 import backend as B


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    out = []
    D = 10
    for N, M in ((512, 1 << 20), (2048, 1 << 19), (8192, 1 << 17), (16384, 1 << 16)):
        X, y = synthetic_problem(N + 8, D)
        theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
        Xc = np.random.default_rng(2).random((M, D))
        be = GPBackend(0)
        be.set_train(X[:N], y[:N])
        be.factorize(KIND_MATERN52, theta)
        dev = torch.device("cuda", 0)
        Xc_d = torch.from_numpy(Xc).to(dev)
        var_d = torch.empty(M, dtype=torch.float64, device=dev)
        Xc_pin = torch.from_numpy(Xc).pin_memory().numpy()
        var_h = torch.empty(M, dtype=torch.float64).pin_memory().numpy()
        picks = {}

        def device_pick():
            be.predict(Xc_d, out_var=var_d, want_mean=False)
            picks["dev"] = be.acquire(B.ACQ_VARIANCE, var_d)[0]

        def host_pick():
            be.predict(Xc_pin, out_var=var_h, want_mean=False)
            picks["host"] = int(np.argmax(var_h))

        device_pick()
        host_pick()
        t_dev, t_host = timed(device_pick), timed(host_pick)
        assert picks["dev"] == picks["host"]
        # one more observation: in-place extension vs factorising the enlarged problem
        t_refac = timed(lambda: be.factorize(KIND_MATERN52, theta))
        k = [0]

        def append_one():
            i = N + k[0]
            k[0] += 1
            be.append_train(X[i:i + 1], y[i:i + 1])

        t_app = timed(append_one, reps=5)
        rec = {"N": N, "M": M, "pick_device_ms": round(t_dev, 3), "pick_host_ms": round(t_host, 3),
               "append_ms": round(t_app, 3), "refactorize_ms": round(t_refac, 3)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        be.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/al_step.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
