"""Race check of the multi-stream schedule: NLL + gradient must be bit-identical (a) across repeated evaluations and
(b) with the overlapped stages switched off (GPB_EARLY_TRTRI=0 GPB_SPLIT_T2=0 GPB_LOOKAHEAD=0 run everything in
dependency order on one stream): every tile is computed by the same task with the same accumulation order, so any
difference is a missing dependency."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = [(2304, 3), (4100, 8), (6144, 5), (6272, 10), (8192, 16), (10000, 4), (12416, 10), (16384, 10)]
if os.environ.get("GPB_STRESS_QUICK"):
    SIZES = [(4100, 8), (6272, 10), (8192, 16)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np

    from _data import synthetic_problem
    from profit_b200 import GPBackend

    be = GPBackend(0)
    out = {}
    for N, D in SIZES:
        X, y = synthetic_problem(N, D)
        theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
        be.set_train(X, y)
        res = []
        for rep in range(int(sys.argv[2])):
            kind = rep % 2
            nll, g = be.nll(kind, theta, want_grad=True)
            Xc = X[:200] + 0.01
            m, v = be.predict(Xc)
            res.append([kind, float(nll).hex(), [float(x).hex() for x in g], float(m.sum()).hex(), float(v.sum()).hex()])
        out[f"{N}x{D}"] = res
    print("RESULT " + json.dumps(out))
    sys.exit(0)


def run(env_extra, reps):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(reps)], env=env, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        raise RuntimeError(p.stdout[-2000:] + p.stderr[-2000:])
    return json.loads(line[0][7:])


reps = 3 if os.environ.get("GPB_STRESS_QUICK") else 8
a = run({}, reps)
b = run({"GPB_EARLY_TRTRI": "0", "GPB_SPLIT_T2": "0", "GPB_LOOKAHEAD": "0"}, 2)
bad = 0
for key, res in a.items():
    by_kind = {0: set(), 1: set()}
    for r in res:
        by_kind[r[0]].add(json.dumps(r[1:]))
    for r in b[key]:
        by_kind[r[0]].add(json.dumps(r[1:]))
    ok = all(len(s) == 1 for s in by_kind.values())
    bad += not ok
    print(f"{key:10s}: {reps} overlapped evaluations + 2 serial ones -> {'bit-identical' if ok else 'MISMATCH ' + str({k: len(s) for k, s in by_kind.items()})}")
print("stress ok" if bad == 0 else f"stress FAILED for {bad} sizes")
sys.exit(1 if bad else 0)
