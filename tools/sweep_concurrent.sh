# A/B of the scheduling switches under two concurrent restart streams (the mode of bench.py's configs[4] leg)
cat > /tmp/cr.py <<'P'
import sys, os, time, threading
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np
from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_RBF
N, D, streams, evals = 8192, 16, int(sys.argv[1]), 30
X, y = synthetic_problem(N, D)
th0 = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
bes = [GPBackend(0) for _ in range(streams)]
for be in bes:
    be.set_train(X, y); be.nll(KIND_RBF, th0, want_grad=True)
def work(i):
    for k in range(evals):
        bes[i].nll(KIND_RBF, th0 * (1 + 1e-3 * (i * 17 + k)), want_grad=True)
ts = [threading.Thread(target=work, args=(i,)) for i in range(streams)]
t0 = time.perf_counter()
for t in ts: t.start()
for t in ts: t.join()
dt = time.perf_counter() - t0
print(f"streams={streams}: {streams * evals / dt:.2f} evals/s")
P
for cfg in "" "GPB_LOOKAHEAD=1" "GPB_FINE_BELOW=148" "GPB_TAIL_BLOCKS=32" "GPB_STASH=0" "GPB_LOOKAHEAD=1 GPB_FINE_BELOW=148 GPB_TAIL_BLOCKS=32 GPB_STASH=0"; do
  echo "== ${cfg:-default}"
  for s in 1 2; do env $cfg python /tmp/cr.py $s; done
done
