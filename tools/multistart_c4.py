"""BASELINE configs[4]: multi-start hyper-parameter restarts, N=8192, D=16, restarts sharded over ranks (and over
concurrent streams per GPU), NCCL/gloo gather of the best NLL.  Usage (1 GPU):  python tools/multistart_c4.py 8 2
(restarts, streams); under torchrun the restarts are sharded r -> rank r mod G."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from profit_b200 import GPBackend, RBF
from profit_b200.distributed import multistart_optimize, world

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N, D = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (8192, 16)
if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dist.init_process_group("nccl")
rank, ws = world()
X, y = bench.synthetic_problem(N, D)
th = bench.theta_ref(D)
# deterministic start list: theta_ref * 10^(h - 1/2) * 0.6 with a van-der-Corput style sequence
def vdc(i, base):
    f, r = 1.0, 0.0
    while i > 0:
        f /= base; r += f * (i % base); i //= base
    return r
primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61]
H = np.array([[vdc(r + 1, primes[p]) for p in range(th.size)] for r in range(R)])
starts = th * 10 ** (0.6 * (H - 0.5))
be = GPBackend(int(os.environ.get("LOCAL_RANK", "0")))
t0 = time.perf_counter()
best, best_nll, table = multistart_optimize(X, y, starts, RBF, eval_gradient=True, backend=be, streams=streams)
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"workload": f"{R} multi-start BFGS restarts, N={N}, D={D}, RBF ARD", "n_gpus": ws, "streams_per_gpu": streams,
                      "seconds": dt, "restarts_per_s": R / dt, "best_nll": best_nll, "distinct_optima": int(len(set(np.round(table[:, 1], 3))))}))
