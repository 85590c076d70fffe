"""BASELINE configs[3] through the acquisition classes: one active-learning pick over 2^22 candidates against N=8192
training points, candidates sharded over the ranks of a torchrun job (NCCL carries two 16-byte gathers per pick).

    python -m torch.distributed.run --nproc-per-node G tools/al_sharded_demo.py [log2_M] [kind]
"""
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from _data import synthetic_problem
from profit_b200 import B200GPSurrogate
from profit_b200 import acquisition as acq

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    dist.init_process_group("nccl")
rank = dist.get_rank() if dist.is_initialized() else 0
ws = dist.get_world_size() if dist.is_initialized() else 1
M = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
kind = sys.argv[2] if len(sys.argv) > 2 else "expected_improvement"
N, D = 8192, 10
X, y = synthetic_problem(N, D)
theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
Xpred = np.random.default_rng(2).random((M, D))
s = B200GPSurrogate(local)
s.Xtrain, s.ytrain, s.ndim, s.trained = X, y, D, True
s.kernel = s.select_kernel("Matern52")
s.hyperparameters = {"length_scale": theta[:-2].copy(), "sigma_f": theta[-2:-1].copy(), "sigma_n": theta[-1:].copy()}
s.optimize = lambda *a, **k: None                    # time the pick, not a refit
variables = types.SimpleNamespace(input=X, output=y)
af = acq._TWINS[kind](Xpred, s, variables)
af.find_next_candidates(1)                           # warm-up: upload the shard, factorise, allocate
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
t0 = time.perf_counter()
cand = af.find_next_candidates(2)
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / 2
if rank == 0:
    rec = {"workload": f"{kind} pick over 2^{int(np.log2(M))} candidates, N={N}, D={D} Matern-5/2, hyper-parameters frozen",
           "n_gpus": ws, "seconds_per_pick": round(dt, 4), "candidates_per_s": round(M / dt),
           "picks": np.round(cand, 6).tolist(), "n_train_after": int(s.Xtrain.shape[0])}
    print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/al_sharded_{kind}_g{ws}.json", "w") as f:
        json.dump(rec, f)
if ws > 1:
    dist.destroy_process_group()
