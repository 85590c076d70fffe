"""Small, fixed workloads for the ncu captures under profiles/ (tools/ncu_capture.sh):
  nll      one NLL+gradient evaluation, Matern-5/2 N=16384 D=10 (after one warm-up evaluation)
  kmat     dense K + dK/dtheta (N, N, 12), Matern-5/2 N=16384 D=10, device resident (two calls)
  predict  65536 candidates against N=8192 (cross kernel + mean, column-norm GEMM, finish), after a warm-up chunk"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_MATERN52

mode = sys.argv[1]
D = 10
theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
be = GPBackend(0)
if mode == "nll":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    X, y = synthetic_problem(N, D)
    be.set_train(X, y)
    be.nll(KIND_MATERN52, theta, want_grad=True)
    print("launches before the captured evaluation:", be.launch_count())
    nll, g = be.nll(KIND_MATERN52, theta * 1.01, want_grad=True)
    print("launches after:", be.launch_count(), "stages", be.stage_ms())
elif mode == "kmat":
    import torch

    N = 16384
    X = torch.from_numpy(np.random.default_rng(0).random((N, D))).cuda()
    K = torch.empty(N, N, dtype=torch.float64, device="cuda")
    dK = torch.empty(N, N, D + 2, dtype=torch.float64, device="cuda")
    for _ in range(2):
        be.kernel_matrix(KIND_MATERN52, X, X, theta[:-2], theta[-2], theta[-1], eval_gradient=True, out_K=K, out_dK=dK)
        print("assemble ms", be.stage_ms()["assemble"])
elif mode == "predict":
    N, M = 8192, 1 << 16
    X, y = synthetic_problem(N, D)
    be.set_train(X, y)
    be.factorize(KIND_MATERN52, theta)
    Xc = np.random.default_rng(2).random((M, D))
    be.predict(Xc[:4096])
    print("launches before the captured chunk:", be.launch_count())
    be.predict(Xc)
    print("launches after:", be.launch_count(), be.stage_ms())
be.close()
