import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from profit_b200 import GPBackend
be = GPBackend(0)
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, K) in ((8192, 8192, 512), (16384, 16384, 128), (6144, 6144, 512)):
    A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
    C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
    ref = C0 - A @ B.T
    for cfg in (0, 1, 2):
        for rep in range(3):
            C = C0.clone()
            torch.cuda.synchronize()
            be.debug_gemm(A, B, C, cfg=cfg, reps=1, alpha=-1.0, beta=1.0)
            torch.cuda.synchronize()
            err = (C - ref).abs()
            nb = (err > 1e-9).sum().item()
            te = err.reshape(M // 128, 128, N // 64, 64).amax(dim=(1, 3))
            bad = (te > 1e-9).nonzero()[:6].tolist()
            print(f"{M}x{N}x{K} cfg{cfg} rep{rep}: max err {err.max().item():.2e} bad elems {nb} first bad 128x64 tiles {bad}", flush=True)
    del A, B, C0, ref, C
