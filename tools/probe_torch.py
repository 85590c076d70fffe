"""Library-side FP64 reference points on the box (cuBLAS DGEMM, cuSOLVER potrf/potri). Not product code."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: torch.matmul(a, b.T))
    res[f"dgemm_nt_{n}"] = {"ms": ms, "tflops": 2 * n**3 / ms * 1e-9}
    ms = timeit(lambda: torch.matmul(a, b))
    res[f"dgemm_nn_{n}"] = {"ms": ms, "tflops": 2 * n**3 / ms * 1e-9}
    del a, b
# long-K skinny (trailing-update-like): C[16384x16384] -= A[16384x256] B^T
n, k = 16384, 256
a = torch.randn(n, k, dtype=torch.float64, device=dev); c = torch.randn(n, n, dtype=torch.float64, device=dev)
ms = timeit(lambda: torch.addmm(c, a, a.T, beta=1, alpha=-1, out=c))
res["dgemm_rankk_16384_k256"] = {"ms": ms, "tflops": 2 * n * n * k / ms * 1e-9}
del a, c
for n in (4096, 8192, 16384):
    x = torch.rand(n, 10, dtype=torch.float64, device=dev)
    d = torch.cdist(x, x)
    K = 2.25 * torch.exp(-0.5 * d * d) + 0.09 * torch.eye(n, dtype=torch.float64, device=dev)
    del d
    ms = timeit(lambda: torch.linalg.cholesky(K))
    res[f"cusolver_potrf_{n}"] = {"ms": ms, "tflops": n**3 / 3 / ms * 1e-9}
    L = torch.linalg.cholesky(K)
    ms = timeit(lambda: torch.cholesky_inverse(L), reps=2)
    res[f"cusolver_potri_{n}"] = {"ms": ms, "tflops": 2 * n**3 / 3 / ms * 1e-9}
    del K, L
# HBM copy
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
ms = timeit(lambda: b.copy_(a), reps=10)
res["copy_gbs"] = 2 * a.numel() * 2 / ms * 1e-6
print(json.dumps(res, indent=1))
open("gpurun_out/probe_torch.json", "w").write(json.dumps(res, indent=1))
