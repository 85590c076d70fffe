"""A/B of the task order of the predict column-norm GEMM (GPB_PRED_SUPER_M / _C, read in gpb_create): one process per
setting, 2^17 candidates against N=8192; prints ms per 65536-candidate chunk and checks the variances agree bit for bit
with the plain order."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from _data import synthetic_problem
    from profit_b200 import GPBackend, KIND_MATERN52
    N, D, M = 8192, 10, 1 << 17
    X, y = synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    be = GPBackend(0); be.set_train(X, y); be.factorize(KIND_MATERN52, theta)
    import torch
    Xc = torch.from_numpy(np.random.default_rng(2).random((M, D))).cuda()
    mean = torch.empty(M, 1, dtype=torch.float64, device="cuda"); var = torch.empty(M, dtype=torch.float64, device="cuda")
    be.predict(Xc, out_mean=mean, out_var=var)
    ms = []
    for _ in range(3):
        be.predict(Xc, out_mean=mean, out_var=var); ms.append(be.stage_ms()["var"])
    import hashlib
    print(f"M={os.environ.get('GPB_PRED_SUPER_M','default')} C={os.environ.get('GPB_PRED_SUPER_C','default')}: variance stage {min(ms):.2f} ms "
          f"for {M} candidates = {M * N**2 / (min(ms) * 1e-3) * 1e-12:.2f} TFLOP/s; sha1(var) {hashlib.sha1(var.cpu().numpy().tobytes()).hexdigest()[:12]}", flush=True)
else:
    for m, c in ((0, 1), (12, 12), (8, 18), (16, 9), (6, 24), (24, 6), (12, 24), (16, 16)):
        env = dict(os.environ, GPB_PRED_SUPER_M=str(m), GPB_PRED_SUPER_C=str(c))
        subprocess.run([sys.executable, __file__, "child"], env=env)
