"""Developer check on a GPU box: GEMM kernel vs torch, factorisation / NLL / predict vs the oracle, timings."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle
from profit_b200 import GPBackend, KIND_RBF, KIND_MATERN52

stage = sys.argv[1] if len(sys.argv) > 1 else "all"
be = GPBackend(0)
dev = torch.device("cuda:0")
out = {}

if stage in ("all", "gemm"):
    for cfg in (0, 1, 2):
        M = N = K = 512
        A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
        C = torch.zeros(M, N, dtype=torch.float64, device=dev)
        be.debug_gemm(A, B, C, cfg=cfg, reps=1)
        torch.cuda.synchronize()
        ref = A @ B.T
        err = (C - ref).abs().max().item() / ref.abs().max().item()
        print(f"gemm cfg{cfg} 512^3 rel err {err:.2e}", flush=True)
        assert err < 1e-13
    for n in (4096, 8192):
        A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
        C = torch.zeros(n, n, dtype=torch.float64, device=dev)
        for cfg in (0, 1, 2):
            ms = be.debug_gemm(A, B, C, cfg=cfg, reps=3)
            print(f"gemm cfg{cfg} {n}^3: {ms:.2f} ms {2*n**3/ms*1e-9:.2f} TF", flush=True)
            out[f"gemm_cfg{cfg}_{n}"] = 2*n**3/ms*1e-9
        del A, B, C
    # short-K (trailing-update like) 16384 x 16384 x 512 and x128
    n = 16384
    for k in (128, 512):
        A = torch.randn(n, k, dtype=torch.float64, device=dev); C = torch.zeros(n, n, dtype=torch.float64, device=dev)
        for cfg in (0, 1):
            ms = be.debug_gemm(A, A, C, cfg=cfg, reps=3)
            print(f"gemm cfg{cfg} {n}x{n}x{k}: {ms:.2f} ms {2*n*n*k/ms*1e-9:.2f} TF", flush=True)
        del A, C

if stage in ("all", "chol"):
    rng = np.random.default_rng(0)
    for n in (64, 128, 200, 384, 1000, 2048):
        Q = rng.standard_normal((n, n)); A = Q @ Q.T / n + np.eye(n)
        L = be.cholesky(A); Lr = np.linalg.cholesky(A)
        e1 = np.abs(L - Lr).max() / np.abs(Lr).max()
        Ki = be.invert_cholesky(Lr); Kr = np.linalg.inv(A)
        e2 = np.abs(Ki - Kr).max() / np.abs(Kr).max()
        b = rng.standard_normal((n, 3)); x = be.solve_cholesky(Lr, b); xr = np.linalg.solve(A, b)
        e3 = np.abs(x - xr).max() / np.abs(xr).max()
        print(f"chol n={n}: L {e1:.2e} inv {e2:.2e} solve {e3:.2e}", flush=True)
        assert max(e1, e2, e3) < 1e-11
    try:
        be.cholesky(-np.eye(10)); raise SystemExit("expected LinAlgError")
    except np.linalg.LinAlgError as e:
        print("non-SPD ->", e)

if stage in ("all", "nll"):
    for (N, D) in ((3, 2), (200, 2), (300, 5), (1000, 8), (1500, 10), (1100, 16), (700, 20), (520, 32)):
        X, y = oracle.synthetic_problem(N, D)
        Xc = np.random.default_rng(2).random((777, D))
        for n_ell in (D, 1):
            theta = np.concatenate([np.linspace(0.6, 1.2, D) if n_ell == D else [0.8], [1.5, 0.3]])
            for kind, kern in ((KIND_RBF, oracle.rbf), (KIND_MATERN52, oracle.matern52)):
                be.set_train(X, y)
                nll0 = be.nll(kind, theta)
                nll, g = be.nll(kind, theta, want_grad=True)
                rn, rg = oracle.nll_cholesky(theta, X, y, kern, eval_gradient=True)
                e_n = abs(nll - rn) / abs(rn); e_g = np.max(np.abs(g - rg) / np.abs(rg))
                mean, var = be.predict(Xc)
                ell = theta[:-2]
                rm, rv = oracle.predict(X, y, Xc, kern, ell if n_ell > 1 else ell[0], theta[-2], theta[-1])
                e_m = np.max(np.abs(mean - rm)) / np.abs(rm).max(); e_v = np.max(np.abs(var - rv.ravel()) / rv.ravel())
                print(f"N={N} D={D} n_ell={n_ell} kind={kind}: nll {e_n:.1e} (nograd {abs(nll0-rn)/abs(rn):.1e}) grad {e_g:.1e} mean {e_m:.1e} var {e_v:.1e}", flush=True)
                assert e_n < 1e-9 and e_g < 1e-8 and e_m < 1e-8 and e_v < 1e-8

if stage == "big":
    for (N, D) in ((4096, 16), (8192, 8), (8192, 16)):
        X, y = oracle.synthetic_problem(N, D)
        theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
        from profit_b200 import kernels
        K = be.kernel_matrix(KIND_RBF, X, X, theta[:D], 1.5, 0.3)
        Kt = torch.from_numpy(K).to(dev)
        Lt = torch.linalg.cholesky(Kt)
        print(N, D, "torch chol ok, min diag", Lt.diagonal().min().item(), "sym err", (Kt-Kt.T).abs().max().item(), flush=True)
        try:
            L = be.cholesky(K)
            print("  gpb chol err", np.abs(L - Lt.cpu().numpy()).max(), flush=True)
        except Exception as e:
            print("  gpb chol FAILED", e, flush=True)
        be.set_train(X, y)
        try:
            print("  nll", be.nll(KIND_RBF, theta), flush=True)
        except Exception as e:
            print("  nll FAILED", e, flush=True)

if stage in ("all", "assembly"):
    for (N, D, kind) in ((8192, 10, KIND_MATERN52), (16384, 10, KIND_MATERN52), (16384, 10, KIND_RBF), (8192, 16, KIND_RBF)):
        X = np.random.default_rng(0).random((N, D))
        ell = np.linspace(0.6, 1.2, D); P = D + 2
        dKt = torch.empty(N, N, P, dtype=torch.float64, device=dev); Kt = torch.empty(N, N, dtype=torch.float64, device=dev)
        for rep in range(3):
            be.kernel_matrix(kind, X, X, ell, 1.5, 0.3, eval_gradient=True, out_K=Kt, out_dK=dKt)
            ms = be.stage_ms()["assemble"]
        gb = N * N * (1 + P) * 8 / 1e9
        be.kernel_matrix(kind, X, X, ell, 1.5, 0.3, eval_gradient=False, out_K=Kt)
        ms0 = be.stage_ms()["assemble"]
        print(f"assembly N={N} D={D} kind={kind}: K+dK {ms:.3f} ms = {gb/ms*1e3:.0f} GB/s ({gb:.1f} GB);  K only {ms0:.3f} ms = {N*N*8/1e9/ms0*1e3:.0f} GB/s", flush=True)
        # spot check against the oracle on a corner
        kern = oracle.rbf if kind == KIND_RBF else oracle.matern52
        rK, rdK = kern(X[:70], X[N-90:], ell, 1.5, 0.3, eval_gradient=True)
        print("   err", (Kt[:70, N-90:].cpu().numpy() - rK).__abs__().max(), np.abs(dKt[:70, N-90:].cpu().numpy() - rdK).max(), flush=True)
        del dKt, Kt

if stage == "time4096":
    X, y = oracle.synthetic_problem(4096, 8)
    theta = np.concatenate([np.linspace(0.6, 1.2, 8), [1.5, 0.3]])
    be.set_train(X, y)
    be.nll(KIND_RBF, theta, want_grad=True)
    be.nll(KIND_RBF, theta, want_grad=True)
    print(be.stage_ms())

if stage in ("all", "time"):
    for (N, D, kind) in ((4096, 8, KIND_RBF), (8192, 16, KIND_RBF), (16384, 10, KIND_MATERN52)):
        X, y = oracle.synthetic_problem(N, D)
        theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
        be.set_train(X, y)
        be.nll(kind, theta, want_grad=True)
        t0 = time.perf_counter(); nll, g = be.nll(kind, theta, want_grad=True); t1 = time.perf_counter()
        st = be.stage_ms()
        print(f"N={N} D={D}: nll+grad {1e3*(t1-t0):.1f} ms wall; stages {json.dumps({k: round(v,2) for k,v in st.items()})}", flush=True)
        flops = N**3/3
        print(f"   potrf {flops/st['potrf']*1e-9:.1f} TF, trtri {flops/st['trtri']*1e-9:.1f} TF, syrk {flops/st['syrk']*1e-9:.1f} TF, total {3*flops/st['total']*1e-9:.1f} TF", flush=True)
        t0 = time.perf_counter(); be.nll(kind, theta); t1 = time.perf_counter()
        print(f"   nll only {1e3*(t1-t0):.1f} ms", flush=True)
        out[f"nllgrad_ms_{N}"] = st
    # predict throughput sample
    N, D, M = 8192, 10, 1 << 17
    X, y = oracle.synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    be.set_train(X, y); be.factorize(KIND_MATERN52, theta)
    Xc = np.random.default_rng(2).random((M, D))
    be.predict(Xc[:16384])
    t0 = time.perf_counter(); be.predict(Xc); t1 = time.perf_counter()
    st = be.stage_ms()
    print(f"predict N={N} M={M}: {t1-t0:.3f} s -> {M/(t1-t0):.0f} cand/s; var GEMM {M*N*N/st['var']*1e-9:.1f} TF; stages {st}", flush=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
print("gpu_check done")
