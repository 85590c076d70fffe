#!/usr/bin/env python
"""Headline benchmark: NLL+gradient evaluations per second of the Matern-5/2 ARD GP, N=16384, D=10, fp64
(BASELINE.json metric, configs[2]).  The same JSON line carries the other BASELINE configs as extras:
"al_predict" = predict_f mean/variance of ONE 2^22-candidate set against N=8192 training points (configs[3]), sharded
over the ranks (strong scaling); "c4_multistart" = 64 BFGS restarts at N=8192, D=16 sharded over the ranks
(configs[4]); "small" = the complete fits of configs[0] (N=200) and configs[1] (N=4096, D=8).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

One "step" = one NLL+gradient evaluation (Cholesky + triangular inverse + K^-1 + gradient contraction).
N > 1 (torchrun, one rank per GPU): the path does not shard inside one evaluation, so ranks evaluate
independent multi-start restarts (weak scaling) and NCCL only gathers (nll, theta) for the best-NLL pick;
candidate prediction shards candidates and gathers the arg-max.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 16384, 10                 # BASELINE configs[2]
N_TRAIN_PRED, M_FULL = 8192, 1 << 22     # BASELINE configs[3]
N_TRAIN_C4, DIM_C4 = 8192, 16            # BASELINE configs[4]
METRIC = "NLL+grad evals/s (N=16384, D=10, fp64)"
UNIT = "evals/s"
NCU_DRAM_BYTES_SYRK_LAUNCH = 15.49e9     # profiles/r02_ncu_syrk.txt (dram__bytes_read.sum 14.38 GB + dram__bytes_write.sum 1.11 GB)
NCU_DRAM_BYTES_COLSQ_LAUNCH = 28.26e9    # profiles/r02_ncu_predict_colsq.txt (28.11 GB read + 0.15 GB written per launch)
NCU_DRAM_BYTES_COLSQ_NOTE = ("dram__bytes_read+write of ONE column-norm GEMM launch (32768 candidates x N=8192, 2.2e12 flop, "
                             "61.6 ms, DMMA sub-pipe 97.5 % active) with the super-tile task order; plain order 71.1 GB "
                             "(profiles/r02_ncu_predict_colsq_plain_order.txt); algorithmic minimum 2.4 GB (K* chunk + W)")
DMMA_CEILING_TFLOPS = 37.05              # profiles/r01_probe_fp64_dmma.txt: register-resident DMMA.8x8x4 issue rate


def synthetic_problem(N, D, seed_x=0, seed_y=1, sigma_n=0.3):
    """Deterministic inputs of SURVEY.md section 8d (same generator as oracle.synthetic_problem; the GPU legs do
    not import the oracle): X ~ U[0,1)^(N,D), y = sin 3x0 + cos 2x1 + sum_{d>=2} x_d + sigma_n N(0,1)."""
    X = np.random.default_rng(seed_x).random((N, D))
    f = np.sin(3 * X[:, 0]) + np.cos(2 * X[:, 1]) + X[:, 2:].sum(axis=1)
    y = f + sigma_n * np.random.default_rng(seed_y).standard_normal(N)
    return X, y.reshape(-1, 1)


def theta_ref(D):
    return np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def max_over_ranks(x, dist, local):
    if dist is None:
        return x
    import torch

    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist):
    import torch

    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ CPU leg
def _host_threads():
    from threadpoolctl import threadpool_info

    return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])


def cpu_nll_grad_eval(n, want_timings=False):
    """ONE NLL+gradient evaluation of the oracle port of gp_functions.negative_log_likelihood_cholesky (row-blocked,
    O(N^2 P) gradient contraction) at N=n, D=10, Matern-5/2 on all host cores.  Returns (seconds, cores[, phases])."""
    import oracle
    from threadpoolctl import threadpool_limits

    X, y = oracle.synthetic_problem(n, DIM)
    th = theta_ref(DIM)
    # torchrun exports OMP_NUM_THREADS=1; give BLAS every host core this process may use
    with threadpool_limits(limits=len(os.sched_getaffinity(0))):
        cores = _host_threads()
        tm = {}
        t0 = time.perf_counter()
        oracle.nll_cholesky(th, X, y, oracle.matern52, eval_gradient=True, timings=tm)
        dt = time.perf_counter() - t0
    return (dt, cores, tm) if want_timings else (dt, cores)


def cpu_nll_grad_sample(n_sample):
    """cpu_baseline of the GPU arm: a BOUNDED sample (one oracle-port evaluation at N=n_sample) scaled to N=16384 phase
    by phase -- Cholesky / solves / inverse with (N/n)^3, assembly and contraction with (N/n)^2 -- and labelled as an
    extrapolation.  The measured (not extrapolated) N=16384 figure is what `bench.py --impl reference` reports."""
    dt, cores, tm = cpu_nll_grad_eval(n_sample, want_timings=True)
    r = N_TRAIN / n_sample
    full = tm["cubic"] * r**3 + tm["quadratic"] * r**2
    return {"value": 1.0 / full, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
            "sample": f"one NLL+grad eval of the oracle port at N={n_sample}, D={DIM} Matern-5/2: {dt:.2f} s "
                      f"({tm['cubic']:.2f} s O(N^3) phases scaled by {r**3:.0f}, {tm['quadratic']:.2f} s O(N^2 P) phases "
                      f"scaled by {r**2:.0f}) -> {full:.0f} s per eval at N=16384 (EXTRAPOLATED; `--impl reference` "
                      f"measures a complete N=16384 evaluation); the unmodified reference cannot hold N=16384 "
                      f"(>70 GB of (N,N,D) temporaries)",
            "sample_seconds": dt, "scaled_seconds": full}


def cpu_predict_sample(n_train=4096, chunk=1024, chunks=2):
    """CPU baseline of the candidate-prediction leg: the oracle port of GPSurrogate.predict on a
    bounded sample, both as the reference behaves (every call rebuilds alpha and K^-1, custom_surrogate.py:24-45,
    114-116) and with the factor cached once.  The sample runs at N=n_train and is scaled to N=8192: factorisation with
    (N/n)^3, the per-candidate work (kernel row + K^-1 quadratic form) with (N/n)^2."""
    import oracle
    from threadpoolctl import threadpool_limits

    X, y = oracle.synthetic_problem(n_train, DIM)
    th = theta_ref(DIM)
    ell, sf, sn = th[:-2], th[-2], th[-1]
    Xc = np.random.default_rng(2).random((chunk * chunks, DIM))
    ncpu = len(os.sched_getaffinity(0))
    with threadpool_limits(limits=ncpu):
        cores = _host_threads()
        t0 = time.perf_counter()
        alpha, Kinv = oracle.predict_cached_factor(X, y, oracle.matern52, ell, sf, sn)
        t_factor = time.perf_counter() - t0
        t0 = time.perf_counter()
        for c in range(chunks):
            Ks = oracle.matern52(X, Xc[c * chunk:(c + 1) * chunk], ell, sf)
            _mean = Ks.T @ alpha
            _var = np.maximum(sf**2 - np.einsum("ic,ic->c", Ks, Kinv @ Ks), 1e-10) + sn**2
        t_chunk = (time.perf_counter() - t0) / chunks
    r = N_TRAIN_PRED / n_train
    cached = chunk / (t_chunk * r**2)
    uncached = chunk / (t_chunk * r**2 + t_factor * r**3)
    return {"value": cached, "unit": "candidates/s", "cores": cores, "kind": "port", "extrapolated": True,
            "uncached_value": uncached,
            "sample": f"oracle port of GPSurrogate.predict at N={n_train}, D={DIM} Matern-5/2: factor (alpha, K^-1) "
                      f"{t_factor:.2f} s, {chunks} chunks of {chunk} candidates {t_chunk:.2f} s each; scaled to N={N_TRAIN_PRED} "
                      f"(factor x{r**3:.0f}, chunk x{r**2:.0f}): {cached:.0f} candidates/s with the factor cached, "
                      f"{uncached:.1f} candidates/s when every {chunk}-candidate call refactorises like the reference"}


def unmodified_reference_sample(n=2048):
    """If the unmodified reference is installed (baseline/_ref, git-ignored), time ITS
    negative_log_likelihood_cholesky(eval_gradient=True) once on a small RBF problem.  Informational: the reference has
    no Matern-5/2 kernel and its O(N^3 P) trace cannot reach N=16384 in memory or time, so the headline CPU number
    comes from the oracle port, which is the faster CPU algorithm."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "profit")):
        return None
    try:
        sys.path.insert(0, ref)
        from profit.sur.gp.backend.gp_functions import negative_log_likelihood_cholesky
        from profit.sur.gp.backend.python_kernels import RBF
        from threadpoolctl import threadpool_limits

        X, y = synthetic_problem(n, DIM)
        with threadpool_limits(limits=len(os.sched_getaffinity(0))):
            t0 = time.perf_counter()
            negative_log_likelihood_cholesky(theta_ref(DIM), X, y, RBF, eval_gradient=True)
            dt = time.perf_counter() - t0
        return {"what": f"unmodified profit negative_log_likelihood_cholesky(eval_gradient=True), RBF ARD, N={n}, D={DIM}",
                "seconds": dt, "evals_per_s": 1.0 / dt}
    except Exception as exc:  # the reference is optional here
        return {"error": repr(exc)}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port; the unmodified reference has no Matern-5/2 kernel and
    cannot hold N=16384) on the host cores.  `value` is MEASURED: complete N=16384 evaluations, as many as fit the time
    budget (at least one, first in the timed region).  The remaining steps are bounded N=2048 samples of the
    same workload (they only cross-check the phase-wise scaling law).  `ms_per_step` is the true wall time of the
    timed region divided by the step count, so steps x ms_per_step is what the run really took."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("GPB_REF_BUDGET_S", "300"))
    n_bounded = 2048                                   # bounded-sample steps: ~1.5 s each on 8 cores
    cpu_nll_grad_eval(1024)                            # page in numpy / BLAS
    for _ in range(args.warmup):
        cpu_nll_grad_eval(2048)
    t_region = time.perf_counter()
    full, samples, cores = [], [], 1
    for k in range(args.steps):
        spent = time.perf_counter() - t_region
        if k == 0 or (spent + 1.15 * max(full) <= budget):
            dt, cores = cpu_nll_grad_eval(N_TRAIN)
            full.append(dt)
        else:
            samples.append(cpu_nll_grad_sample(n_bounded))
    region = time.perf_counter() - t_region
    sec = statistics.mean(full)
    sample = (f"{len(full)} complete NLL+grad evaluation(s) of the oracle port at N={N_TRAIN}, D={DIM} Matern-5/2, measured: "
              + ", ".join(f"{t:.1f}" for t in full) + f" s on {cores} threads (value = 1 / mean)")
    if samples:
        est = statistics.mean(s["scaled_seconds"] for s in samples)
        sample += (f"; the other {len(samples)} steps are bounded N={n_bounded} samples "
                   f"({statistics.mean(s['sample_seconds'] for s in samples):.2f} s each) whose phase-wise "
                   f"extrapolation gives {est:.0f} s per eval")
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": region / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "value_is": "1 / (measured seconds of a complete N=16384 evaluation); ms_per_step is the wall time of the "
                        "timed region / steps (full evaluations + bounded samples), NOT 1000 / value",
            "full_eval_seconds": full, "full_evals": len(full), "bounded_sample_steps": len(samples),
            "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "unmodified_reference_sample": unmodified_reference_sample()}
    if samples:
        line["extrapolated_seconds_from_bounded_samples"] = statistics.mean(s["scaled_seconds"] for s in samples)
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "Matern-5/2 ARD GP NLL+gradient, N=16384, D=10, P=12 (BASELINE configs[2])",
            "n_train": N_TRAIN, "dim": DIM, "kernel": "Matern52", "theta": "l=linspace(0.6,1.2,10), sf=1.5, sn=0.3",
            "l2": "inputs larger than L2: every evaluation streams the 2.1 GB factor (and two more 2.1 GB "
                  "triangles for the inverse) through HBM, 126 MB L2",
            "parallelism": "replicas (independent multi-start restarts per GPU; NCCL gathers best NLL only)"}


# ------------------------------------------------------------------------------------------------ GPU legs
def measured_hbm_peak():
    """HBM roofline denominator: the driver-measured copy bandwidth, else the profiling guide's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"


def fp64_gemm_peak(local):
    """cuBLAS DGEMM 8192^3, best of 5 after warm-up -- the measured FP64 tensor roofline denominator
    (MEASURED_PEAKS.json carries HBM and bf16 only)."""
    import torch

    dev = torch.device("cuda", local)
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b.T)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b.T)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n**3 / best * 1e-9


def halton(n, dim):
    """Halton points, index 1..n, prime bases (same sequence as profit.util.halton.halton: halton(128, 2)[:3] =
    [[1/2, 1/3], [1/4, 2/3], [3/4, 1/9]]; restated here because bench.py cannot import the reference on the GPU box)."""
    primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101]
    out = np.empty((n, dim))
    for j in range(dim):
        base = primes[j]
        for i in range(n):
            k, f, r = i + 1, 1.0, 0.0
            while k > 0:
                f /= base
                r += f * (k % base)
                k //= base
            out[i, j] = r
    return out


def c4_starts(restarts):
    """theta0^(r) = theta_ref * 10^(h_r - 1/2), h = halton(R, P)."""
    th = theta_ref(DIM_C4)
    return th * 10 ** (halton(restarts, th.size) - 0.5)


def small_fits(local):
    """configs[0] (RBF fit + predict, N=200 Halton points in 2-D; the reference needs 0.272 s for the fit and 0.455 s
    for the 2500-candidate predict on 8 CPU cores) and configs[1] (complete RBF hyper-parameter
    fit, N=4096, D=8) through the surrogate / function-level API with host arrays."""
    from profit_b200 import RBF, B200GPSurrogate, GPBackend, gp_functions as gpf

    out = {}
    X0 = halton(200, 2)
    y0 = (np.sin(2 * X0[:, 0]) + np.cos(3 * X0[:, 1]) + 0.05 * np.random.default_rng(1).standard_normal(200)).reshape(-1, 1)
    g = np.linspace(0, 1, 50)
    Xg = np.stack(np.meshgrid(g, g), axis=-1).reshape(-1, 2)
    B200GPSurrogate().train(X0, y0)                                   # warm-up
    fit_ms, pred_ms = [], []
    for _ in range(3):
        s = B200GPSurrogate()
        t0 = time.perf_counter()
        s.train(X0, y0)
        fit_ms.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        s.predict(Xg)
        pred_ms.append((time.perf_counter() - t0) * 1e3)
    out["c0"] = {"workload": "B200GPSurrogate.train + predict(2500 grid points), RBF, N=200 2-D Halton (BASELINE configs[0])",
                 "train_ms": min(fit_ms), "predict_ms": min(pred_ms),
                 "reference_cpu_train_ms": 272.0, "reference_cpu_predict_ms": 455.0,
                 "reference_source": "BASELINE.md section 2 / SURVEY.md section 6: unmodified reference, 8 CPU cores"}
    N1, D1 = 4096, 8
    X1, y1 = synthetic_problem(N1, D1)
    th1 = theta_ref(D1)
    a0 = th1 * 10 ** (0.3 * (np.random.default_rng(5).random(D1 + 2) - 0.5))
    be = GPBackend(local)
    gpf.optimize(X1[:256], y1[:256], a0, RBF, eval_gradient=True, backend=be)
    be.set_train(X1, y1)
    be.nll(0, th1, want_grad=True)
    ev = []
    for _ in range(5):
        be.nll(0, th1, want_grad=True)
        ev.append(be.stage_ms()["total"])
    t0 = time.perf_counter()
    hyp, res = gpf.optimize(X1, y1, a0, RBF, eval_gradient=True, backend=be, return_result=True)
    dt = time.perf_counter() - t0
    out["c1"] = {"workload": "gp_functions.optimize (SciPy BFGS as gp_functions.py:59-66), RBF ARD, N=4096, D=8 "
                             "(BASELINE configs[1]); start = theta_ref * 10^U(-0.15, 0.15)",
                 "fit_seconds": dt, "nfev": int(res.nfev), "nit": int(res.nit),
                 "ms_per_eval_incl_host": dt / res.nfev * 1e3, "nll_grad_device_ms": min(ev),
                 "fp64_tflops_one_eval": float(N1) ** 3 / (min(ev) * 1e-3) * 1e-12,
                 "reference_cpu_seconds_per_eval": 34.1,
                 "reference_source": "SURVEY.md section 6: unmodified negative_log_likelihood_cholesky, N=4096 D=8, 8 cores"}
    be.close()
    return out


def run_gpu(args):
    import torch

    from profit_b200 import GPBackend, KIND_MATERN52, Matern52, gp_functions as gpf
    from profit_b200.distributed import gather_argmax, gather_argmin_rows, shard_bounds

    rank, world, local, dist = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    be = GPBackend(local)
    out = {}

    # ---------------- NLL + gradient (headline) ----------------
    X, y = synthetic_problem(N_TRAIN, DIM)
    th0 = theta_ref(DIM)
    # every rank/step gets its own hyper-parameters (multi-start restarts): nothing can be cached across steps
    def theta_of(step, r=rank):
        return th0 * 10 ** (0.05 * np.sin(1.0 + 0.7 * (r * 131 + step) + np.arange(th0.size)))

    be.set_train(X, y)
    for w in range(args.warmup):
        be.nll(KIND_MATERN52, theta_of(-1 - w), want_grad=True)
    stages = {k: 0.0 for k in ("assemble", "potrf", "trtri", "syrk", "grad", "total")}
    results = []
    barrier(dist)
    launches0 = be.launch_count()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for k in range(args.steps):
            nll, g = be.nll(KIND_MATERN52, theta_of(k), want_grad=True)
            results.append(np.concatenate([[rank * args.steps + k, nll], theta_of(k)]))
            st = be.stage_ms()
            for key in stages:
                stages[key] += st[key]
        torch.cuda.synchronize()
        elapsed = time.perf_counter() - t0
    launches = be.launch_count() - launches0
    barrier(dist)
    elapsed = max_over_ranks(elapsed, dist, local)
    best_row, _ = gather_argmin_rows(np.array(results))      # the only collective: (nll, theta) per restart
    value = world * args.steps / elapsed
    dense_ms = (stages["potrf"] + stages["trtri"] + stages["syrk"]) / args.steps

    # end to end: the reference-facing function with HOST buffers (pinned), H2D of (X, y, theta) and D2H of
    # (nll, grad) inside every step
    Xp = torch.from_numpy(X).pin_memory()
    yp = torch.from_numpy(y).pin_memory()
    Xh, yh = Xp.numpy(), yp.numpy()
    gpf.negative_log_likelihood(np.log10(theta_of(-9)), Xh, yh, Matern52, True, True, backend=None)
    barrier(dist)
    t0 = time.perf_counter()
    for k in range(args.steps):
        gpf.negative_log_likelihood(np.log10(theta_of(1000 + k)), Xh, yh, Matern52, True, True, backend=None)
    torch.cuda.synchronize()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0, dist, local)
    barrier(dist)
    P = th0.size
    e2e = {"value": world * args.steps / e2e_elapsed, "unit": UNIT,
           "h2d_bytes_per_step": int(X.nbytes + y.nbytes + P * 8), "d2h_bytes_per_step": int((P + 1) * 8 + 4),
           "api": "profit_b200.gp_functions.negative_log_likelihood(hyp_log10, X, y, Matern52, eval_gradient=True, "
                  "log_scale_hyp=True) with pinned host arrays"}

    # ---------------- candidate prediction: ONE 2^22-candidate set (configs[3]) sharded over the ranks --------
    pred = None
    if not args.skip_predict:
        Xt, yt = synthetic_problem(N_TRAIN_PRED, DIM)
        M_total = args.predict_candidates if args.predict_candidates > 0 else M_FULL
        lo, hi = shard_bounds(M_total, world, rank)
        Xc = np.ascontiguousarray(np.random.default_rng(2).random((M_total, DIM))[lo:hi])   # same set for every world size
        bp = GPBackend(local)
        bp.set_train(Xt, yt)
        bp.factorize(KIND_MATERN52, th0)
        dXc = torch.from_numpy(Xc).cuda()
        dmean = torch.empty(hi - lo, 1, dtype=torch.float64, device="cuda")
        dvar = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        bp.predict(dXc, out_mean=dmean, out_var=dvar)                       # warm-up (allocates the chunk buffers)
        psteps = max(1, min(args.steps, 3))
        barrier(dist)
        l0 = bp.launch_count()
        var_ms = cross_ms = 0.0
        t0 = time.perf_counter()
        for _ in range(psteps):
            bp.predict(dXc, out_mean=dmean, out_var=dvar)
            st = bp.stage_ms()
            var_ms += st["var"]
            cross_ms += st["cross"]
        torch.cuda.synchronize()
        p_elapsed = max_over_ranks(time.perf_counter() - t0, dist, local)
        p_launch = bp.launch_count() - l0
        li, lv = bp.argmax(dvar)
        gidx, gval = gather_argmax(lv, lo + li)
        # end to end from pinned host candidates to host mean/var
        Xcp = torch.from_numpy(Xc).pin_memory().numpy()
        hm = torch.empty(hi - lo, 1, dtype=torch.float64).pin_memory().numpy()
        hv = torch.empty(hi - lo, dtype=torch.float64).pin_memory().numpy()
        barrier(dist)
        t0 = time.perf_counter()
        for _ in range(psteps):
            bp.predict(Xcp, out_mean=hm, out_var=hv)
        pe_elapsed = max_over_ranks(time.perf_counter() - t0, dist, local)
        flop_per_cand = float(N_TRAIN_PRED) ** 2                            # triangular L^-1 k*: N^2 flop
        # whole timed step (cross kernel + variance GEMM + finish + chunk loop), all ranks
        tf_step = M_total * psteps * flop_per_cand / p_elapsed * 1e-12 / world
        tf_var = (hi - lo) * psteps * flop_per_cand / (var_ms * 1e-3) * 1e-12 if var_ms > 0 else None
        pred = {"metric": "AL candidate predictions/s", "unit": "candidates/s",
                "value": M_total * psteps / p_elapsed, "scaling": "strong",
                "e2e": {"value": M_total * psteps / pe_elapsed, "unit": "candidates/s",
                        "h2d_bytes_per_step": int(Xc.nbytes), "d2h_bytes_per_step": int(hm.nbytes + hv.nbytes)},
                "n_train": N_TRAIN_PRED, "candidates_per_step": int(M_total), "steps": psteps,
                "ms_per_step": p_elapsed / psteps * 1e3,
                "workload": ("2^22 candidates (BASELINE configs[3])" if M_total == M_FULL else
                             f"{M_total} candidates = bounded sample of the 2^22-candidate set of configs[3]")
                            + f", N={N_TRAIN_PRED}, D={DIM} Matern-5/2, one set sharded over {world} GPU(s)",
                "argmax": [int(gidx), float(gval)], "gpu_launches": int(p_launch),
                "step_tflops_per_gpu": tf_step, "variance_gemm_tflops": tf_var,
                "stage_ms_per_step_rank0": {"cross_mean": cross_ms / psteps, "variance": var_ms / psteps}}
        bp.close()
        del dXc, dmean, dvar
        torch.cuda.empty_cache()

    # ---------------- configs[4]: 64 multi-start BFGS restarts, N=8192, D=16, sharded over the ranks ----------
    c4 = None
    if not args.skip_c4:
        from profit_b200 import RBF
        from profit_b200.distributed import multistart_optimize

        Xm, ym = synthetic_problem(N_TRAIN_C4, DIM_C4)
        starts = c4_starts(args.c4_restarts)
        bm = GPBackend(local)
        gpf.optimize(Xm[:512], ym[:512], starts[0], RBF, eval_gradient=True, backend=bm)     # warm-up: SciPy, workspaces
        bm.set_train(Xm, ym)
        bm.nll(0, theta_ref(DIM_C4), want_grad=True)
        barrier(dist)
        l0 = bm.launch_count()
        t0 = time.perf_counter()
        best, best_nll, table = multistart_optimize(Xm, ym, starts, RBF, eval_gradient=True, backend=bm,
                                                    streams=args.c4_streams)
        torch.cuda.synchronize()
        c_elapsed = max_over_ranks(time.perf_counter() - t0, dist, local)
        c4 = {"metric": "multi-start restarts/s", "unit": "restarts/s", "value": args.c4_restarts / c_elapsed,
              "scaling": "strong", "seconds": c_elapsed, "restarts": int(args.c4_restarts),
              "workload": f"{args.c4_restarts} BFGS restarts (gp_functions.optimize, SciPy BFGS defaults), RBF ARD, "
                          f"N={N_TRAIN_C4}, D={DIM_C4}, P={DIM_C4 + 2} (BASELINE configs[4]); starts = theta_ref * "
                          f"10^(halton - 1/2); restarts handed out one at a time to the {world} rank(s) from a counter in the process "
                          f"group's store; {args.c4_streams} concurrent restart "
                          f"stream(s) per GPU; one all_gather of (nll, theta) picks the best",
              "best_nll": float(best_nll), "best_restart": int(table[np.argmin(table[:, 1]), 0]),
              "finite_restarts": int(np.isfinite(table[:, 1]).sum()),
              "distinct_optima": int(len(set(np.round(table[np.isfinite(table[:, 1]), 1], 3)))),
              "gpu_launches_rank0_first_stream": int(bm.launch_count() - l0),
              "evaluations_rank0": int(getattr(multistart_optimize, "last_evaluations", 0)),
              "restart_queue": getattr(multistart_optimize, "last_queue_mode", None),
              "evaluations_note": "NLL+gradient evaluations rank 0 spent on its restarts: the BFGS path lengths (SciPy's "
                                  "line searches, sensitive to the last bits of the gradient) decide the wall time as "
                                  "much as the 18 ms per evaluation do"}
        c4["ms_per_evaluation_rank0"] = (c_elapsed * 1e3 / c4["evaluations_rank0"]) if c4["evaluations_rank0"] else None
        bm.close()
        torch.cuda.empty_cache()

    # ---------------- configs[0] and configs[1]: complete fits at the small end (rank 0) -----------------------
    small = None
    if rank == 0 and not args.skip_small:
        small = small_fits(local)

    # ---------------- dense K + dK/dtheta assembly (python_kernels.RBF twin), HBM-bound ----------------
    asm = None
    if rank == 0 and not args.skip_assembly:
        P = th0.size
        Kt = torch.empty(N_TRAIN, N_TRAIN, dtype=torch.float64, device="cuda")
        dKt = torch.empty(N_TRAIN, N_TRAIN, P, dtype=torch.float64, device="cuda")
        dX = torch.from_numpy(X).cuda()
        ms = []
        for rep in range(args.warmup + 3):
            be.kernel_matrix(KIND_MATERN52, dX, dX, th0[:-2], th0[-2], th0[-1], eval_gradient=True, out_K=Kt, out_dK=dKt)
            ms.append(be.stage_ms()["assemble"])
        ms = statistics.mean(ms[args.warmup:])
        nbytes = float(N_TRAIN) ** 2 * (1 + P) * 8
        hbm_peak, hbm_src = measured_hbm_peak()
        asm = {"workload": f"K and dK/dtheta (N,N,{P}) of the Matern-5/2 ARD kernel, N={N_TRAIN}, D={DIM}, device resident",
               "ms": ms, "algorithmic_bytes": nbytes,
               "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": nbytes / (ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None, "peak_source": hbm_src,
                            "kernel": "gpb::kernel_matrix_kernel (bulk-TMA stores of the derivative tiles)"}}
        del Kt, dKt, dX
        torch.cuda.empty_cache()

    if rank != 0:
        be.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peak = fp64_gemm_peak(local)
    achieved = float(N_TRAIN) ** 3 / (dense_ms * 1e-3) * 1e-12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_SYRK_LAUNCH,
                "traffic_note": "dram__bytes_read+write of the K^-1 = U U^T launch (1.466e12 flop) from "
                                "profiles/r02_ncu_syrk.txt (super-tile task order; round 1: 37.3 GB); algorithmic minimum 2.1 GB (read U, write K^-1); compute-bound, DMMA sub-pipe 97.3 % active",
                "kernel": "gpb::dgemm_tasks_kernel (DMMA.8x8x4 tile GEMM behind potrf / trtri / syrk)",
                "algorithmic_flops_per_eval": float(N_TRAIN) ** 3,
                "dense_ms_per_eval": dense_ms,
                "peak_source": "cuBLAS DGEMM fp64 8192^3 measured in this run (MEASURED_PEAKS.json has no fp64 "
                               f"entry); DMMA.8x8x4 issue ceiling {DMMA_CEILING_TFLOPS} TFLOP/s",
                # the leading sub-trees of the triangular inverse run inside the potrf stage (early phases on a
                # low-priority stream), so those two stages are only meaningful together
                "stage_tflops": {"potrf+trtri": (2 * float(N_TRAIN) ** 3 / 3)
                                 / ((stages["potrf"] + stages["trtri"]) / args.steps * 1e-3) * 1e-12,
                                 "syrk": (float(N_TRAIN) ** 3 / 3) / (stages["syrk"] / args.steps * 1e-3) * 1e-12},
                "stage_ms": {s: stages[s] / args.steps for s in stages}}
    # the element-wise (HBM-side) stages of the same evaluation against the measured copy bandwidth: algorithmic bytes
    # per DESIGN.md section 3.3 -- the lower 128-blocks of K plus the two radial-factor stash arrays written by the
    # assembly; K^-1, the stash and U (alpha = U z) read by the gradient stage
    hbm_peak, hbm_src = measured_hbm_peak()
    nb = -(-N_TRAIN // 128)
    tri = nb * (nb + 1) / 2 * 128 * 128 * 8.0                       # bytes of one lower block triangle
    asm_bytes, grad_bytes = 3 * tri, 4 * tri                         # Matern-5/2: K + g + k_pure;  K^-1 + g + k_pure + U
    roofline["elementwise"] = {
        "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
        "assemble": {"ms": stages["assemble"] / args.steps, "algorithmic_bytes": asm_bytes,
                     "achieved": asm_bytes / (stages["assemble"] / args.steps * 1e-3) * 1e-9,
                     "frac": asm_bytes / (stages["assemble"] / args.steps * 1e-3) * 1e-9 / hbm_peak,
                     "traffic": 3.204e9, "kernel": "gpb::assemble_sym_kernel (FP64 exp/sqrt-bound, profiles/r02_ncu_assemble_sym_stash.txt)"},
        "gradient_stage": {"ms": stages["grad"] / args.steps, "algorithmic_bytes": grad_bytes,
                           "achieved": grad_bytes / (stages["grad"] / args.steps * 1e-3) * 1e-9,
                           "frac": grad_bytes / (stages["grad"] / args.steps * 1e-3) * 1e-9 / hbm_peak,
                           "traffic": 3.270e9 + 1.075e9,
                           "kernel": "gpb::grad_contract_stash_kernel (bulk-TMA ring) + trigemv_kernel "
                                     "(profiles/r02_ncu_grad_contract_stash.txt, r02_ncu_trigemv.txt)"}}
    if pred is not None:
        pred["roofline"] = {"bound": "tensor", "achieved": pred["step_tflops_per_gpu"], "peak": peak,
                            "unit": "TFLOP/s", "frac": pred["step_tflops_per_gpu"] / peak,
                            "traffic": NCU_DRAM_BYTES_COLSQ_LAUNCH,
                            "traffic_note": NCU_DRAM_BYTES_COLSQ_NOTE,
                            "achieved_is": "N^2 flop per candidate x candidates of the step / the WHOLE timed step "
                                           "(cross kernel, variance GEMM, finish, chunk loop), per GPU",
                            "kernel": "gpb::dgemm_tasks_kernel<128,64,32,32,4,EPI_COLSQ,2> (||L^-1 k*||^2 tiles)"}
    # CPU baselines: rank 0 of the single-GPU run only (the scaling runs would just repeat them)
    want_cpu = not args.skip_cpu and world == 1
    cpu = cpu_nll_grad_sample(args.cpu_sample_n) if want_cpu else None
    if pred is not None and want_cpu:
        pred["cpu_baseline"] = cpu_predict_sample()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(), "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "best_restart": {"index": int(best_row[0]), "nll": float(best_row[1])},
            "al_predict": pred, "c4_multistart": c4, "small": small, "assembly": asm}
    print(json.dumps(line), flush=True)
    be.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-n", type=int, default=4096, help="size of the bounded CPU sample evaluation")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-predict", action="store_true")
    ap.add_argument("--skip-assembly", action="store_true")
    ap.add_argument("--predict-candidates", type=int, default=0,
                    help="total candidates of the predict leg; 0 = the 2^22 of configs[3] (default)")
    ap.add_argument("--predict-full", action="store_true", help="(default now) run all 2^22 candidates of configs[3]")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--c4-restarts", type=int, default=64)
    ap.add_argument("--c4-streams", type=int, default=2, help="concurrent restarts per GPU (own handle + stream each)")
    ap.add_argument("--skip-small", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not os.path.exists(os.path.join(ROOT, "profit_b200", "libgpb.so")):
            import __graft_entry__ as entry

            entry.build()
        run_gpu(args)


if __name__ == "__main__":
    main()
