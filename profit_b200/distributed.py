"""Multi-GPU sharding of the two paths that partition naturally.

One process per GPU (torchrun); ``torch.distributed`` is only used for tiny result gathers:

* candidate prediction -- contiguous shards of ceil(M/G) candidates per rank, X/y/theta replicated, every rank
  factorises on its own; one all_gather of (best value, global index) = 16 B per rank, then a local arg-max
  with lowest-index tie-break (np.argmax semantics, aquisition_functions.py:68).
* multi-start hyper-parameter optimisation -- the restarts are handed out one at a time from a counter in the
  process group's key-value store (BFGS paths differ in length by an order of magnitude once a start wanders into the
  reference's eigen-decomposition exit, so a static r -> rank r mod G map leaves most GPUs idle behind the unlucky one;
  it remains the fallback without a store); one all_gather of (nll, theta[P]) per restart, arg-min with tie-break on
  the restart index.  The reference has a single start (gp_functions.py:59-66); per restart the optimiser is exactly
  ``gp_functions.optimize``, and a restart gives the same result on whichever rank it runs.

A single NLL+gradient evaluation does not shard ("replicas only").  No data-path collective exists.
"""
import numpy as np


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_bounds(M, world_size, rank):
    """Contiguous shard [lo, hi) of ceil(M/G) items for this rank (the last shards may be short or empty)."""
    per = -(-int(M) // int(world_size))
    lo = min(int(M), rank * per)
    hi = min(int(M), lo + per)
    return lo, hi


def _all_gather_rows(row):
    """all_gather a small float64 vector from every rank -> (G, len) numpy array (same on all ranks)."""
    d = _dist()
    row = np.asarray(row, dtype=np.float64)
    if d is None:
        return row[None, :]
    import torch

    if d.get_backend() == "nccl":
        # the rank's own GPU (the one GPBackend picks: LOCAL_RANK under torchrun), whatever torch's current device is
        from .backend import _default_device

        dev = torch.device("cuda", _default_device())
    else:
        dev = torch.device("cpu")
    mine = torch.from_numpy(row.copy()).to(dev)
    out = [torch.empty_like(mine) for _ in range(d.get_world_size())]
    d.all_gather(out, mine)
    return torch.stack(out).cpu().numpy()


def gather_argmax(local_value, local_global_index):
    """Global (index, value) of the maximum over ranks with np.argmax's order: the first NaN wins, otherwise the
    largest value, ties -> lowest global index; ranks with an empty shard pass index -1."""
    rows = _all_gather_rows([local_value, float(local_global_index)])
    best_val, best_idx = -np.inf, -1
    for val, idx in rows:
        idx = int(idx)
        if idx < 0:
            continue
        if best_idx < 0:
            better = True
        elif np.isnan(val) != np.isnan(best_val):
            better = bool(np.isnan(val))
        elif not np.isnan(val) and val != best_val:
            better = val > best_val
        else:
            better = idx < best_idx
        if better:
            best_val, best_idx = val, idx
    return best_idx, best_val


def gather_minmax(local_stats):
    """Combine per-shard [min, max] rows (ncol, 2) into the global ones; empty shards pass [+inf, -inf]."""
    local_stats = np.atleast_2d(np.asarray(local_stats, dtype=np.float64))
    rows = _all_gather_rows(local_stats.ravel()).reshape(-1, local_stats.shape[0], 2)
    return np.stack([rows[:, :, 0].min(axis=0), rows[:, :, 1].max(axis=0)], axis=1)


def gather_concat(local, total):
    """Concatenate the contiguous shards of a 1-D array (shard_bounds layout) -> full array of length ``total``."""
    rank, ws = world()
    if ws == 1:
        return np.asarray(local, dtype=np.float64)
    per = -(-int(total) // ws)
    padded = np.zeros(per)
    padded[: len(local)] = local
    return _all_gather_rows(padded).ravel()[:total]        # shard r starts at r * per: rank-ordered rows line up


_QUEUE_CALLS = [0]


def restart_queue(total):
    """Thread-safe ``next_index()`` over range(total), shared by every worker of every rank: a counter in the default
    process group's store (``store.add`` is atomic; the key is unique per call because all ranks call this function the
    same number of times).  Without a process group, or if the store cannot be reached, rank r walks r, r + G, ...
    Returns (next_index, mode) with mode "store" or "static"."""
    import itertools
    import threading

    rank, ws = world()
    lock = threading.Lock()
    _QUEUE_CALLS[0] += 1
    store = None
    if ws > 1:
        try:
            from torch.distributed import distributed_c10d

            store = distributed_c10d._get_default_store()
            key = f"profit_b200/restart_queue/{_QUEUE_CALLS[0]}"
            store.add(key, 0)
        except Exception:                      # noqa: BLE001 -- any failure falls back to the static map
            store = None
    if store is not None:
        def next_index():
            with lock:
                i = int(store.add(key, 1)) - 1
            return i if i < total else None

        return next_index, "store"
    static = iter(range(rank, int(total), ws))

    def next_index():
        with lock:
            return next(static, None)

    return next_index, "static"


def gather_argmin_rows(local_rows):
    """local_rows: (k, 2+P) rows [restart_index, nll, theta...]; returns the row with the smallest nll over all
    ranks (ties -> smallest restart index).  Ranks may hold different numbers of rows."""
    local_rows = np.atleast_2d(np.asarray(local_rows, dtype=np.float64))
    d = _dist()
    width = local_rows.shape[1]
    counts = _all_gather_rows([float(local_rows.shape[0])])[:, 0].astype(int)
    kmax = int(counts.max())
    padded = np.full((kmax, width), np.inf)
    padded[:, 0] = -1
    padded[: local_rows.shape[0]] = local_rows
    allrows = _all_gather_rows(padded.ravel()).reshape(-1, width) if d else padded
    allrows = allrows[allrows[:, 0] >= 0]
    order = np.lexsort((allrows[:, 0], allrows[:, 1]))
    return allrows[order[0]], allrows[np.argsort(allrows[:, 0])]


def predict_sharded(backend, Xpred, add_data_variance=True):
    """Predict this rank's contiguous shard of the candidate set.  Returns (lo, hi, mean, var)."""
    rank, ws = world()
    lo, hi = shard_bounds(len(Xpred), ws, rank)
    if hi > lo:
        mean, var = backend.predict(np.ascontiguousarray(Xpred[lo:hi], dtype=np.float64),
                                    add_data_variance=1 if add_data_variance else 0)
    else:
        mean, var = np.empty((0, backend.n_out)), np.empty(0)
    return lo, hi, mean, var


def acquisition_argmax(backend, Xpred, add_data_variance=True, mask=None):
    """argmax_c var(c) * mask(c) over the whole candidate set, sharded over ranks
    (SimpleExploration: aquisition_functions.py:63-74,107-119).  Returns (global index, value)."""
    lo, hi, _, var = predict_sharded(backend, Xpred, add_data_variance)
    if hi > lo:
        score = var if mask is None else var * np.asarray(mask, dtype=np.float64)[lo:hi]
        li, lv = backend.argmax(np.ascontiguousarray(score))
        return gather_argmax(lv, lo + li)
    return gather_argmax(-np.inf, -1)


def multistart_optimize(xtrain, ytrain, starts, kernel, fixed_sigma_n=False, eval_gradient=True, backend=None,
                        streams=1):
    """Run ``gp_functions.optimize`` from every start in ``starts`` (R, P), restarts handed to the ranks' workers one
    at a time (``restart_queue``).

    ``streams`` > 1 runs that many restarts of this rank concurrently, each on its own device handle and stream
    (host threads; the C calls release the GIL).  One evaluation cannot fill the GPU while its Cholesky walks the
    panel chain, so independent restarts overlap: measured on B200 +4 % at N=8192 on plain evaluations
    (tools/sweep_concurrent.sh; +13 % before the chain was shortened) and 59 -> 47 s on 12 BFGS restarts, where the second
    stream also hides the latency-bound eig exits of an unlucky restart.

    Returns (best_theta, best_nll, table) where table has one row per restart: [restart, nll, theta...].
    The number of objective evaluations this rank spent (the length of the BFGS paths, which decides the wall time as
    much as the speed of one evaluation does) is left in ``multistart_optimize.last_evaluations``.
    """
    import threading

    from . import gp_functions as gpf
    from .backend import GPBackend, shared_backend

    rank, ws = world()
    starts = np.atleast_2d(np.asarray(starts, dtype=np.float64))
    if starts.size == 0:
        raise ValueError("multistart_optimize needs at least one start point")
    streams = max(1, min(int(streams), -(-starts.shape[0] // ws)))
    first = backend or shared_backend()
    backends = [first] + [GPBackend(first.device) for _ in range(streams - 1)]
    rows, lock = [], threading.Lock()
    calls0 = [getattr(be, "nll_calls", 0) for be in backends]

    def one(r, be):
        try:
            theta = gpf.optimize(xtrain, ytrain, starts[r], kernel, fixed_sigma_n=fixed_sigma_n,
                                 eval_gradient=eval_gradient, backend=be)
            full = np.append(theta, starts[r][-1]) if fixed_sigma_n else theta
            nll = gpf.negative_log_likelihood_cholesky(full, xtrain, ytrain, kernel, backend=be)
        except np.linalg.LinAlgError:
            theta, nll = starts[r][:-1] if fixed_sigma_n else starts[r], np.inf
        with lock:
            rows.append(np.concatenate([[r, nll], theta]))

    # Every restart is deterministic on whichever handle of whichever rank it runs, so the table does not depend on
    # who took what.
    next_restart, multistart_optimize.last_queue_mode = restart_queue(starts.shape[0])

    def worker(k):
        while True:
            r = next_restart()
            if r is None:
                return
            one(r, backends[k])

    if streams == 1:
        worker(0)
    else:
        threads = [threading.Thread(target=worker, args=(k,)) for k in range(streams)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    multistart_optimize.last_evaluations = sum(getattr(be, "nll_calls", 0) - c for be, c in zip(backends, calls0))
    for be in backends[1:]:
        be.close()
    width = 2 + (starts.shape[1] - (1 if fixed_sigma_n else 0))
    local = np.array(sorted(rows, key=lambda row: row[0])) if rows else np.empty((0, width))
    best, table = gather_argmin_rows(local) if len(local) or ws > 1 else (None, local)
    return best[2:], float(best[1]), table
