"""Host-side plan logic without a GPU: the launch / tile-task lists that libgpb would run (``gpb_debug_plan``) are
replayed with numpy, in issue order, on a small SPD matrix and must produce the Cholesky factor with the forward
substitution of y, both triangular inverses and K^-1 -- for block counts that exercise every scheduling variant
(narrowing panels, the trailing updates split over two streams, the early phases of the triangular inverse).
Cross-stream ordering itself is checked on the GPU (tools/stress_streams.py: overlapped == serial, bit for bit)."""
import ctypes

import numpy as np
import pytest

import oracle

M_L, M_U, M_W, M_DINV, M_NONE = 0, 1, 2, 3, 9
TILE = {0: (128, 128), 1: (128, 64), 2: (64, 128), 3: (32, 128), 4: (32, 64)}


@pytest.fixture(scope="module")
def lib(built_lib):
    from profit_b200 import _lib

    return _lib.load()


def get_plan(lib, nb, which):
    counts = (ctypes.c_longlong * 8)()
    assert lib.gpb_debug_plan(nb, which, None, 0, None, 0, counts) == 0
    nt, nl = int(counts[0]), int(counts[1])
    tasks = np.zeros((max(nt, 1), 8), dtype=np.int32)
    launches = np.zeros((max(nl, 1), 16))
    assert lib.gpb_debug_plan(nb, which, tasks.ctypes.data_as(ctypes.c_void_p), nt, launches.ctypes.data_as(ctypes.c_void_p),
                              nl, counts) == 0
    cuts = [int(c) for c in counts[3:3 + int(counts[2])]]
    return tasks[:nt], launches[:nl], cuts


def replay(tasks, launches, mats):
    """Run one plan in issue order.  mats: {id: 2-D array}.  Diagonal-block launches factor L_kk in place and store
    its inverse in Dinv (what potrf_diag4_kernel does)."""
    for L in launches:
        typ, k, cfg, mA, mB, mC, mCt = (int(x) for x in L[:7])
        alpha, beta, off, nt = L[7], L[8], int(L[9]), int(L[10])
        if typ == 0:
            blk = mats[M_L][k * 128:(k + 1) * 128, k * 128:(k + 1) * 128]
            Lkk = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
            blk[...] = Lkk
            mats[M_DINV][k * 128:(k + 1) * 128, :] = np.linalg.inv(Lkk)
            continue
        bm, bn = TILE[cfg]
        A, B, C = mats[mA], mats[mB], mats[mC]
        for a_row, a_k, b_row, b_k, nk, c_row, c_col, _aux in tasks[off:off + nt]:
            prod = A[a_row:a_row + bm, a_k:a_k + 16 * nk] @ B[b_row:b_row + bn, b_k:b_k + 16 * nk].T
            out = alpha * prod + (beta * C[c_row:c_row + bm, c_col:c_col + bn] if beta != 0.0 else 0.0)
            C[c_row:c_row + bm, c_col:c_col + bn] = out
            if mCt != M_NONE:
                mats[mCt][c_col:c_col + bn, c_row:c_row + bm] = out.T


def spd_problem(nb, n_out=2):
    n = nb * 128
    X = np.random.default_rng(nb).random((n, 3))
    K = oracle.rbf(X, X, 0.25, 1.0, 0.6)
    Y = np.random.default_rng(nb + 1).standard_normal((n, n_out))
    return K, Y


def seed_diag(mats, nb):
    for k in range(nb):
        d = mats[M_DINV][k * 128:(k + 1) * 128, :]
        mats[M_W][k * 128:(k + 1) * 128, k * 128:(k + 1) * 128] = d
        mats[M_U][k * 128:(k + 1) * 128, k * 128:(k + 1) * 128] = d.T


@pytest.mark.parametrize("nb", [1, 2, 5, 21, 34])
def test_replayed_plans_factor_and_invert(lib, nb):
    n = nb * 128
    K, Y = spd_problem(nb)
    Lbuf = np.zeros((n + 128, n))
    Lbuf[:n] = np.tril(K)                                        # the assembly writes the lower tiles only
    for k in range(nb):                                         # ... with full diagonal tiles
        Lbuf[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128] = K[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128]
    Lbuf[n:n + Y.shape[1]] = Y.T
    mats = {M_L: Lbuf, M_U: np.zeros((n, n)), M_W: np.zeros((n, n)), M_DINV: np.zeros((n, 128))}
    tasks, launches, _ = get_plan(lib, nb, 0)
    assert [int(L[1]) for L in launches if L[0] == 0] == list(range(nb))     # every diagonal block once, in order
    replay(tasks, launches, mats)
    Lref = np.linalg.cholesky(K)
    Lgot = np.tril(Lbuf[:n])
    assert np.max(np.abs(Lgot - Lref)) <= 1e-11 * np.abs(Lref).max()
    zref = np.linalg.solve(Lref, Y)
    assert np.max(np.abs(Lbuf[n:n + Y.shape[1]].T - zref)) <= 1e-10 * np.abs(zref).max()
    # triangular inverse (full plan) and K^-1
    seed_diag(mats, nb)
    tasks, launches, _ = get_plan(lib, nb, 1)
    replay(tasks, launches, mats)
    Winv = np.linalg.inv(Lref)
    assert np.max(np.abs(np.tril(mats[M_W]) - Winv)) <= 1e-10 * np.abs(Winv).max()
    assert np.max(np.abs(np.triu(mats[M_U]) - Winv.T)) <= 1e-10 * np.abs(Winv).max()
    assert not np.any(np.triu(mats[M_W], 1)) and not np.any(np.tril(mats[M_U], -1))
    tasks, launches, _ = get_plan(lib, nb, 2)
    replay(tasks, launches, mats)
    Kinv = np.linalg.inv(K)
    assert np.max(np.abs(np.tril(Lbuf[:n]) - np.tril(Kinv))) <= 1e-9 * np.abs(Kinv).max()


def test_split_trailing_updates_use_both_streams_and_every_tile_once(lib):
    """Nb = 40: the large trailing updates go out as even / odd block columns on streams 0 and 2; over the whole plan
    every lower tile (i, j) receives exactly 8 j k-slabs of updates and one panel solve."""
    nb = 40
    tasks, launches, _ = get_plan(lib, nb, 0)
    upd = np.zeros((nb + 1, nb), dtype=np.int64)
    solves = np.zeros((nb + 1, nb), dtype=np.int64)
    streams = set()
    for L in launches:
        if L[0] == 0:
            continue
        cfg, mB, beta, off, nt, stream = int(L[2]), int(L[4]), L[8], int(L[9]), int(L[10]), int(L[12])
        bm, bn = TILE[cfg]
        streams.add(stream)
        for a_row, a_k, b_row, b_k, nk, c_row, c_col, _aux in tasks[off:off + nt]:
            frac = (bm * bn) / (128.0 * 128.0)            # a 128x128 tile is covered by several finer tasks
            i, j = c_row // 128, c_col // 128
            if mB == M_DINV:
                solves[i, j] += round(frac * 1024)
            else:
                assert beta == 1.0
                upd[i, j] += round(frac * nk * 1024)
    assert streams == {0, 1, 2}
    for j in range(nb):
        for i in range(j, nb + 1):
            assert upd[i, j] == 8 * j * 1024, (i, j, upd[i, j])
            assert solves[i, j] == (1024 if i > j else 0), (i, j)
    split = [L for L in launches if L[0] == 1 and int(L[12]) == 2]
    assert split and all(int(t[6]) // 128 % 2 == 1 for L in split for t in tasks[int(L[9]):int(L[9]) + int(L[10])])


def test_early_phases_partition_the_triangular_inverse(lib):
    """Nb = 48 (the smallest size with early phases): phases + late part replayed in order give the same U and W as
    the single plan, phase i only touches blocks left of its cut, and its cut is reached by the Cholesky first."""
    nb = 48
    n = nb * 128
    K, _ = spd_problem(nb, 1)
    Lref = np.linalg.cholesky(K)
    Dinv = np.zeros((n, 128))
    for k in range(nb):
        Dinv[k * 128:(k + 1) * 128] = np.linalg.inv(Lref[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128])
    _, _, cuts = get_plan(lib, nb, 0)
    assert cuts == [12, 24, 36, 42]
    mats = {M_L: np.vstack([Lref.copy(), np.zeros((128, n))]), M_U: np.zeros((n, n)), M_W: np.zeros((n, n)), M_DINV: Dinv}
    seed_diag(mats, nb)
    for i, cut in enumerate(cuts):
        tasks, launches, _ = get_plan(lib, nb, 10 + i)
        assert len(tasks) and tasks[:, 5].max() < cut * 128 and tasks[:, 6].max() < cut * 128
        assert all(int(L[2]) == 3 for L in launches)           # short 32x128 tiles
        replay(tasks, launches, mats)
    tasks, launches, _ = get_plan(lib, nb, 20)
    replay(tasks, launches, mats)
    Winv = np.linalg.inv(Lref)
    assert np.max(np.abs(np.tril(mats[M_W]) - Winv)) <= 1e-10 * np.abs(Winv).max()
    assert np.max(np.abs(np.triu(mats[M_U]) - Winv.T)) <= 1e-10 * np.abs(Winv).max()
