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
    counts = (ctypes.c_longlong * 11)()
    assert lib.gpb_debug_plan(nb, which, None, 0, None, 0, counts) == 0
    nt, nl = int(counts[0]), int(counts[1])
    tasks = np.zeros((max(nt, 1), 8), dtype=np.int32)
    launches = np.zeros((max(nl, 1), 17))
    assert lib.gpb_debug_plan(nb, which, tasks.ctypes.data_as(ctypes.c_void_p), nt, launches.ctypes.data_as(ctypes.c_void_p),
                              nl, counts) == 0
    cuts = [int(c) for c in counts[3:3 + int(counts[2])]]
    get_plan.phase_events = [int(c) for c in counts[7:7 + int(counts[2])]]
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


@pytest.mark.parametrize("nb", [1, 2, 5, 17, 21, 34])
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
    assert streams == {0, 1, 2, 3}
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


# ------------------------------------------------------------------------------------------ static race check
def _footprints(tasks, launches, nb, extra_rows=1):
    """Per launch: boolean read / write masks over 128x128 tiles of the four matrices (L has one extra block row)."""
    shapes = {M_L: (nb + extra_rows, nb), M_U: (nb, nb), M_W: (nb, nb), M_DINV: (nb, 1)}
    offs, tot = {}, 0
    for m, (r, c) in shapes.items():
        offs[m] = (tot, r, c)
        tot += r * c

    def mark(mask, m, r0, r1, c0, c1):
        o, r, c = offs[m]
        view = mask[o:o + r * c].reshape(r, c)
        view[r0 // 128:(r1 + 127) // 128, c0 // 128:(c1 + 127) // 128] = True

    R = np.zeros((len(launches), tot), dtype=bool)
    W = np.zeros((len(launches), tot), dtype=bool)
    for n, L in enumerate(launches):
        typ, k, cfg, mA, mB, mC, mCt = (int(x) for x in L[:7])
        beta, off, nt = L[8], int(L[9]), int(L[10])
        if typ == 0:                                   # diagonal block: L_kk in place, inverse into Dinv
            mark(R[n], M_L, k * 128, k * 128 + 128, k * 128, k * 128 + 128)
            mark(W[n], M_L, k * 128, k * 128 + 128, k * 128, k * 128 + 128)
            mark(W[n], M_DINV, k * 128, k * 128 + 128, 0, 128)
            continue
        if typ == 2:                                   # seed of the U / W diagonal tiles from Dinv, blocks [k, cfg)
            for b in range(k, cfg):
                mark(R[n], M_DINV, b * 128, b * 128 + 128, 0, 128)
                mark(W[n], M_U, b * 128, b * 128 + 128, b * 128, b * 128 + 128)
                mark(W[n], M_W, b * 128, b * 128 + 128, b * 128, b * 128 + 128)
            continue
        bm, bn = TILE[cfg]
        for a_row, a_k, b_row, b_k, nk, c_row, c_col, _aux in tasks[off:off + nt]:
            mark(R[n], mA, a_row, a_row + bm, a_k, a_k + 16 * nk)
            mark(R[n], mB, b_row, b_row + bn, b_k, b_k + 16 * nk)
            mark(W[n], mC, c_row, c_row + bm, c_col, c_col + bn)
            if beta != 0.0:
                mark(R[n], mC, c_row, c_row + bm, c_col, c_col + bn)
            if mCt != M_NONE:
                mark(W[n], mCt, c_col, c_col + bn, c_row, c_row + bm)
    return R, W


def _happens_before(launches):
    """reach[j] = bitmask of launches ordered before j by stream order and event record -> wait pairs (run_plan)."""
    last_on_stream, last_record, reach = {}, {}, []
    for j, L in enumerate(launches):
        stream, waits, rec = int(L[12]), [int(L[13]), int(L[14]), int(L[16])], int(L[15])
        preds = [last_on_stream[stream]] if stream in last_on_stream else []
        preds += [last_record[e] for e in waits if e >= 0 and e in last_record]
        bits = 1 << j
        for p in preds:
            bits |= reach[p]
        reach.append(bits)
        last_on_stream[stream] = j
        if rec >= 0:
            last_record[rec] = j
    return reach


def _assert_race_free(tasks, launches, nb):
    R, W = _footprints(tasks, launches, nb)
    reach = _happens_before(launches)
    RW = R | W
    for j in range(len(launches)):
        if j == 0:
            continue
        conflict = (W[:j] & RW[j]).any(axis=1) | (R[:j] & W[j]).any(axis=1)
        for i in np.nonzero(conflict)[0]:
            assert (reach[j] >> int(i)) & 1, (
                f"launches {int(i)} (stream {int(launches[i][12])}) and {j} (stream {int(launches[j][12])}) touch the same "
                f"tiles without an ordering path")


@pytest.mark.parametrize("nb", [3, 9, 16, 17, 24, 33, 40, 65, 70])
def test_cholesky_plan_orders_every_conflicting_pair_of_launches(lib, nb):
    """Static race check of the multi-stream Cholesky: any two launches whose tile footprints conflict (write/write or
    read/write) must be ordered by stream order or by an event record -> wait pair."""
    tasks, launches, _ = get_plan(lib, nb, 0)
    _assert_race_free(tasks, launches, nb)
    if nb >= 24:
        assert {int(L[12]) for L in launches} == {0, 1, 2, 3}


def test_removing_an_event_is_detected(lib):
    """The checker itself: dropping the wait of one trailing update on its panel's event must raise."""
    tasks, launches, _ = get_plan(lib, 24, 0)
    broken = launches.copy()
    victim = next(n for n, L in enumerate(broken) if L[0] == 1 and int(L[12]) == 0 and int(L[13]) >= 0)
    broken[victim, 13] = -1
    with pytest.raises(AssertionError, match="without an ordering path"):
        _assert_race_free(tasks, broken, 24)


def test_two_panel_lookahead_orders_the_next_column_update(lib):
    """Two panels of look-ahead: the update of panel J+1's columns (panel stream) waits for the small launch that applied
    panel J-1 to the same columns on the look-ahead update stream (T2a), not for the whole trailing update; dropping that wait
    must be detected."""
    nb = 24
    tasks, launches, _ = get_plan(lib, nb, 0)
    t1 = [n for n, L in enumerate(launches) if L[0] == 1 and int(L[12]) == 1 and int(L[13]) >= 0 and L[8] == 1.0]
    assert len(t1) >= 16
    for n in t1[-8:-3]:                                                       # the single-block panels of the tail
        ev = int(launches[n][13])
        rec = next(m for m, L in enumerate(launches) if int(L[15]) == ev)
        assert rec < n and int(launches[rec][12]) == 3                      # recorded on the look-ahead update stream, earlier in issue order
        col = int(tasks[int(launches[n][9])][6]) // 128
        cols = {int(t[6]) // 128 for t in tasks[int(launches[rec][9]):int(launches[rec][9]) + int(launches[rec][10])]}
        assert cols == {col}                                                # ... by a launch that only touches that block column
    broken = launches.copy()
    broken[t1[-5], 13] = -1
    with pytest.raises(AssertionError, match="without an ordering path"):
        _assert_race_free(tasks, broken, nb)


def test_early_inverse_phases_are_ordered_after_the_columns_they_read(lib):
    """The early phases of the triangular inverse run on their own stream while the Cholesky continues: appended to
    the Cholesky's launch list the way assemble_and_factor issues them (wait on the phase event, seed the diagonal
    tiles, run the phase), the combined list must still be race free."""
    nb = 48
    tasks, launches, cuts = get_plan(lib, nb, 0)
    events = list(get_plan.phase_events)               # the events assemble_and_factor makes stream3 wait for
    all_tasks, all_launches = [tasks], [launches]
    off = len(tasks)
    k0 = 0
    for i, cut in enumerate(cuts):
        ev = events[i]
        assert any(int(L[15]) == ev for L in launches)
        seed = np.zeros((1, 17))
        seed[0, :3] = [2, k0, cut]                      # pseudo launch: seed blocks [k0, cut)
        seed[0, 3:7] = M_NONE
        seed[0, 12:17] = [9, ev, -1, -1, -1]
        k0 = cut
        pt, pl, _ = get_plan(lib, nb, 10 + i)
        pl = pl.copy()
        pl[:, 9] += off
        pl[:, 12] = 9                                   # stream3 of the handle (not a plan stream)
        pl[:, 13:17] = -1
        all_tasks.append(pt)
        all_launches += [seed, pl]
        off += len(pt)
    _assert_race_free(np.vstack(all_tasks), np.vstack(all_launches), nb)
