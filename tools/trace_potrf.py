"""Occupancy timeline of one NLL+gradient evaluation from the per-CTA trace of the debug build
(nvcc -DGPB_TRACE ... -o profit_b200/libgpb_trace.so): fraction of the 296 CTA slots (2 per SM) in use per time
window, split by kernel class.  Diagnostic only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from profit_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libgpb_trace.so")
from _data import synthetic_problem
from profit_b200 import GPBackend, KIND_RBF

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
win_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
D = 10
be = GPBackend(0)
lib = be._lib
lib.gpb_debug_trace.restype = ctypes.c_int
lib.gpb_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
X, y = synthetic_problem(N, D)
theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
be.set_train(X, y)
for _ in range(2):
    be.nll(KIND_RBF, theta, want_grad=True)
cap = 1 << 20
buf = torch.zeros(4 * cap, dtype=torch.int64, device="cuda:0")
assert lib.gpb_debug_trace(be._h, ctypes.c_void_p(buf.data_ptr()), cap, None) == 0
be.nll(KIND_RBF, theta, want_grad=True)
st = be.stage_ms()
cnt = ctypes.c_uint()
assert lib.gpb_debug_trace(be._h, None, 0, ctypes.byref(cnt)) == 0
n = min(cnt.value, cap)
rec = buf.cpu().numpy().view(np.uint64).reshape(-1, 4)[:n]
t0, t1 = rec[:, 0].astype(np.float64), rec[:, 1].astype(np.float64)
grid = (rec[:, 2] >> np.uint64(32)).astype(np.int64)
tag = rec[:, 3].astype(np.int64)
base = t0.min()
t0, t1 = (t0 - base) * 1e-6, (t1 - base) * 1e-6      # ms
nk, bm, has_beta, has_ct = tag & 0xFFF, (tag >> 16) & 0xFFF, (tag >> 28) & 1, (tag >> 29) & 1
keep = (tag != 0xD1A7) & (tag != 0xD1A8)      # phase stamps of the diag kernel, analysed at the end
rec_all, tag_all = rec, tag
rec, t0, t1, grid, tag, nk, bm, has_beta, has_ct = (a[keep] for a in (rec, t0, t1, grid, tag, nk, bm, has_beta, has_ct))
n = len(rec)
is_diag = tag == 0xD1A6
cls = np.where(is_diag, 0, np.where(has_beta == 1, np.where(nk >= 16, 1, 2), np.where(has_ct == 1, 4, np.where(bm == 64, 3, 5))))
names = ["diag", "update K>=256", "update K=128", "panel solve", "trtri(U,W)", "other beta=0 (T, syrk, early)"]
# slot weight: a 128x128 CTA fills the SM alone; the others share it pairwise; the diag kernel fills it alone
weight = np.where(is_diag | (bm == 128) & (((tag >> 16) & 0xFFF) == 128) & False, 2.0, 1.0)
weight[is_diag] = 2.0
end = t1.max()
print(f"N={N}: {n} CTAs traced, span {end:.2f} ms, stages {{potrf {st['potrf']:.2f}, trtri {st['trtri']:.2f}, syrk {st['syrk']:.2f}}} ms")
for c, nm in enumerate(names):
    m = cls == c
    if m.any():
        print(f"  {nm:32s}: {m.sum():7d} CTAs, mean {1e3 * (t1[m] - t0[m]).mean():7.1f} us, slot-time {(weight[m] * (t1[m] - t0[m])).sum() / 296:8.3f} ms")
edges = np.arange(0.0, end + win_ms, win_ms)
print(f"window[ms]  slots in use (of 296)  by class: " + " | ".join(nm.split()[0] for nm in names))
for a, b in zip(edges[:-1], edges[1:]):
    ov = np.clip(np.minimum(t1, b) - np.maximum(t0, a), 0.0, None) * weight
    tot = ov.sum() / (b - a)
    parts = [ov[cls == c].sum() / (b - a) for c in range(len(names))]
    print(f"{a:7.1f}  {tot:6.1f}  " + " ".join(f"{p:6.1f}" for p in parts))
# diag kernel phases: records 0xD1A6 (start..end), 0xD1A7 (t0 = loads done), 0xD1A8 (t0 = compute done)
ra, rb, rc = rec_all[tag_all == 0xD1A6], rec_all[tag_all == 0xD1A7], rec_all[tag_all == 0xD1A8]
if len(ra) and len(ra) == len(rb) == len(rc):
    o = lambda r: r[np.argsort(r[:, 1])]
    ra, rb, rc = o(ra), o(rb), o(rc)
    load = (rb[:, 0].astype(np.float64) - ra[:, 0].astype(np.float64)) * 1e-3
    comp = (rc[:, 0].astype(np.float64) - rb[:, 0].astype(np.float64)) * 1e-3
    store = (ra[:, 1].astype(np.float64) - rc[:, 0].astype(np.float64)) * 1e-3
    print(f"diag kernel phases (us): load {load.mean():.1f} (max {load.max():.1f}), compute {comp.mean():.1f}, "
          f"finalise+store {store.mean():.1f}")
