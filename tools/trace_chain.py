"""Panel-chain timeline of one Cholesky factorisation from the per-CTA trace of the debug build
(nvcc -DGPB_TRACE ... -o profit_b200/libgpb_trace.so): for every 128-column block the diagonal-block kernel and the
launches up to the next diagonal block, with the idle gaps between them.  Diagnostic only."""
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

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
want_grad = (sys.argv[2] != "0") if len(sys.argv) > 2 else False
be = GPBackend(0)
lib = be._lib
lib.gpb_debug_trace.restype = ctypes.c_int
lib.gpb_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
X, y = synthetic_problem(N, 8)
theta = np.concatenate([np.linspace(0.6, 1.2, 8), [1.5, 0.3]])
be.set_train(X, y)
for _ in range(3):
    be.nll(KIND_RBF, theta, want_grad=want_grad)
cap = 1 << 20
buf = torch.zeros(4 * cap, dtype=torch.int64, device="cuda:0")
assert lib.gpb_debug_trace(be._h, ctypes.c_void_p(buf.data_ptr()), cap, None) == 0
be.nll(KIND_RBF, theta, want_grad=want_grad)
st = be.stage_ms()
cnt = ctypes.c_uint()
assert lib.gpb_debug_trace(be._h, None, 0, ctypes.byref(cnt)) == 0
rec = buf.cpu().numpy().view(np.uint64).reshape(-1, 4)[:min(cnt.value, cap)]
tag = rec[:, 3].astype(np.int64)
rec = rec[(tag != 0xD1A7) & (tag != 0xD1A8)]
tag = rec[:, 3].astype(np.int64)
t0 = rec[:, 0].astype(np.float64)
t1 = rec[:, 1].astype(np.float64)
base = t0.min()
t0, t1 = (t0 - base) * 1e-3, (t1 - base) * 1e-3          # us
grid = (rec[:, 2] >> np.uint64(32)).astype(np.int64)
is_diag = (tag == 0xD1A6) | (tag == 0xD1A9)
order = np.argsort(t0)
d_idx = [i for i in order if is_diag[i]]
print(f"N={N} want_grad={want_grad}: potrf stage {st['potrf']:.3f} ms, {len(d_idx)} diagonal blocks, {len(rec)} CTAs")
print(" blk  diag[us]  | launches until the next diagonal block: (grid x nk, BM) start-after-diag-end .. end-after-diag-end")
nk, bm, beta = tag & 0xFFF, (tag >> 16) & 0xFFF, (tag >> 28) & 1
rows = []
for b, i in enumerate(d_idx):
    end = t1[i]
    nxt = t0[d_idx[b + 1]] if b + 1 < len(d_idx) else np.inf
    m = (~is_diag) & (t0 >= end - 0.5) & (t0 < nxt)
    # group CTAs by launch signature
    sig = {}
    for j in np.nonzero(m)[0]:
        key = (int(grid[j]), int(nk[j]), int(bm[j]), int(beta[j]))
        s = sig.setdefault(key, [np.inf, 0.0, 0])
        s[0], s[1], s[2] = min(s[0], t0[j]), max(s[1], t1[j]), s[2] + 1
    parts = sorted(sig.items(), key=lambda kv: kv[1][0])
    txt = "  ".join(f"({k[0]}x{k[1]},{k[2]}{'b' if k[3] else ''}) {v[0] - end:5.1f}..{v[1] - end:5.1f}" for k, v in parts[:4])
    rows.append((t1[i] - t0[i], (nxt - end) if np.isfinite(nxt) else np.nan))
    if b < 6 or b % 8 == 0 or b >= len(d_idx) - 3:
        print(f" {b:3d}  {t1[i] - t0[i]:7.1f}   | next diag starts {nxt - end:6.1f} us after this one ends | {txt}")
rows = np.array(rows)
print(f"mean diag {np.nanmean(rows[:, 0]):.1f} us, mean diag-end -> next-diag-start {np.nanmean(rows[:-1, 1]):.1f} us, "
      f"chain per block {np.nanmean(rows[:-1, 0] + rows[:-1, 1]):.1f} us")
