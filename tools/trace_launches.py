"""Per-launch timeline of one evaluation from the per-CTA trace of the debug build
(nvcc -DGPB_TRACE ... -o profit_b200/libgpb_trace.so): for every kernel launch (CTAs grouped by grid size, k length, tile
height and epilogue) the first start, the last end, the mean CTA time and the DMMA rate its CTAs reached.  Diagnostic only.
    python tools/trace_launches.py [N] [want_grad 0/1] [t_end_ms]"""
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
t_from = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
t_to = float(sys.argv[4]) if len(sys.argv) > 4 else 1e9
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
smid = (rec[:, 2] & np.uint64(0xFFFFFFFF)).astype(np.int64)
is_diag = (tag == 0xD1A6) | (tag == 0xD1A9)
nk, bn, bm, beta, ct = tag & 0xFFF, 32 * ((tag >> 12) & 0xF), (tag >> 16) & 0xFFF, (tag >> 28) & 1, (tag >> 29) & 1
print(f"N={N} want_grad={want_grad}: stages {{potrf {st['potrf']:.3f}, trtri {st['trtri']:.3f}, syrk {st['syrk']:.3f}}} ms, {len(rec)} CTAs")
# launches: CTAs of the same signature in start order; a group is complete when it holds `grid` CTAs
launches = []
order = np.argsort(t0)
groups = {}
for j in order:
    key = ("diag", int(tag[j])) if is_diag[j] else (int(grid[j]), int(nk[j]), int(bm[j]), int(beta[j]), int(ct[j]), int(bn[j]))
    g = groups.get(key)
    if g is None or is_diag[j] or len(g["idx"]) >= key[0]:
        g = {"key": key, "idx": [], "last": t0[j]}
        groups[key] = g
        launches.append(g)
    g["idx"].append(j)
    g["last"] = t0[j]
print("   start     end    span  CTAs  meanCTA  SMs  kind                      GF/s per CTA   launch TF/s over its span")
for g in sorted(launches, key=lambda g: t0[g["idx"][0]]):
    ix = np.array(g["idx"])
    a, b = t0[ix].min(), t1[ix].max()
    if b * 1e-3 < t_from or a * 1e-3 > t_to:
        continue
    k = g["key"]
    if k[0] == "diag":
        print(f"{a:8.1f} {b:8.1f} {b - a:7.1f} {len(ix):5d} {np.mean(t1[ix] - t0[ix]):8.1f} {len(set(smid[ix])):4d}  diag/panel tag {k[1]:#x}")
        continue
    bn_ = k[5] if k[5] else 128     # the panel-solve kernel reports no width: 32 x 128 stripes
    flop = 2.0 * k[2] * bn_ * 16 * k[1]
    per_cta = flop / (np.mean(t1[ix] - t0[ix]) * 1e-6) * 1e-9
    whole = flop * len(ix) / ((b - a) * 1e-6) * 1e-12
    kind = f"grid {k[0]:5d} K={16 * k[1]:4d} {k[2]:3d}x{bn_:3d} {'beta ' if k[3] else '     '}{'Ct' if k[4] else '  '}"
    print(f"{a:8.1f} {b:8.1f} {b - a:7.1f} {len(ix):5d} {np.mean(t1[ix] - t0[ix]):8.1f} {len(set(smid[ix])):4d}  {kind} {per_cta:10.1f} {whole:12.2f}")
