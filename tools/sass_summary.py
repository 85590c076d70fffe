"""SASS opcode summary of profit_b200/libgpb.so for profiles/ (run here: cuobjdump cross-reads sm_100a cubins).
    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "profit_b200", "libgpb.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEEP = ("DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UTMA", "UBLKCP", "SYNCS", "LDS", "STS", "RED", "ATOM", "SHFL", "BAR",
        "HMMA", "UTC", "LDGSTS", "ACQBULK")
WIDTH = ("DMMA", "UTMA", "UBLKCP", "SYNCS", "LDS", "STS")
cur, per = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        parts = m.group(1).split(".")
        key = parts[0] + ("." + parts[1] if parts[0].startswith(WIDTH) and len(parts) > 1 else "")
        per[cur][key] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS opcode summary of profit_b200/libgpb.so (sm_100a): cuobjdump -sass, instructions per mnemonic")
print("# build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 (see __graft_entry__.py)")
print("# DMMA.8x8x4 = the FP64 tensor instruction (mma.sync.m8n8k4.f64); UTMALDG.2D = cp.async.bulk.tensor.2d loads;")
print("# UBLKCP = cp.async.bulk (non-tensor bulk copies / stores); SYNCS.* = mbarrier arrive / expect-tx / try_wait.")
print("# No UTC*MMA: tcgen05 has no f64 kind, so the FP64 contraction cannot use TMEM.\n")
print("whole library (%d kernels, %d instructions): " % (len(per), sum(tot.values()))
      + ", ".join(f"{k}={v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if k.startswith(KEEP)) + "\n")
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
for (f, c), name in sorted(zip(per.items(), names), key=lambda t: -sum(t[0][1].values())):
    sel = {k: v for k, v in c.items() if k.startswith(KEEP)}
    name = re.sub(r"\(.*", "", name)
    print(f"{name}\n    total={sum(c.values())} " + " ".join(f"{k}={v}" for k, v in sorted(sel.items(), key=lambda kv: -kv[1])))
