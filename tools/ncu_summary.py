"""Summarise ncu outputs into small text files for profiles/: launch list shares and key raw metrics."""
import csv, collections, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    cols = rows[hdr]
    ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        name = r[ki].split('(')[0][:90]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(t for _, t in agg.values())
    out = [f"# {path}: {sum(c for c, _ in agg.values())} launches, {tot/1e3:.2f} ms summed (cold-cache, serialised)"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{t/1e3:10.3f} ms {100*t/tot:6.2f}% {c:5d} launches  avg {t/c:9.1f} us  {k}")
    return "\n".join(out)

def raw(path, pattern):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2:]
    out = [f"# {path}"]
    import re
    for v in vals:
        for n, u, x in zip(names, units, v):
            if re.search(pattern, n):
                out.append(f"{n} [{u}] = {x}")
    return "\n".join(out)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        pat = sys.argv[3] if len(sys.argv) > 3 else (
            r"Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|"
            r"gpu__dram_throughput|sm__throughput.avg.pct|sm__inst_executed_pipe_fp64|sm__pipe_fp64|sm__pipe_tensor|"
            r"sm__warps_active.avg.pct|launch__registers_per_thread|launch__occupancy_limit|smsp__cycles_active.avg|"
            r"sm__cycles_elapsed.avg$|lts__t_bytes.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
            r"smsp__average_warp.*stall|smsp__issue_active.avg.pct|launch__shared_mem_per_block|launch__waves")
        print(raw(sys.argv[2], pat))
