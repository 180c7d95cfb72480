"""Summarise an .ncu-rep (raw + source pages) for the first profiled kernel: python tools/ncu_summary.py rep [warps]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2]
get = lambda k: next((data[i] for i, h in enumerate(hdr) if h == k), None)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg"]
print("kernel:", get("Kernel Name"))
for k in keys:
    i = hdr.index(k) if k in hdr else -1
    if i >= 0:
        print(f"{k:70s} {units[i]:14s} {data[i]}")
for i, h in enumerate(hdr):
    if "average_warps_issue_stalled" in h and float(data[i] or 0) > 0.05:
        print("stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), data[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
isrc, iex, ist = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Warp Stall Sampling (All Samples)")
ops, total, L = collections.Counter(), 0, []
for r in rows[2:]:
    try:
        ex = int(r[iex])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    ops[m.group(2) if m else "?"] += ex
    total += ex
    L.append((int(r[ist] or 0), ex, r[isrc]))
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print("instructions executed total", total, "per unit", total / nw)
print({op: round(c / nw, 1) for op, c in ops.most_common(28)})
tot = sum(x[0] for x in L) or 1
for s, e, t in sorted(L, reverse=True)[:14]:
    print(f"{100*s/tot:5.1f}% ex/unit={e/nw:5.2f} {t[:100]}")
