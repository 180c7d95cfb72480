"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except (ValueError, IndexError):
        continue
    name = r[kn].split("(")[0][:90]
    agg[name][0] += 1
    agg[name][1] += v
total = sum(t for _, t in agg.values())
print(f"{sum(n for n, _ in agg.values())} launches, {total / 1e6:.2f} ms of kernel time (ncu: cold cache, serialised)")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:6d} launches {t / 1e3:11.1f} us total {t / n / 1e3:10.2f} us avg {100 * t / total:5.1f} %  {k}")
