"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>"
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] == "us":
        v *= 1e3
    elif r["Metric Unit"] == "ms":
        v *= 1e6
    rows.append((r["Kernel Name"], r["Grid Size"], v))
tot = sum(v for _, _, v in rows)
agg = collections.OrderedDict()
for name, grid, v in rows:
    short = re.sub(r"\(.*", "", name)
    key = (short, grid)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
print("ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none)" % cmd)
print("per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes")
print("total kernel time %.1f us over %d launches\n" % (tot / 1e3, len(rows)))
for (name, grid), (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5d launches %10.1f us %5.1f%% avg %7.2f us grid %-14s %s"
          % (n, v / 1e3, 100 * v / tot, v / n / 1e3, grid, name[:110]))
