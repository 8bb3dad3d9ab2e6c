"""Summarise an `ncu --page source --csv` dump: executed instructions per
opcode and the hottest SASS lines by stall samples (developer tool).

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_source_summary.py src.csv [kernel-index]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# the file holds one block per profiled launch: "Kernel Name" row, header row, data
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {k: i for i, k in enumerate(b["hdr"])}
ops = collections.Counter()
samples = collections.Counter()
total = 0
lines = []
for r in b["data"]:
    src = r[h["Source"]].strip()
    toks = [t for t in src.split() if not t.startswith("@")]
    if not toks:
        continue
    op = toks[0].split(".")[0]
    n = int(float(r[h["Instructions Executed"]] or 0))
    s = int(float(r[h["# Samples"]] or 0))
    ops[op] += n
    samples[op] += s
    total += n
    lines.append((s, n, src))
print(b["name"])
print("warp instructions executed:", total)
for op, n in ops.most_common(30):
    print("  %-8s %10d  %5.1f%%   samples %d" % (op, n, 100.0 * n / total, samples[op]))
stall_cols = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter()
for r in b["data"]:
    for k in stall_cols:
        tot[k] += int(float(r[h[k]] or 0))
print("stall samples:", ", ".join("%s %d" % kv for kv in tot.most_common(10)))
print("hottest lines:")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print("  %6d %9d  %s" % (s, n, src))
