"""Attribute executed SASS instructions of an ncu capture to CUDA source
lines (developer tool).

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libmtraffic.so && nvdisasm -g *.cubin > all.dis
    python tools/ncu_line_attribution.py src.csv all.dis <mangled kernel name>
"""
import collections
import csv
import re
import sys

src_csv, dis, kernel = sys.argv[1:4]
# offset -> (file, line) from nvdisasm
line_of, cur, on = {}, None, False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text." + kernel + ":"):
        on = True
        continue
    if on and ln.startswith("//-----") or (on and ln.startswith(".text.") and kernel not in ln):
        break
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+\S", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = {k: i for i, k in enumerate(rows[start])}
data = []
for r in rows[start + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if r:
        data.append(r)
base = int(data[0][h["Address"]], 16)
per_line = collections.Counter()
fp64_line = collections.Counter()
total = 0
for r in data:
    off = int(r[h["Address"]], 16) - base
    n = int(float(r[h["Instructions Executed"]] or 0))
    where = line_of.get(off, ("?", 0))
    per_line[where] += n
    op = [t for t in r[h["Source"]].split() if not t.startswith("@")][0]
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"):
        fp64_line[where] += n
    total += n
print("total warp instructions", total)
files = {}
for (f, l), n in per_line.most_common(60):
    if f not in files:
        try:
            files[f] = open("mbrl_traffic_b200/csrc/" + f).read().splitlines()
        except OSError:
            files[f] = []
    text = files[f][l - 1].strip()[:90] if 0 < l <= len(files[f]) else ""
    print("%6.2f%%  fp64 %5.1f%%  %-16s %4d  %s" % (
        100.0 * n / total, 100.0 * fp64_line[(f, l)] / max(n, 1), f, l, text))
