"""Opcode histogram of one kernel in a built library (developer tool).

    python tools/sass_hist.py <substring of mangled name> [lib]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "mbrl_traffic_b200/libmtraffic.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE,
                     universal_newlines=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(2).strip())
for name, ins in funcs.items():
    if sys.argv[1] in name:
        print(name, len(ins), "instructions")
        ops = collections.Counter()
        for i in ins:
            toks = [t for t in i.split() if not t.startswith("@")]
            ops[toks[0].split(".")[0]] += 1
        print(", ".join("%s %d" % kv for kv in ops.most_common(45)))
        if len(sys.argv) > 3:
            print("\n".join(ins))
