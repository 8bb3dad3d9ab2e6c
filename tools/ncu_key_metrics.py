"""Cuts the metrics DESIGN.md quotes out of an `ncu --page raw --csv` dump, one
row per profiled launch (developer tool; output is what profiles/ keeps).

    python tools/ncu_key_metrics.py gpurun_out/r2b_kernels_raw.csv [labels,comma,separated]
"""
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.per_cycle_active", "issue"),
    ("smsp__warps_eligible.avg.per_cycle_active", "elig"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("smsp__inst_executed.sum", "inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankc"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("local_load", None),
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
spill = [k for k in hdr if k.startswith("smsp__inst_executed_op_local")]
print("launch | case | kernel | " + " | ".join(n for _, n in KEYS if n) + " | local_ld/st")
for li, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void (mt::)?", "", name)
    vals = []
    for k, n in KEYS:
        if n is None:
            continue
        if k in col:
            u = units[col[k]]
            v = r[col[k]].replace(",", "")
            try:
                f = float(v)
                if u in ("ns", "nsecond"):
                    v = "%.2fus" % (f / 1e3)
                elif u in ("us", "usecond"):
                    v = "%.2fus" % f
                elif u in ("ms", "msecond"):
                    v = "%.2fus" % (f * 1e3)
                elif u == "byte":
                    v = "%.2fMB" % (f / 1e6)
                elif u == "Kbyte":
                    v = "%.2fMB" % (f / 1e3)
                elif u == "Mbyte":
                    v = "%.2fMB" % f
                elif u == "Gbyte":
                    v = "%.2fMB" % (f * 1e3)
                elif abs(f) >= 1000:
                    v = "%d" % f
                else:
                    v = "%.2f" % f
            except ValueError:
                pass
            vals.append(v)
        else:
            vals.append("-")
    sp = "/".join(r[col[k]] for k in spill) if spill else "-"
    lab = labels[li] if li < len(labels) else ""
    print("%d | %s | %s | %s | %s" % (li, lab, name[:70], " | ".join(vals), sp))
