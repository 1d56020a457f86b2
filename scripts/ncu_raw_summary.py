#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into a markdown table of the metrics the roofline uses.
usage: scripts/ncu_raw_summary.py raw.csv [raw2.csv ...] > profiles/xxx.md"""
import csv
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sectors_op_read.sum", "L2 rd sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 MMA % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
]


def fmt(v, u):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return "%.2f MB" % (f / 1e6)
    if u in ("ns", "us", "ms"):
        f *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
        return "%.1f us" % f
    if u == "%":
        return "%.1f" % f
    return ("%d" % f) if f == int(f) else "%.2f" % f


for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    if len(rows) < 3:
        print("\n(%s: no kernels captured)\n" % path)
        continue
    hdr, units = rows[0], rows[1]
    cols = []
    seen = set()
    for key, label in COLS:
        if key in hdr and label not in seen:
            cols.append((hdr.index(key), label))
            seen.add(label)
    ki = hdr.index("Kernel Name")
    print("\n### %s\n" % path.split("/")[-1])
    print("| # | kernel | " + " | ".join(l for _, l in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(rows[2:]):
        name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
        print("| %d | `%s` | " % (n, name) + " | ".join(fmt(r[i], units[i]) for i, _ in cols) + " |")
