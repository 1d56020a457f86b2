#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: scripts/summarize_launches.py launches.csv > profiles/xxx.md"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k[:100], n, t, 100 * t / tot, t / n))
print("\ntotal: %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))
