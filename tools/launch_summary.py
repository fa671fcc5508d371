#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: launch_summary.py launches.csv "<header comment>" > profiles/rNN_launches_summary.csv"""
import csv
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot, cnt = defaultdict(int), defaultdict(int)
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    tot[name] += int(float(r["Metric Value"]))
    cnt[name] += 1
s = sum(tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("kernel,launches,total_ns,avg_ns,share")
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%s,%d,%d,%d,%.3f" % (k, cnt[k], tot[k], tot[k] // cnt[k], tot[k] / s))
