#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row["Kernel Name"][:70], []).append(float(row["Metric Value"].replace(",", "")))
total = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'n':>4s} {'mean_us':>10s} {'sum_us':>10s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:72s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / total:5.1f}%")
