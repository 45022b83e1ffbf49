#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (argv[1] = csv)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"][:78]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(k, [0, 0.0, []])
    a[0] += 1; a[1] += v; a[2].append(round(v, 2))
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{k:78s} n={a[0]:3d} total={a[1]:9.3f} ms share={a[1] / tot:.3f} last={a[2][-8:]}")
