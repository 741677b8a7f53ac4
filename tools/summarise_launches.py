#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: summarise_launches.py file.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(u, 1.0)
    a = agg[k]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("%-28s %6s %14s %7s %12s" % ("kernel", "n", "total_us", "share", "max_us"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-28s %6d %14.1f %6.1f%% %12.1f" % (k, a[0], a[1], 100 * a[1] / tot, a[2]))
print("%-28s %6d %14.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
