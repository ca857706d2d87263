#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv) of the depth-map chain."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"dm_(px|tile)_kernel<", "", row["Kernel Name"])[:70]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (c, t) in agg.items():
    print(f"{c:4d} launches {t / c:9.1f} us each {t:10.1f} us total  {k}")
print(f"total {sum(t for _, t in agg.values()) / 1000:.2f} ms")
