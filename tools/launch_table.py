#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel / grid."""
import collections
import csv
import re
import sys


def table(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name, grid, t = r[4], r[8], float(r[-1])
        m = re.search(r"dm_px_kernel<(.*?)>\(", name)
        short = m.group(1) if m else re.sub(r"\(.*", "", name)[:48]
        a = agg.setdefault((short, grid), [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    out = []
    for (k, g), v in agg.items():
        out.append(f"{k:44s} {g:18s} n={v[0]:4d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%")
    out.append(f"total {tot / 1e6:.3f} ms over {len(rows)} launches")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        print(table(p))
