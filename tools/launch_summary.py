#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py <launches.csv>   (prints a table; the raw list is kept beside it)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("snapgpu::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:58]:58s} {a[0]:8d} {a[1]:10.3f} {a[1] / a[0]:9.3f} {a[1] / tot:6.3f}")
print(f"{'all':58s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")
