#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean duration and share per kernel.
Usage: python tools/launch_summary.py launches.csv > profiles/rN_launches_summary.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
acc = OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn].split("(")[0]
    for junk in ("void ", "snpg::", "<unnamed>::", "(anonymous namespace)::"):
        name = name.replace(junk, "")
    acc.setdefault(name, []).append(float(r[mv].replace(",", "")))
total = sum(sum(v) for v in acc.values())
print(f"{'kernel':42s} {'launches':>8s} {'mean_ns':>11s} {'share_of_listed':>17s}")
for k, v in acc.items():
    print(f"{k[:40]:42s} {len(v):8d} {sum(v) / len(v):11.1f} {100 * sum(v) / total:16.1f}%")
