#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel, launches, total ms, share.
usage: ncu_launch_list.py launches.csv "header text" > profiles/<name>_launch_list.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# kernel, launches, total ms, share, avg ms per launch")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {n:5d} {ms:10.3f} ms {100 * ms / tot:5.1f}%  {ms / n:8.3f}")
