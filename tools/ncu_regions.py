#!/usr/bin/env python3
"""Stall-sample and instruction share by source region. usage: ncu_regions.py rep nodes file:lo-hi=name ..."""
import csv, subprocess, sys, io
rep = sys.argv[1]; nodes = float(sys.argv[2]); regs = []
for a in sys.argv[3:]:
    loc, name = a.split("="); f, r = loc.split(":"); lo, hi = r.split("-"); regs.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdrseen = False; acc = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdrseen = True; continue
    if not hdrseen or len(r) < 8 or not r[0].isdigit(): continue
    try: s, i = int(r[4]), int(r[7])
    except ValueError: continue
    ln = int(r[0]); name = "other:" + cur
    for f, lo, hi, nm in regs:
        if cur == f and lo <= ln <= hi: name = nm; break
    a = acc.setdefault(name, [0, 0]); a[0] += s; a[1] += i
ts = sum(a[0] for a in acc.values()) or 1
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0] / ts * 100:5.1f}% stall {a[1] / nodes:7.1f} i/node  {k}")
