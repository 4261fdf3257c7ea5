#!/usr/bin/env python3
"""Stall-reason mix inside a source region. usage: ncu_region_stalls.py rep file lo hi"""
import csv, subprocess, sys, io
rep, f, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None; inreg = False; acc = {}; ninst = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 50: continue
    if r[0].isdigit(): inreg = (cur == f and lo <= int(r[0]) <= hi); continue
    if not inreg: continue
    try: ninst += int(r[7])
    except ValueError: pass
    for i in range(32, 49):
        try: acc[hdr[i]] = acc.get(hdr[i], 0) + int(r[i])
        except ValueError: pass
tot = sum(acc.values()) or 1
print("warp instructions", ninst, "samples", tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]: print(f"  {v / tot * 100:5.1f}% {k}")
