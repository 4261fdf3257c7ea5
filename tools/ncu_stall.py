#!/usr/bin/env python3
"""Top SASS instructions for one stall reason. usage: ncu_stall.py rep reason [top]   (reason: long_sb, short_sb, wait, barrier...)"""
import csv, subprocess, sys, io
rep, reason = sys.argv[1], "stall_" + sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
col = None; cur = None; line = None; items = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": col = r.index(reason); continue
    if col is None or len(r) <= col: continue
    if r[0].isdigit(): line = (cur, int(r[0]), r[1].strip()[:70]); continue
    try: v = int(r[col])
    except ValueError: continue
    if v: items.append((v, line, r[3].strip()[:60]))
tot = sum(v for v, _, _ in items) or 1
for v, l, sass in sorted(items, key=lambda t: -t[0])[:top]:
    print(f"{v / tot * 100:5.1f}%  {l[0]}:{l[1]}  {sass}   | {l[2]}")
