#!/usr/bin/env python3
"""Per-source-line warp-instruction counts and stall samples of an ncu report, sorted by
instructions.  usage: ncu_lines.py report.ncu-rep nodes [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nodes = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdrseen = False; lines = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdrseen = True; continue
    if not hdrseen or len(r) < 8: continue
    if r[0].isdigit():
        try: lines[(cur, int(r[0]), r[1].strip()[:90])] = (int(r[4]), int(r[7]))
        except ValueError: pass
ts = sum(v[0] for v in lines.values()) or 1
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / nodes:7.1f} i/node {v[0] / ts * 100:5.1f}% stall  {k[0]}:{k[1]} {k[2]}")
