#!/usr/bin/env python3
"""Shared-memory wavefronts (ideal / excessive) and executed instructions per source line of an ncu report.
usage: ncu_smem.py report.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    try:
        k = (cur, int(r[0]))
        agg[k][0] += int(d.get("L1 Wavefronts Shared", 0) or 0)
        agg[k][1] += int(d.get("L1 Wavefronts Shared Excessive", 0) or 0)
        agg[k][2] += int(d.get("Instructions Executed", 0) or 0)
        agg[k][3] = r[1].strip()[:80]
    except ValueError:
        pass
tw = sum(v[0] for v in agg.values()) or 1; te = sum(v[1] for v in agg.values())
print(f"total shared wavefronts {tw}, excessive {te} ({100.0 * te / tw:.1f} %)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * v[0] / tw:5.1f}% wf  excess {100.0 * v[1] / max(v[0], 1):5.1f}%  inst {v[2]:>11d}  {k[0]}:{k[1]} {v[3]}")
