#!/usr/bin/env python3
"""Stall samples, executed warp instructions and shared-memory wavefronts per source FUNCTION of an ncu report.
usage: ncu_funcs.py report.ncu-rep nodes src1.cuh [src2.cu ...]"""
import csv, io, subprocess, sys, collections, re, os
rep = sys.argv[1]; nodes = float(sys.argv[2]); srcs = sys.argv[3:]
defs = {}
for s in srcs:
    d = []
    for n, l in enumerate(open(s), 1):
        m = re.match(r"(?:__host__ __device__|__device__|__global__|static|template).*?\b(\w+)\s*\(", l)
        if m and not l.startswith("template <"): d.append((n, m.group(1)))
    defs[os.path.basename(s)] = d
def func(f, line):
    name = f
    for n, fn in defs.get(f, []):
        if n <= line: name = fn
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    d = dict(zip(hdr, r)); k = func(cur, int(r[0]))
    for key, col in (("samples", "# Samples"), ("inst", "Instructions Executed"), ("wf", "L1 Wavefronts Shared"), ("wfx", "L1 Wavefronts Shared Excessive")):
        try: agg[k][key] += int(d.get(col, 0) or 0)
        except ValueError: pass
    for col in hdr:
        if col.startswith("stall_") and "Not Issued" not in col:
            try: agg[k][col] += int(d.get(col, 0) or 0)
            except ValueError: pass
ts = sum(v["samples"] for v in agg.values()) or 1
print(f"{'function':24s} {'samples':>8s} {'inst/node':>9s} {'wf/node':>8s} {'excess':>7s}  top stalls")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:25]:
    st = sorted(((c, n) for n, c in v.items() if n.startswith("stall_")), reverse=True)[:4]
    tot = sum(c for c, _ in st) or 1
    print(f"{k:24s} {100.0 * v['samples'] / ts:7.1f}% {v['inst'] / nodes:9.1f} {v['wf'] / nodes:8.1f} {100.0 * v['wfx'] / max(v['wf'], 1):6.1f}%  " +
          " ".join(f"{n[6:]}:{100.0 * c / max(v['samples'], 1):.0f}%" for c, n in st))
