#!/usr/bin/env python3
"""Summarise an ncu report (read on the GPU-less box): headline metrics, stall reasons, per-region
instruction mix and the hottest source lines.  usage: ncu_summary.py report.ncu-rep [nodes]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
nodes = float(sys.argv[2]) if len(sys.argv) > 2 else None

def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout

rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
print("== headline")
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {vals[i]:>18s} {units[i]}")
st = []
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try: st.append((float(vals[i].replace(",", "")), h.split("stalled_")[1]))
        except ValueError: pass
tot = sum(v for v, _ in st) or 1
print("== stall reasons (all samples)")
for v, h in sorted(st, reverse=True)[:10]:
    print(f"  {v / tot * 100:5.1f}% {h}")

rows = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
cur = None; line = None; hdrseen = False
lines = {}; mix = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdrseen = True; continue
    if not hdrseen or len(r) < 8: continue
    if r[0].isdigit():
        line = int(r[0])
        try: lines[(cur, line, r[1].strip()[:100])] = (int(r[4]), int(r[7]))
        except ValueError: pass
        continue
    if r[2] in ("-", "..."): continue
    sass = r[3].strip().split()
    if not sass: continue
    op = sass[1] if sass[0].startswith("@") and len(sass) > 1 else sass[0]
    try: ie = int(r[7]); s = int(r[4])
    except ValueError: continue
    d = mix.setdefault((cur, line), {"all": 0, "fp64": 0, "smem": 0, "glob": 0, "local": 0, "samples": 0})
    d["all"] += ie; d["samples"] += s
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU")): d["fp64"] += ie
    if op.startswith(("LDS", "STS")): d["smem"] += ie
    if op.startswith(("LDG", "STG")): d["glob"] += ie
    if op.startswith(("LDL", "STL")): d["local"] += ie
tot = sum(v[0] for v in lines.values()) or 1
print("== hottest source lines (stall samples, warp instructions)")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"  {v[0] / tot * 100:5.1f}% inst={v[1]:>11d} {k[0]}:{k[1]} {k[2]}")
tot_all = {k: sum(d[k] for d in mix.values()) for k in ("all", "fp64", "smem", "glob", "local")}
print("== totals (warp instructions)" + (" per node" if nodes else ""))
for k, v in tot_all.items():
    print(f"  {k:6s} {v / nodes if nodes else v:12.1f}")
# per file/function-ish region: bucket by file and 20-line windows is too fine; print per-file totals
byfile = {}
for (f, l), d in mix.items():
    b = byfile.setdefault(f, {"all": 0, "fp64": 0, "smem": 0, "glob": 0, "local": 0, "samples": 0})
    for k in b: b[k] += d[k]
ts = sum(b["samples"] for b in byfile.values()) or 1
for f, b in byfile.items():
    sc = (1.0 / nodes) if nodes else 1.0
    print(f"  {f:20s} all={b['all'] * sc:10.1f} fp64={b['fp64'] * sc:10.1f} smem={b['smem'] * sc:9.1f} glob={b['glob'] * sc:8.1f} local={b['local'] * sc:8.1f} samples={b['samples'] / ts * 100:5.1f}%")
