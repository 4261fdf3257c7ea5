#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27 -c 1 -s 9 -o gpurun_out/s4e_hex27 -f \
  python tools/bench_hex27.py 48 1 1 > gpurun_out/s4e_ncu.log 2>&1
tail -3 gpurun_out/s4e_ncu.log | cut -c1-300
ls -la gpurun_out/s4e_hex27.ncu-rep
