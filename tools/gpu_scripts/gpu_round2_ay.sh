#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27_kernel -c 1 -s 9 -o gpurun_out/s4u_hex27 -f \
  python tools/bench_hex27.py 48 1 1 > gpurun_out/s4u_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27_onedof -c 1 -s 4 -o gpurun_out/s4u_hex27_onedof -f \
  python tools/bench_hex27.py 48 1 1 > gpurun_out/s4u_ncu2.log 2>&1
ls -la gpurun_out/s4u_*.ncu-rep
