#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_cont_box -c 1 -s 2 -o gpurun_out/s4g_cont -f \
  python tools/bench_continuity.py 256 2 continuity_adv > gpurun_out/s4g_ncu.log 2>&1
tail -2 gpurun_out/s4g_ncu.log | cut -c1-300
