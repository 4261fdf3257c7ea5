#!/bin/bash
# full GPU suite + ncu --set full of the fused-list tiled kernel at 256^3 + launch list of the bench command
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/s2g_gpu_suite.txt 2>&1
tail -25 gpurun_out/s2g_gpu_suite.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_box5 -s 2 -c 1 -o gpurun_out/s2g_box5_list_256 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-continuity --no-oracle-check > gpurun_out/s2g_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2g_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-oracle-check > gpurun_out/s2g_launches_bench.log 2>&1
tail -2 gpurun_out/s2g_ncu.log | cut -c1-300
