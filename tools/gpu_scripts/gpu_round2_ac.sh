#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_urow_owner_kernel -c 1 -s 2 -o gpurun_out/s3y_urow_central -f \
  python tools/bench_momentum_list.py 128 1 1 nobox > gpurun_out/s3y_ncu.log 2>&1
tail -2 gpurun_out/s3y_ncu.log | cut -c1-200
