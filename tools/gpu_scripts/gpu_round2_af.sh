#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex27.py -q -x 2>&1 | tail -15
timeout 300 python tools/bench_hex27.py 16 5 2 > gpurun_out/s4d_hex27_16.json 2> gpurun_out/s4d_err.txt; tail -c 1500 gpurun_out/s4d_hex27_16.json; tail -5 gpurun_out/s4d_err.txt
timeout 600 python tools/bench_hex27.py 48 10 3 > gpurun_out/s4d_hex27_48.json 2>> gpurun_out/s4d_err.txt; tail -c 1800 gpurun_out/s4d_hex27_48.json; tail -5 gpurun_out/s4d_err.txt
