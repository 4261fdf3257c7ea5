#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_hex27.py -q -m gpu -k "vs_oracle and 3-2-2 or partitioned" > gpurun_out/s4k_memcheck.txt 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/s4k_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_hex27.py -q -m gpu -k "deterministic" > gpurun_out/s4k_racecheck.txt 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/s4k_racecheck.txt
