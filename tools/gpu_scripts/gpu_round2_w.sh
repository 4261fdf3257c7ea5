#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_thread_unstructured" 2>&1 | tail -5
python tools/bench_config2.py 64 10 > gpurun_out/s3o_config2.json 2> gpurun_out/s3o_err.txt
cat gpurun_out/s3o_config2.json; tail -3 gpurun_out/s3o_err.txt
