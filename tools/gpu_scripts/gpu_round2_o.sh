#!/bin/bash
# full GPU suite + smoke()
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/s3f_gpu_suite.txt 2>&1
tail -30 gpurun_out/s3f_gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3f_smoke.txt 2>&1
tail -6 gpurun_out/s3f_smoke.txt
