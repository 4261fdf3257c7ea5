#!/bin/bash
# full GPU suite + smoke()
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/s2i_gpu_suite.txt 2>&1
tail -30 gpurun_out/s2i_gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2i_smoke.txt 2>&1
cat gpurun_out/s2i_smoke.txt | tail -6
