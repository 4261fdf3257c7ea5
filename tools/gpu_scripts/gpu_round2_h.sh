#!/bin/bash
# GPU suite at HEAD + smoke + general-kernel timing and ncu capture (coloured scatter, momentum list, 96^3)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/s2j_gpu_suite.txt 2>&1
tail -3 gpurun_out/s2j_gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2j_smoke.txt 2>&1; tail -2 gpurun_out/s2j_smoke.txt
python tools/bench_momentum_list.py 128 3 > gpurun_out/s2j_list128.json 2> gpurun_out/s2j_list128.err; cat gpurun_out/s2j_list128.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_general -s 8 -c 1 -o gpurun_out/s2j_general_96 \
  python tools/bench_momentum_list.py 96 1 > gpurun_out/s2j_ncu.log 2>&1
ls -la gpurun_out/s2j_general_96.ncu-rep
