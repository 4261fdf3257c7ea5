#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_cpp.py -q -m gpu 2>&1 | tail -12
timeout 600 python tools/bench_hex27.py 48 10 3 > gpurun_out/s4q_hex27_1gpu.json 2> gpurun_out/s4q_err.txt
tail -c 400 gpurun_out/s4q_err.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/s4q_hex27_1gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), d.get('one_dof_kernels'), d['checks']['sampled_rows_vs_oracle']['passed'])
PY
