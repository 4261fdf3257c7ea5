#!/bin/bash
# quick A/B of a kernel change: fused-list + box tests, then the default bench without the CPU baseline leg; tag = $1
cd "$(dirname "$0")/../.."
T=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or box_kernel or row_thread or exodus or sst_des or ksgs" > gpurun_out/s2h_${T}_tests.log 2>&1
tail -3 gpurun_out/s2h_${T}_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s2h_${T}_bench.json 2> gpurun_out/s2h_${T}_bench.err
tail -c 300 gpurun_out/s2h_${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s2h_${T}_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value'],1), 'frac', round(r['frac'],4), 'ms', round(r['kernel_ms'],2))
print('single', r['single_kernel']['value'], r['single_kernel']['list_over_single'])
print('m+c', r['momentum_plus_continuity']['frac'], 'e2e', d['e2e']['value'])
print('checks', d['checks'])
PY
