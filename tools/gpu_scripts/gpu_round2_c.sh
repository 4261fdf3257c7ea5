#!/bin/bash
# fused list: tests + 1-GPU bench (default) 
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or system_parity or box_kernel" > gpurun_out/s2d_tests.log 2>&1
tail -3 gpurun_out/s2d_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s2d_bench_1gpu.json 2> gpurun_out/s2d_bench_1gpu.err
tail -c 400 gpurun_out/s2d_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2d_bench_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value'],1), 'frac', round(r['frac'],4), 'kernel', r['kernel'], 'ms', round(r['kernel_ms'],2))
print('single', r.get('single_kernel'))
print('cont', r.get('continuity'))
print('m+c', r.get('momentum_plus_continuity'))
print('e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'])
print('checks', d['checks'])
print('cpu', d['cpu_baseline'])
PY
