#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex27.py -q -x 2>&1 | tail -15
timeout 600 python tools/bench_hex27.py 48 10 3 > gpurun_out/s4o_hex27_1gpu.json 2> gpurun_out/s4o_err.txt
tail -c 300 gpurun_out/s4o_err.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/s4o_hex27_1gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'asm', round(d['assemble_ms'],2), d.get('momentum_list'), d['checks'])
PY
