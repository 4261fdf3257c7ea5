#!/bin/bash
# slist kernel after the two-pass source terms: timing at the config-4 per-GPU size + parity tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config4" 2>&1 | tail -3
timeout 900 python tools/bench_config4.py --nx 512 --layers 64 --steps 10 --no-momentum --no-oracle-check > gpurun_out/s3e_config4_slist.json 2> gpurun_out/s3e_err.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/s3e_config4_slist.json').read().strip().splitlines()[-1])
for k in ('tke','sdr'): print(k, d[k]['assemble_ms'], d[k]['kernel_melem_s_per_gpu'])
PY
