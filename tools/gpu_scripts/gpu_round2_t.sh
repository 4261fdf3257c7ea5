#!/bin/bash
# row-thread unstructured kernel: register bound / occupancy A/B after the load restructuring
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for mb in 2 3 4; do
  python tools/bench_momentum_list.py 128 5 1 nobox 0 $mb > gpurun_out/s3k_urow_minb$mb.json 2> gpurun_out/s3k_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3k_urow_minb$mb.json').read().strip().splitlines()[-1]); print('minb $mb: urow', round(d['auto']['Melem_per_s'],1), 'Melem/s', round(d['auto']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'], d['split_vs_general_rhs_relerr'])"
done
tail -3 gpurun_out/s3k_err.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_thread_unstructured or exodus" 2>&1 | tail -3
