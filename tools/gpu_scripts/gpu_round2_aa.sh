#!/bin/bash
# quad form of the row-owner momentum kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_thread_unstructured or exodus" 2>&1 | tail -5
for qm in 0 2 3 4; do
  python tools/bench_momentum_list.py 128 5 1 nobox 0 2 1 $qm > gpurun_out/s3s_urow_quad$qm.json 2> gpurun_out/s3s_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3s_urow_quad$qm.json').read().strip().splitlines()[-1]); print('quad minb $qm: urow', round(d['auto']['Melem_per_s'],1), 'Melem/s', round(d['auto']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'], d['split_vs_general_rhs_relerr'])"
done
tail -3 gpurun_out/s3s_err.txt
