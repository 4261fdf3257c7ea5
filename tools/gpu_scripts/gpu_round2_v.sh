#!/bin/bash
# row-owner kernel with the block values read one column ahead
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_thread_unstructured or exodus" 2>&1 | tail -5
for cfg in "2 1" "3 1"; do
  set -- $cfg
  python tools/bench_momentum_list.py 128 5 1 nobox 0 $1 $2 > gpurun_out/s3n_urow_$1_$2.json 2> gpurun_out/s3n_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3n_urow_$1_$2.json').read().strip().splitlines()[-1]); print('minb $1 owner $2: urow', round(d['auto']['Melem_per_s'],1), 'Melem/s', round(d['auto']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'], d['split_vs_general_rhs_relerr'])"
done
tail -3 gpurun_out/s3n_err.txt
