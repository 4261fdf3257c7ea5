#!/bin/bash
# row-thread unstructured kernel: (chunk, colour) launch groups so that a chunk's rows stay in L2 across its colours
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for ch in 0 262144 65536 32768 16384; do
  python tools/bench_momentum_list.py 128 5 1 nobox $ch > gpurun_out/s3i_urow_chunk$ch.json 2> gpurun_out/s3i_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3i_urow_chunk$ch.json').read().strip().splitlines()[-1]); print('chunk $ch: general', round(d['general']['Melem_per_s'],1), 'urow', round(d['auto']['Melem_per_s'],1), 'Melem/s', round(d['auto']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'], d['split_vs_general_rhs_relerr'])"
done
tail -3 gpurun_out/s3i_err.txt
