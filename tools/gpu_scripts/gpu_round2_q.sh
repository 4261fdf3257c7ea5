#!/bin/bash
# row-thread unstructured momentum kernel: parity tests, 128^3 declared as a general mesh (urow vs general kernel)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_thread_unstructured or exodus or determin" 2>&1 | tail -12
python tools/bench_momentum_list.py 128 5 1 nobox > gpurun_out/s3h_urow_128.json 2> gpurun_out/s3h_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s3h_urow_128.json').read().strip().splitlines()[-1]); print('general', round(d['general']['Melem_per_s'],1), 'auto(mode', d['auto']['kernel_family'], ')', round(d['auto']['Melem_per_s'],1), 'Melem/s', round(d['auto']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'], d['split_vs_general_rhs_relerr'])"
tail -3 gpurun_out/s3h_err.txt
