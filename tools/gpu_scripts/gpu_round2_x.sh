#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_owner_unstructured_scalar or exodus or periodic" 2>&1 | tail -8
python tools/bench_config2.py 64 10 > gpurun_out/s3p_config2.json 2> gpurun_out/s3p_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s3p_config2.json').read().strip().splitlines()[-1]); p=d['periodic']; print('periodic', p['kernel_family'], round(p['ms'],3), 'ms', round(p['Melem_per_s'],1), 'Melem/s', p['vs_oracle']); print('nonperiodic tiled', round(d['non_periodic_same_size']['Melem_per_s'],1))"
tail -3 gpurun_out/s3p_err.txt
