#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config4 or row_owner_unstructured_onedof or exodus" 2>&1 | tail -6
timeout 600 python tools/bench_config4.py --nx 128 --layers 128 --steps 3 --nobox --no-momentum --no-oracle-check > gpurun_out/s3u_config4_nobox_128.json 2> gpurun_out/s3u_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s3u_config4_nobox_128.json').read().strip().splitlines()[-1]); print('config4 nobox: tke', d['tke']['assemble_mode'], round(d['tke']['kernel_melem_s_per_gpu'],1), 'sdr', round(d['sdr']['kernel_melem_s_per_gpu'],1), 'Melem/s')"
tail -3 gpurun_out/s3u_err.txt
