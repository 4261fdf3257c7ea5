#!/bin/bash
# general kernel: coloured scatter by reductions vs load/add/store (128^3 momentum list; config-4 scalar lists), determinism tests, smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for sc in 0 1; do
  python tools/bench_momentum_list.py 128 3 $sc > gpurun_out/s3g_general_scatter$sc.json 2> gpurun_out/s3g_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3g_general_scatter$sc.json').read().strip().splitlines()[-1]); print('scatter $sc general', round(d['general']['Melem_per_s'],1), 'Melem/s', round(d['general']['ms'],2), 'ms; relerr', d['split_vs_general_values_relerr'])"
done
timeout 600 python tools/bench_config4.py --nx 128 --layers 128 --steps 3 --mode general --no-momentum --no-oracle-check > gpurun_out/s3g_config4_general_128.json 2>> gpurun_out/s3g_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s3g_config4_general_128.json').read().strip().splitlines()[-1]); print('config4 general tke', d['tke']['kernel_melem_s_per_gpu'], 'sdr', d['sdr']['kernel_melem_s_per_gpu'])"
timeout 900 python -m pytest tests -q -m gpu -k "unstructured or exodus or determin or system_parity or periodic or face" 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
