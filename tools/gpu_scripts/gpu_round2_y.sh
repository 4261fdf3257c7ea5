#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for k in continuity_adv scalar_adv_diff; do
  python tools/bench_continuity.py 128 5 $k nobox > gpurun_out/s3q_${k}_nobox_128.json 2> gpurun_out/s3q_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s3q_${k}_nobox_128.json').read().strip().splitlines()[-1]); print('$k', 'general', round(d['general']['Melem_per_s'],1), 'row-owner(mode', d['auto']['kernel_family'], ')', round(d['auto']['Melem_per_s'],1), 'Melem/s; relerr', d['tiled_vs_general_values_relerr'], d['tiled_vs_general_rhs_relerr'])"
done
tail -3 gpurun_out/s3q_err.txt
