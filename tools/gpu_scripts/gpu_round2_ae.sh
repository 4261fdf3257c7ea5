#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "row_owner or row_thread or exodus or config4_lists_on_unstructured" 2>&1 | tail -4
python tools/bench_momentum_list.py 128 5 1 nobox > gpurun_out/s4a_urow_128.json 2> gpurun_out/s4a_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s4a_urow_128.json').read().strip().splitlines()[-1]); print('momentum list nobox: row-owner', d['auto']['kernel_family'], round(d['auto']['Melem_per_s'],1))"
python tools/bench_config1_list.py 128 5 > gpurun_out/s4a_config1_list_128.json 2>> gpurun_out/s4a_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s4a_config1_list_128.json').read().strip().splitlines()[-1]); print('config1 list: general', round(d['general']['Melem_per_s'],1), 'row-owner', d['auto']['kernel_family'], round(d['auto']['Melem_per_s'],1))"
for k in continuity_adv scalar_adv_diff; do
  python tools/bench_continuity.py 128 5 $k nobox > gpurun_out/s4a_${k}_nobox_128.json 2>> gpurun_out/s4a_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s4a_${k}_nobox_128.json').read().strip().splitlines()[-1]); print('$k row-owner', round(d['auto']['Melem_per_s'],1))"
done
timeout 600 python tools/bench_config4.py --nx 128 --layers 128 --steps 3 --nobox --no-momentum --no-oracle-check > gpurun_out/s4a_config4_nobox_128.json 2>> gpurun_out/s4a_err.txt
python -c "import json; d=json.loads(open('gpurun_out/s4a_config4_nobox_128.json').read().strip().splitlines()[-1]); print('config4 nobox: tke', round(d['tke']['kernel_melem_s_per_gpu'],1), 'sdr', round(d['sdr']['kernel_melem_s_per_gpu'],1))"
tail -3 gpurun_out/s4a_err.txt
