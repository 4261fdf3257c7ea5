#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=${1:-s4h}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_hex27.py -q -m gpu -k "cont or scalar or config4 or slist or tke or sdr or hex27 or golden" 2>&1 | tail -4
for k in continuity_adv scalar_adv_diff; do
  python tools/bench_continuity.py 256 10 $k > gpurun_out/${T}_${k}_256.json 2>> gpurun_out/${T}_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/${T}_${k}_256.json').read().strip().splitlines()[-1]); print('$k', {k:(round(v['Melem_per_s'],1), round(v['ms'],3)) for k,v in d.items() if isinstance(v,dict) and 'Melem_per_s' in v})"
done
timeout 600 python tools/bench_config4.py --nx 256 --layers 64 --steps 5 --no-momentum > gpurun_out/${T}_config4.json 2>> gpurun_out/${T}_err.txt
python -c "import json; d=json.loads(open('gpurun_out/${T}_config4.json').read().strip().splitlines()[-1]); print('config4 256x256x64: tke', round(d['tke']['kernel_melem_s_per_gpu'],1), 'sdr', round(d['sdr']['kernel_melem_s_per_gpu'],1), {k:d[k].get('oracle_check') for k in ('tke','sdr')})"
tail -3 gpurun_out/${T}_err.txt
