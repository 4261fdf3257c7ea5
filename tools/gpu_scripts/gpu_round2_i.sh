#!/bin/bash
# general kernel: registers bounded for 2 / 3 / 4 CTAs per SM (A/B at 128^3, momentum list)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for mb in 2 3 4; do
  NALU_GENERAL_MINB=$mb python tools/bench_momentum_list.py 128 3 > gpurun_out/s2k_general_minb$mb.json 2> gpurun_out/s2k_err.txt
  python -c "import json; d=json.loads(open('gpurun_out/s2k_general_minb$mb.json').read().strip().splitlines()[-1]); print('minb $mb general', round(d['general']['Melem_per_s'],1), 'Melem/s', round(d['general']['ms'],2), 'ms')"
done
