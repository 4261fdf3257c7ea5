#!/bin/bash
# end-of-round evidence on one B200: full GPU suite, smoke(), the driver's bench command (both arms), the ncu launch list of
# the same bench command
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s4s_gpu_suite.txt 2>&1
tail -4 gpurun_out/s4s_gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4s_smoke.txt 2>&1
tail -6 gpurun_out/s4s_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s4s_bench_1gpu.json 2> gpurun_out/s4s_bench_1gpu.err
tail -c 400 gpurun_out/s4s_bench_1gpu.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/s4s_bench_reference_arm.json 2> gpurun_out/s4s_ref.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s4s_bench_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value'],1), 'frac', round(r['frac'],4), 'kernel ms', round(r['kernel_ms'],2), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'])
print('single', r['single_kernel']['value'], 'm+c', r['momentum_plus_continuity']['frac'], 'fp64', r.get('fp64'))
print('checks', d['checks'], d['clocks'])
print('other', d.get('other_configs'))
rr=json.loads(open('gpurun_out/s4s_bench_reference_arm.json').read().strip().splitlines()[-1])
print('reference arm', rr['value'], rr['unit'], rr['cpu_baseline']['cores'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s4s_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-oracle-check > gpurun_out/s4s_ncu_bench.log 2>&1
tail -2 gpurun_out/s4s_ncu_bench.log | cut -c1-300
