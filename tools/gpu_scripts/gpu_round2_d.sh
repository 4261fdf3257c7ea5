#!/bin/bash
# N-GPU: weak + strong scaling bench with the library's loadComplete (N = $1)
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
for sc in weak strong; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/s2e_bench_${N}gpu_$sc.json 2> gpurun_out/s2e_bench_${N}gpu_$sc.err
  tail -c 600 gpurun_out/s2e_bench_${N}gpu_$sc.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s2e_bench_${N}gpu_$sc.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('$sc N=$N value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), 'exposed halo', round(r['exposed_halo_ms'],3), 'e2e', round(d['e2e']['value'],1))
    print('  checks', d['checks'])
    print('  halo', d['config']['halo'])
except Exception as e:
    print('parse failed', e)
PY
done
