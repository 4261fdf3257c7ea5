#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/s4t_bench_${G}gpu_weak.json 2> gpurun_out/s4t_err.txt
tail -c 300 gpurun_out/s4t_err.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/s4t_bench_${G}gpu_weak.json').read().strip().splitlines()[-1])
print('$G gpu weak value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'graph_s', round(d['config']['graph_build_s'],1), d['checks'].get('sampled_rows_vs_oracle',{}).get('passed'))
PY
