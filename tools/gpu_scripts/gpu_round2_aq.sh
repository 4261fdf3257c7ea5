#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=${1:-8}
for S in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 10 --warmup 3 --scaling $S > gpurun_out/s4m_bench_${G}gpu_$S.json 2> gpurun_out/s4m_err_$S.txt
  tail -c 300 gpurun_out/s4m_err_$S.txt
  python - <<PY
import json
d=json.loads(open('gpurun_out/s4m_bench_${G}gpu_$S.json').read().strip().splitlines()[-1])
print('$G gpu $S value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'graph_s', round(d['config']['graph_build_s'],1), d['checks'].get('sampled_rows_vs_oracle'))
PY
done
