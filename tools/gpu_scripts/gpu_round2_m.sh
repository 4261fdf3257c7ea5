#!/bin/bash
# bench.py at N ranks: does binding each rank to its GPU's NUMA node fix the e2e scaling?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/s3d_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name" > gpurun_out/s3d_lscpu.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s3d_bench_${N}gpu.json 2> gpurun_out/s3d_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s3d_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('serial_ms_per_step'), d['config'].get('numa_binding'))
PY
tail -3 gpurun_out/s3d_bench_${N}gpu.err; cat gpurun_out/s3d_lscpu.txt
