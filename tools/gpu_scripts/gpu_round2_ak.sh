#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s4i_bench_2gpu_weak.json 2> gpurun_out/s4i_err.txt
tail -c 600 gpurun_out/s4i_err.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/s4i_bench_2gpu_weak.json').read().strip().splitlines()[-1])
print('2gpu weak value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['scaling'], d['checks'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s4i_ref_2gpu.json 2>> gpurun_out/s4i_err.txt
tail -c 300 gpurun_out/s4i_ref_2gpu.json
