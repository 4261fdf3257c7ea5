#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=${1:-2}
timeout 600 python tools/bench_hex27.py 48 10 3 > gpurun_out/s4j_hex27_1gpu.json 2> gpurun_out/s4j_err.txt
tail -c 300 gpurun_out/s4j_err.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29519 tools/bench_hex27.py 48 10 3 > gpurun_out/s4j_hex27_${G}gpu.json 2>> gpurun_out/s4j_err.txt
tail -c 600 gpurun_out/s4j_err.txt
python - <<PY
import json
for f in ("gpurun_out/s4j_hex27_1gpu.json", "gpurun_out/s4j_hex27_${G}gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'asm', round(d['assemble_ms'],2), d['checks'])
    except Exception as e:
        print(f, 'failed', e)
PY
