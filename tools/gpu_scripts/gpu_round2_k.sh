#!/bin/bash
# BASELINE config 4 as specified: 512^3 ABL-style box on 8 GPUs (z-slabs of 64 layers), tke / sdr / momentum lists + loadComplete,
# sampled rows vs the oracle across the rank interfaces
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/bench_config4.py --nx 512 --layers 64 --steps 10 > gpurun_out/s3b_config4_${N}gpu.json 2> gpurun_out/s3b_config4_${N}gpu.err
tail -c 3500 gpurun_out/s3b_config4_${N}gpu.json; tail -5 gpurun_out/s3b_config4_${N}gpu.err
