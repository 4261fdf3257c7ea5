#!/bin/bash
# scalar-list tiled kernel (config 4): new parity tests, per-GPU config-4 bench (512x512x64 slab, tke / sdr / momentum), general
# path for comparison at a smaller size, ncu capture of the slist kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config4 or system_parity" > gpurun_out/s3a_tests.txt 2>&1
tail -15 gpurun_out/s3a_tests.txt
timeout 900 python tools/bench_config4.py --nx 512 --layers 64 --steps 10 > gpurun_out/s3a_config4_1gpu.json 2> gpurun_out/s3a_config4_1gpu.err
tail -c 3000 gpurun_out/s3a_config4_1gpu.json; tail -5 gpurun_out/s3a_config4_1gpu.err
timeout 600 python tools/bench_config4.py --nx 128 --layers 128 --steps 3 --mode general --no-momentum --no-oracle-check > gpurun_out/s3a_config4_general_128.json 2> gpurun_out/s3a_config4_general.err
tail -c 1500 gpurun_out/s3a_config4_general_128.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_slist -c 1 -s 3 -o gpurun_out/s3a_slist_tke -f \
  python tools/bench_config4.py --nx 512 --layers 64 --steps 1 --no-momentum --no-oracle-check > gpurun_out/s3a_ncu.log 2>&1
tail -3 gpurun_out/s3a_ncu.log
