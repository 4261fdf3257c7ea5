#!/bin/bash
# quick A/B: box tests + 256^3 bench of assemble_box5 (+ optional z-chunk sweep)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box_kernel or row_thread or zrange" > gpurun_out/s2b_tests.log 2>&1
tail -2 gpurun_out/s2b_tests.log
for zc in 0 $ZCHUNKS; do
  NALU_BOX_ZCHUNK=$zc python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-continuity > gpurun_out/s2b_bench_zc$zc.json 2> gpurun_out/s2b_bench.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/s2b_bench_zc$zc.json').read().strip().splitlines()[-1]); print('zchunk $zc', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1))"
done
