#!/bin/bash
# first GPU visit of round 2: box tests, A/B bench of the two tiled momentum kernels, ncu capture of assemble_box5
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "box_kernel or row_thread or zrange or smoke" > gpurun_out/s2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/s2a_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s2a_bench_box5.json 2> gpurun_out/s2a_bench_box5.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-continuity --box-kernel 2 > gpurun_out/s2a_bench_box2.json 2> gpurun_out/s2a_bench_box2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_box5 -s 2 -c 1 -o gpurun_out/s2a_box5_256 \
  python bench.py --n 192 --steps 1 --warmup 1 --no-cpu-baseline --no-continuity > gpurun_out/s2a_ncu.log 2>&1
tail -3 gpurun_out/s2a_tests.log
cat gpurun_out/s2a_bench_box5.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('box5', d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('momentum_list'))"
cat gpurun_out/s2a_bench_box2.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('box2', d['value'], d['roofline']['frac'])"
