#!/bin/bash
# thin slab (256 x 256 x 32, one rank of the strong-scaled 256^3 box on 8 GPUs): z chunks 16 + 16 + 1 (old) vs balanced (new)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for Z in 16 0; do
  NALU_BOX_ZCHUNK=$Z timeout 600 python tools/bench_config4.py --nx 256 --layers 32 --steps 10 --no-oracle-check > gpurun_out/s4w_slab32_z$Z.json 2> gpurun_out/s4w_err_$Z.txt
  python -c "import json; d=json.loads(open('gpurun_out/s4w_slab32_z$Z.json').read().strip().splitlines()[-1]); m=d['momentum']; print('zchunk', $Z, {k:(round(v,3) if isinstance(v,float) else v) for k,v in m.items() if k in ('kernel_ms','kernel_melem_s_per_gpu','ms_per_step','assemble_mode')})"
  tail -2 gpurun_out/s4w_err_$Z.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "box_kernel or fused_lumped or zrange" 2>&1 | tail -3
