#!/bin/bash
# round 2, visit 15 (1 GPU): k_matvec2 (several ring slots per warp) on the 8x8 matrices of p123: tile shapes, parity
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; b=r['back_to_back']
print('$2', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'per-launch events', round(r['avg_launch_ms'],4), round(r['frac'],4), '| back to back', round(b['avg_launch_ms'],4), round(b['frac'],4))"; }
for tune in 0 2 3 4 5 6; do
PF_TUNE=$tune timeout 600 python bench.py --program p123 --cube 100 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_15_B_t$tune.json 2> gpurun_out/r2_15_B_t$tune.err; show gpurun_out/r2_15_B_t$tune.json B_p123_100_tune$tune
done
for tune in 0 2 4; do
PF_TUNE=$tune timeout 600 python bench.py --program p123 --cube 200 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_15_p123_200_t$tune.json 2> gpurun_out/r2_15_p123_200_t$tune.err; show gpurun_out/r2_15_p123_200_t$tune.json p123_200_tune$tune
done
for tune in 2 3 4 5 6; do
PF_TUNE=$tune timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_transient.py tests/test_gpu_explicit.py -q -k "p123 or p124 or p125 or transient or explicit" 2>&1 | tail -1
done
