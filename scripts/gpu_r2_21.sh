#!/bin/bash
# round 2, visit 21 (1 GPU): does a common shared-memory carve-out for the kernels of an iteration remove SM
# reconfiguration between them?  A/B at configs B, C, D and matrix-free (graph replay, ms per step)
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'])"; }
for c in 0 1 0 1; do
PF_CARVEOUT=$c timeout 600 python bench.py --program p123 --cube 100 --steps 100 --warmup 10 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_21_B_c$c.json 2>/dev/null; show gpurun_out/r2_21_B_c$c.json B_carveout$c
done
for c in 0 1; do
PF_CARVEOUT=$c timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_21_C_c$c.json 2>/dev/null; show gpurun_out/r2_21_C_c$c.json C_carveout$c
PF_CARVEOUT=$c timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_21_mf2_c$c.json 2>/dev/null; show gpurun_out/r2_21_mf2_c$c.json mf2_carveout$c
done
