#!/bin/bash
# round 2, visit 14 (1 GPU): the mat-vec kernel timed back to back (pf_measure_matvec) beside the per-launch-event figure:
# configs B, C, D and the matrix-free modes
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; b=r['back_to_back']
print('$2', round(d['value'],1), 'per-launch events', round(r['avg_launch_ms'],4), round(r['frac'],4), '| back to back', round(b['avg_launch_ms'],4), round(b['frac'],4))"; }
timeout 600 python bench.py --program p123 --cube 100 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_14_B.json 2> gpurun_out/r2_14_B.err; show gpurun_out/r2_14_B.json B_p123_100
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_14_C.json 2> gpurun_out/r2_14_C.err; show gpurun_out/r2_14_C.json C_hex20_125
timeout 600 python bench.py --hex 8 --cube 200 --steps 20 --warmup 5 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_14_D.json 2> gpurun_out/r2_14_D.err; show gpurun_out/r2_14_D.json D_hex8_200
timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_14_mf2.json 2> gpurun_out/r2_14_mf2.err; show gpurun_out/r2_14_mf2.json mf2
timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free 1 > gpurun_out/r2_14_mf1.json 2> gpurun_out/r2_14_mf1.err; show gpurun_out/r2_14_mf1.json mf1
timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --layout 1 > gpurun_out/r2_14_sym.json 2> gpurun_out/r2_14_sym.err; show gpurun_out/r2_14_sym.json sym_hex20
timeout 600 python bench.py --hex 8 --cube 200 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --layout 1 > gpurun_out/r2_14_sym8.json 2> gpurun_out/r2_14_sym8.err; show gpurun_out/r2_14_sym8.json sym_hex8
tail -3 gpurun_out/r2_14_*.err | tail -12
