#!/bin/bash
# r01 round 22: CUDA-graph replay of the PCG iteration (single rank): parity, then with / without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for g in 0 1; do
  PF_GRAPH=$g timeout 300 python bench.py --program p124 --cube 100 --steps 100 > gpurun_out/r22_p124_g$g.json 2> gpurun_out/r22.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r22_p124_g$g.json') if l.startswith('{')][-1]); print('p124 100 graph=$g', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['pcg_iterations_timed'], d['gpu_launches'])"; tail -2 gpurun_out/r22.err
  PF_GRAPH=$g timeout 300 python bench.py --program p123 --cube 100 --steps 200 --no-cpu --no-solve > gpurun_out/r22_p123_g$g.json 2> gpurun_out/r22.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r22_p123_g$g.json') if l.startswith('{')][-1]); print('p123 100 graph=$g (value is profiled, e2e not)', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"; tail -2 gpurun_out/r22.err
  PF_GRAPH=$g timeout 300 python bench.py --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/r22_c_g$g.json 2> gpurun_out/r22.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r22_c_g$g.json') if l.startswith('{')][-1]); print('C graph=$g', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"; tail -2 gpurun_out/r22.err
done
