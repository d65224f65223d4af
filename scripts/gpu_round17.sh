#!/bin/bash
# r01 round 17: first GPU visit of the 8f-rank-3 drivers (p124 transient, p125 explicit, xx2 materials) +
# a sanity bench of the headline path (the mat-vec launch now takes a selectable matrix set)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transient.py tests/test_gpu_materials.py tests/test_gpu_explicit.py -m gpu -q 2>&1 | tail -25
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/r17_c.json 2> gpurun_out/r17.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r17_c.json') if l.startswith('{')][-1]); print('C full', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/r17.err
timeout 300 python bench.py --program p124 --cube 100 --steps 100 > gpurun_out/r17_p124.json 2> gpurun_out/r17.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r17_p124.json') if l.startswith('{')][-1]); print('p124 100', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['pcg_iterations_timed'], d['gpu_launches'])"; tail -2 gpurun_out/r17.err
timeout 300 python bench.py --program p125 --cube 200 --steps 500 > gpurun_out/r17_p125.json 2> gpurun_out/r17.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r17_p125.json') if l.startswith('{')][-1]); print('p125 200', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; tail -2 gpurun_out/r17.err
