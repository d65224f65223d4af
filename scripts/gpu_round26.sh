#!/bin/bash
# r01 round 26: diagonal-only mode of k_form_km_tiled (matrix-free setup)
set -x
timeout 600 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_transient.py -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --matrix-free 2 --steps 50 --no-cpu --no-solve > gpurun_out/r26_mf2.json 2>gpurun_out/r26.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r26_mf2.json') if l.startswith('{')][-1]); print('MF2', round(d['value']), d['ms_per_step'], d['setup_s'])"; tail -2 gpurun_out/r26.err
