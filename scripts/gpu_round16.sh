#!/bin/bash
# r01 round 16: k_scatter with four contributions in flight
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_symmetric.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/r16_c.json 2> gpurun_out/r16.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r16_c.json') if l.startswith('{')][-1]); print('C full', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/r16.err
timeout 300 python bench.py --hex 8 --cube 200 --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/r16_d.json 2> gpurun_out/r16.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r16_d.json') if l.startswith('{')][-1]); print('hex8 200 full', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/r16.err
timeout 300 python bench.py --program p123 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/r16_p.json 2> gpurun_out/r16.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r16_p.json') if l.startswith('{')][-1]); print('p123 200', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/r16.err
