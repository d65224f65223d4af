#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
for mode in 1 2; do for tune in 0 1; do
  PF_TUNE=$tune timeout 300 python bench.py --matrix-free $mode --steps 100 --no-cpu --no-solve > gpurun_out/mf_m${mode}_t${tune}.json 2> gpurun_out/mf.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/mf_m${mode}_t${tune}.json') if l.startswith('{')][-1]); print('MF mode $mode tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/mf.err
done; done
for tune in 0 1 2 3 4; do
  PF_TUNE=$tune timeout 300 python bench.py --hex 8 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/hex8_t${tune}.json 2> gpurun_out/h8.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/hex8_t${tune}.json') if l.startswith('{')][-1]); print('hex8 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/h8.err
done
for tune in 0 1 2; do
  PF_TUNE=$tune timeout 300 python bench.py --program p123 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/p123_t${tune}.json 2> gpurun_out/p123.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/p123_t${tune}.json') if l.startswith('{')][-1]); print('p123 200^3 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/p123.err
done
