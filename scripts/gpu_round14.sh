#!/bin/bash
# r01 round 14: fewer barriers in k_pcg_update, exit test folded into k_pupdate; MF unroll factors
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --no-cpu --no-solve --no-variants > gpurun_out/r14_full.json 2> gpurun_out/r14.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r14_full.json') if l.startswith('{')][-1]); print('C full', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; tail -2 gpurun_out/r14.err
for tune in 0 1 2; do
  PF_TUNE=$tune timeout 300 python bench.py --matrix-free 2 --steps 100 --no-cpu --no-solve > gpurun_out/r14_mf_t${tune}.json 2> gpurun_out/r14.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r14_mf_t${tune}.json') if l.startswith('{')][-1]); print('MF2 unroll-tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/r14.err
done
timeout 300 python bench.py --program p123 --cube 100 --steps 200 --no-cpu > gpurun_out/r14_p123_100.json 2> gpurun_out/r14.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r14_p123_100.json') if l.startswith('{')][-1]); print('p123 100 (config B)', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['time_to_solution'])"; tail -2 gpurun_out/r14.err
