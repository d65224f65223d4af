#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -16
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-6000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
PF_FUSE=0 timeout 600 python bench.py --steps 100 --no-cpu --no-solve --no-variants > gpurun_out/bench_nofuse.json 2>/dev/null; python -c "import json; d=json.loads([l for l in open('gpurun_out/bench_nofuse.json') if l.startswith('{')][-1]); print('nofuse', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])"
timeout 600 python bench.py --cube 100 --steps 100 --no-cpu --no-solve --no-variants > gpurun_out/bench_weak_base_n100.json 2>/dev/null; python -c "import json; d=json.loads([l for l in open('gpurun_out/bench_weak_base_n100.json') if l.startswith('{')][-1]); print('n100', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'])"
for tune in 0 2 3; do
  PF_TUNE=$tune timeout 300 python bench.py --hex 8 --cube 200 --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/hex8_t${tune}.json 2> gpurun_out/h8.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/hex8_t${tune}.json') if l.startswith('{')][-1]); print('hex8 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/h8.err
  PF_TUNE=$tune timeout 300 python bench.py --program p123 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/p123_t${tune}.json 2> gpurun_out/p123.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/p123_t${tune}.json') if l.startswith('{')][-1]); print('p123 200^3 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/p123.err
done
