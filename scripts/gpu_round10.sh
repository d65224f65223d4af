#!/bin/bash
# r01 round 10: MF kernel v4 (node loops not fully unrolled, factors two points ahead, bulk L2 prefetch)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrix_free.py -m gpu -x -q 2>&1 | tail -8
for mode in 2 1; do
  timeout 300 python bench.py --matrix-free $mode --steps 100 --no-cpu --no-solve > gpurun_out/mf6_m${mode}.json 2> gpurun_out/mf.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/mf6_m${mode}.json') if l.startswith('{')][-1]); print('MF mode $mode', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])"; tail -2 gpurun_out/mf.err
done
timeout 300 python bench.py --matrix-free 2 --hex 8 --cube 200 --steps 100 --no-cpu --no-solve > gpurun_out/mf6_hex8.json 2> gpurun_out/mf.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/mf6_hex8.json') if l.startswith('{')][-1]); print('MF hex8 mode 2', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/mf.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf -s 4 -c 1 -f -o gpurun_out/prof_mf6_n125 \
    python bench.py --matrix-free 2 --steps 3 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_mf6.log 2>&1
tail -2 gpurun_out/ncu_mf6.log | cut -c1-200
