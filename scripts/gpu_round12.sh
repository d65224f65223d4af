#!/bin/bash
# r01 round 12: packed-lower-triangle storkm layout -- parity, tile shapes, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_symmetric.py -m gpu -x -q 2>&1 | tail -8
for tune in 0 1 2; do
  PF_TUNE=$tune timeout 300 python bench.py --layout 1 --steps 50 --no-cpu --no-solve > gpurun_out/sym_t${tune}.json 2> gpurun_out/sym.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/sym_t${tune}.json') if l.startswith('{')][-1]); print('SYM hex20 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'])"; tail -2 gpurun_out/sym.err
done
for tune in 0 1; do
  PF_TUNE=$tune timeout 300 python bench.py --layout 1 --hex 8 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/sym_hex8_t${tune}.json 2> gpurun_out/sym.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/sym_hex8_t${tune}.json') if l.startswith('{')][-1]); print('SYM hex8 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
done
timeout 300 python bench.py --layout 1 --program p123 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/sym_p123.json 2> gpurun_out/sym.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/sym_p123.json') if l.startswith('{')][-1]); print('SYM p123', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_sym -s 4 -c 1 -f -o gpurun_out/prof_sym_n125 \
    python bench.py --layout 1 --steps 3 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_sym.log 2>&1
tail -2 gpurun_out/ncu_sym.log | cut -c1-200
