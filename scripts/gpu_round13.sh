#!/bin/bash
# r01 round 13: register-pipelined gather in k_matvec / k_matvec_sym -- full GPU suite, default bench, tile shapes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_default_r13.json 2> gpurun_out/bench_default_r13.err; tail -2 gpurun_out/bench_default_r13.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default_r13.json') if l.startswith('{')][-1])
print('DEFAULT', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['clocks'])
print('  tts', d['time_to_solution'])
for k,v in d['variants'].items():
    print('  ', k, round(v['value']), round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['kernel_ms_per_step'].items()}, round(v['roofline']['frac'],3), v.get('time_to_solution',{}).get('solve_s'))
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
for tune in 1 2; do
  PF_TUNE=$tune timeout 300 python bench.py --layout 1 --steps 50 --no-cpu --no-solve > gpurun_out/sym2_t${tune}.json 2> gpurun_out/sym.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/sym2_t${tune}.json') if l.startswith('{')][-1]); print('SYM hex20 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
done
for tune in 0 1 2; do
  PF_TUNE=$tune timeout 300 python bench.py --layout 1 --hex 8 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/sym2_hex8_t${tune}.json 2> gpurun_out/sym.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/sym2_hex8_t${tune}.json') if l.startswith('{')][-1]); print('SYM hex8 tune $tune', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
done
for lay in 0 1; do
timeout 300 python bench.py --layout $lay --program p123 --cube 200 --steps 50 --no-cpu --no-solve > gpurun_out/p123_l${lay}.json 2> gpurun_out/sym.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/p123_l${lay}.json') if l.startswith('{')][-1]); print('p123 200 layout $lay', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
done
timeout 300 python bench.py --hex 8 --cube 200 --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/hex8_full_r13.json 2> gpurun_out/sym.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/hex8_full_r13.json') if l.startswith('{')][-1]); print('hex8 200 full', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/sym.err
