#!/bin/bash
# r01 final visit: full GPU suite, smoke, the default bench line (both arms), ncu launch list of the bench command
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; tail -2 gpurun_out/bench_default_final.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default_final.json') if l.startswith('{')][-1])
print('DEFAULT', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_read_only_stream'], 'e2e', round(d['e2e']['value']), d['clocks'], d['gpu_launches'])
print('  tts', d['time_to_solution'])
for k,v in d['variants'].items():
    print('  ', k, round(v['value']), round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['kernel_ms_per_step'].items()}, round(v['roofline']['frac'],3), v.get('time_to_solution',{}).get('solve_s'))
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
timeout 600 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/bench_reference_final.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_n125.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_launch_final.log 2>&1
tail -2 gpurun_out/ncu_launch_final.log | cut -c1-300
