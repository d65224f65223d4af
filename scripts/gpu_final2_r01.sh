#!/bin/bash
# r01 final visit of this session: full GPU suite, smoke, the default bench line (both arms), ncu launch list of
# the bench command, and the other BASELINE configs that fit one GPU (D: hex8 200^3, B: p123 100^3)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; tail -2 gpurun_out/bench_default_final.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default_final.json') if l.startswith('{')][-1])
print('DEFAULT', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_read_only_stream'], 'e2e', round(d['e2e']['value']), d['clocks'], d['gpu_launches'], d['setup_s'])
print('  tts', d['time_to_solution'])
for k,v in d['variants'].items():
    print('  ', k, round(v['value']), round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['kernel_ms_per_step'].items()}, round(v['roofline']['frac'],3), v.get('time_to_solution',{}).get('solve_s'))
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
timeout 600 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/bench_reference_final.json 2>/dev/null; cut -c1-200 gpurun_out/bench_reference_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_n125.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_launch_final.log 2>&1
tail -1 gpurun_out/ncu_launch_final.log | cut -c1-200
timeout 300 python bench.py --hex 8 --cube 200 --steps 100 --no-cpu --no-variants > gpurun_out/bench_hex8_n200_final.json 2>/dev/null
python -c "import json; d=json.loads([l for l in open('gpurun_out/bench_hex8_n200_final.json') if l.startswith('{')][-1]); print('D hex8 200', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['time_to_solution'])"
timeout 300 python bench.py --program p123 --cube 100 --steps 200 --no-cpu > gpurun_out/bench_p123_n100_final.json 2>/dev/null
python -c "import json; d=json.loads([l for l in open('gpurun_out/bench_p123_n100_final.json') if l.startswith('{')][-1]); print('B p123 100', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['time_to_solution'])"
