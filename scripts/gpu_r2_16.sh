#!/bin/bash
# round 2, visit 16 (1 GPU): full GPU suite + smoke + the default bench line (timed) after p1210, k_matvec2, the
# reciprocal inverse of the tensor-kernel family and the extra roofline keys
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_16_pytest.log
tail -9 gpurun_out/r2_16_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_16_bench_default.json 2> gpurun_out/r2_16_bench_default.err
echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
tail -c 400 gpurun_out/r2_16_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_16_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['back_to_back']['frac'], d['parity_check']['bit_equal'], d['clocks'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), v['kernel_ms_per_step'], round(v['roofline']['frac'],4), round(v['roofline']['back_to_back']['frac'],4))
for k,v in d['variants'].items(): print(k, round(v['value'],1), v['kernel_ms_per_step'], round(v['roofline']['frac'],4), v['roofline'].get('frac_of_dfma_peak'), (v['roofline'].get('hbm_side') or {}).get('frac'), (v.get('time_to_solution') or {}).get('solve_s'))
print('weak', d['weak']['value']); print('cpu', d['cpu_baseline'])
PY
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r2_16_bench_reference.json 2> gpurun_out/r2_16_bench_reference.err
echo "reference arm rc=$? wall $(( $(date +%s) - t0 )) s"; tail -c 600 gpurun_out/r2_16_bench_reference.json
