#!/bin/bash
# round 2, visit 24 (1 GPU): the round's last validation -- full GPU suite, smoke, default bench line, ncu --set full of
# the p1210 tensor-core kernel at 100^3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/r2_24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_24_pytest.log
tail -7 gpurun_out/r2_24_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_24_bench_default.json 2> gpurun_out/r2_24_bench_default.err
echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_24_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_check']['bit_equal'], d['clocks'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['roofline']['frac'],4))
for k,v in d['variants'].items(): print(k, round(v['value'],1), round(v['roofline']['frac'],4))
PY
timeout 600 ncu --set full --clock-control none -k regex:k_apply_mf4 -s 4 -c 1 -f -o gpurun_out/r2_24_prof_p1210 \
    python bench.py --program p1210 --cube 100 --steps 3 --warmup 3 --matrix-free 1 > gpurun_out/r2_24_ncu_p1210.log 2>&1
ncu -i gpurun_out/r2_24_prof_p1210.ncu-rep --page raw --csv > gpurun_out/r2_24_prof_p1210_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_24_prof_p1210_raw.csv')))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'): print(h,u,v)
PY
rm -f gpurun_out/*.ncu-rep
