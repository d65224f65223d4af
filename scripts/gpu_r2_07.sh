#!/bin/bash
# round 2, visit 7 (1 GPU): matrix-free kernel variants (lane pairing, warps per SM): kernel time + a few ncu counters
mkdir -p gpurun_out
M="l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_07_$name.json 2> gpurun_out/r2_07_$name.err
  tail -1 gpurun_out/r2_07_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(d['roofline']['frac'],4))"
  env "$@" timeout 600 ncu --metrics $M --clock-control none -k regex:k_apply_mf -s 6 -c 1 --csv --log-file gpurun_out/r2_07_ncu_$name.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > /dev/null 2>&1
  grep -v "^==" gpurun_out/r2_07_ncu_$name.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; i=h.index('Metric Name'); j=h.index('Metric Value')
print('   ', {r[i].split('__')[-1][:40]: r[j] for r in rows[1:]})"
}
run pair1_w16 PF_MF2=1
run pair16_w16 PF_MF2=16
run pair1_w12 PF_MF2=1 PF_MF2W=12
run pair16_w12 PF_MF2=16 PF_MF2W=12
run onelane PF_MF=1lane

