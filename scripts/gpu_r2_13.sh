#!/bin/bash
# round 2, visit 13 (1 GPU): the tensor-kernel family with ONE reciprocal per Gauss point (inv3_recip; factors written by
# k_apply_mf3<GEOM 1>): matrix-free parity tests in every kernel selection, timings of both modes, bench MF lines with the HBM side
mkdir -p gpurun_out
for sel in 4 3 2lane 1lane; do
PF_MF=$sel timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_fullsize.py -q -k "matrix_free" > gpurun_out/r2_13_pytest_$sel.log 2>&1
echo "PF_MF=$sel pytest rc=$?" >> gpurun_out/r2_13_pytest_$sel.log; tail -2 gpurun_out/r2_13_pytest_$sel.log
done
run() {
  name=$1; mode=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free $mode > gpurun_out/r2_13_$name.json 2> gpurun_out/r2_13_$name.err
  tail -1 gpurun_out/r2_13_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(r['frac'],4), round(r['frac_of_dfma_peak'],4), 'hbm side', round(r['hbm_side']['frac'],3), r['hbm_side']['traffic'])"
}
run mode2 2 PF_X=0
run mode1 1 PF_X=0
run mode1_c8 1 PF_MF4C=8
for mode in 2 1; do
timeout 600 python bench.py --hex 8 --cube 200 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free $mode > gpurun_out/r2_13_hex8_mode$mode.json 2>/dev/null
tail -1 gpurun_out/r2_13_hex8_mode$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('hex8 200^3 mode$mode', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(r['frac'],4), 'hbm side', round(r['hbm_side']['frac'],3))"
done
timeout 600 ncu --set full --clock-control none -k regex:k_apply_mf4 -s 6 -c 1 -f -o gpurun_out/r2_13_prof_mf4_mode1 \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 1 > gpurun_out/r2_13_ncu_mode1.log 2>&1
ncu -i gpurun_out/r2_13_prof_mf4_mode1.ncu-rep --page raw --csv > gpurun_out/r2_13_prof_mf4_mode1_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_13_prof_mf4_mode1_raw.csv')))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed'): print(h,u,v)
PY
rm -f gpurun_out/r2_13_prof_mf4_mode1.ncu-rep
