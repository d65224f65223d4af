#!/bin/bash
# round 2, visit 10 (1 GPU): k_apply_mf3 (FP64 tensor-core matrix-free kernel): parity tests, timings in both modes
# at config C and at 200^3 8-node bricks, warps-per-SM sweep, A/B against the two-lane kernel, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_fullsize.py -q -k "matrix_free" --durations=4 > gpurun_out/r2_10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_10_pytest.log; tail -12 gpurun_out/r2_10_pytest.log
run() {
  name=$1; mode=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free $mode > gpurun_out/r2_10_$name.json 2> gpurun_out/r2_10_$name.err
  tail -1 gpurun_out/r2_10_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(d['roofline']['frac'],4))"
}
run mode2_mf4_c8 2 PF_MF4C=8
run mode2_mf4_c12 2 PF_MF4C=12
run mode1_mf4_c8 1 PF_MF4C=8
run mode1_mf4_c12 1 PF_MF4C=12
for w in 8 12; do
for mode in 2 1; do
PF_MF4C=$w timeout 600 python bench.py --hex 8 --cube 200 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free $mode > gpurun_out/r2_10_hex8_mode${mode}_c$w.json 2>/dev/null
tail -1 gpurun_out/r2_10_hex8_mode${mode}_c$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hex8 200^3 mode$mode c$w', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(d['roofline']['frac'],4))"
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf4 -s 6 -c 1 -f -o gpurun_out/r2_10_prof_mf4 \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_10_ncu_mf4.log 2>&1
ncu -i gpurun_out/r2_10_prof_mf4.ncu-rep --page raw --csv > gpurun_out/r2_10_prof_mf4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_10_prof_mf4.ncu-rep --page source --csv > gpurun_out/r2_10_prof_mf4_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_10_prof_mf4_raw.csv gpurun_out/r2_10_prof_mf4_src.csv 300 > gpurun_out/r2_10_prof_mf4_summary.txt 2>&1
head -40 gpurun_out/r2_10_prof_mf4_summary.txt
rm -f gpurun_out/r2_10_prof_mf4.ncu-rep
