#!/bin/bash
# round 2, visit 9 (1 GPU): matrix-free defaults after tuning (half-warp pairs, 12 warps, one exchange per two node pairs);
# parity tests; mode 1 through the two-lane kernel (A/B); synccheck with default settings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_fullsize.py -q -k "matrix_free" --durations=4 > gpurun_out/r2_09_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_09_pytest.log; tail -6 gpurun_out/r2_09_pytest.log
run() {
  name=$1; mode=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free $mode > gpurun_out/r2_09_$name.json 2> gpurun_out/r2_09_$name.err
  tail -1 gpurun_out/r2_09_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(d['roofline']['frac'],4))"
}
run mode2_default 2 PF_X=0
run mode2_pair1 2 PF_MF2=1
run mode2_w16 2 PF_MF2W=16
run mode1_onelane 1 PF_X=0
run mode1_twolane 1 PF_MF1=2lane
timeout 600 python bench.py --hex 8 --cube 200 --steps 50 --warmup 5 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_09_hex8_mode2.json 2>/dev/null
tail -1 gpurun_out/r2_09_hex8_mode2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hex8 200^3 mode2', round(d['value'],1), d['kernel_ms_per_step']['matvec'], round(d['roofline']['frac'],4))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf2 -s 6 -c 1 -f -o gpurun_out/r2_09_prof_mf2 \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_09_ncu_mf2.log 2>&1
ncu -i gpurun_out/r2_09_prof_mf2.ncu-rep --page raw --csv > gpurun_out/r2_09_prof_mf2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_09_prof_mf2.ncu-rep --page source --csv > gpurun_out/r2_09_prof_mf2_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_09_prof_mf2_raw.csv gpurun_out/r2_09_prof_mf2_src.csv 300 > gpurun_out/r2_09_prof_mf2_summary.txt 2>&1
head -32 gpurun_out/r2_09_prof_mf2_summary.txt
rm -f gpurun_out/r2_09_prof_mf2.ncu-rep
