#!/bin/bash
# round 2, visit 5 (1 GPU): two-lane matrix-free kernel, p122, pcg_km, tetrahedron rules; bench; ncu of the variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_plastic.py tests/test_gpu_tetrahedra.py tests/test_gpu_parity.py -q --durations=5 > gpurun_out/r2_05_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_05_pytest.log
tail -25 gpurun_out/r2_05_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/r2_05_bench_mf2lane.json 2> gpurun_out/r2_05_bench_mf2lane.err
PF_MF=1lane timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu --no-solve --matrix-free 2 --no-variants > gpurun_out/r2_05_bench_mf1lane.json 2> gpurun_out/r2_05_bench_mf1lane.err
for f in gpurun_out/r2_05_bench_mf*.json; do echo $f; tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])
for k,v in (d.get('variants') or {}).items(): print(' ',k,v['value'],v['kernel_ms_per_step']['matvec'],v['roofline']['frac'])
"; done
# ncu --set full: the two-lane matrix-free kernel (mode 2) and the packed-layout mat-vec at config C (DRAM traffic)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf2 -s 6 -c 1 -f -o gpurun_out/r2_05_prof_mf2 \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_05_ncu_mf2.log 2>&1
ncu -i gpurun_out/r2_05_prof_mf2.ncu-rep --page raw --csv > gpurun_out/r2_05_prof_mf2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_05_prof_mf2.ncu-rep --page source --csv > gpurun_out/r2_05_prof_mf2_src.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:k_matvec_sym -s 6 -c 1 -f -o gpurun_out/r2_05_prof_sym \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --layout 1 > gpurun_out/r2_05_ncu_sym.log 2>&1
ncu -i gpurun_out/r2_05_prof_sym.ncu-rep --page raw --csv > gpurun_out/r2_05_prof_sym_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_05_prof_mf2_raw.csv gpurun_out/r2_05_prof_mf2_src.csv 300 > gpurun_out/r2_05_prof_mf2_summary.txt 2>&1
head -40 gpurun_out/r2_05_prof_mf2_summary.txt
ls -la gpurun_out | grep r2_05
