#!/bin/bash
# round 2, visit 18 (1 GPU): ncu launch list of the bench command (round-2 build), ncu --set full of k_matvec2 (config B)
# and of the shipped k_apply_mf4 (mode 2) -> profiles/
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_18_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-variants --no-cpu --no-solve > gpurun_out/r2_18_launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_18_launches.csv')) if len(r) > 5]
hdr = rows[0]; k = hdr.index('Kernel Name'); v = hdr.index('Metric Value'); u = hdr.index('Metric Unit')
tot = collections.OrderedDict()
for r in rows[1:]:
    try: t = float(r[v].replace(',', ''))
    except ValueError: continue
    t *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0}.get(r[u], 1e-3)
    name = r[k].split('(')[0][:60]
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:12]: print(f'{n:62s} {c:5d} launches {t:10.3f} ms')
PY
timeout 600 ncu --set full --clock-control none -k regex:k_matvec2 -s 6 -c 1 -f -o gpurun_out/r2_18_prof_matvec2 \
    python bench.py --program p123 --cube 100 --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants > gpurun_out/r2_18_ncu_matvec2.log 2>&1
ncu -i gpurun_out/r2_18_prof_matvec2.ncu-rep --page raw --csv > gpurun_out/r2_18_prof_matvec2_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf4 -s 6 -c 1 -f -o gpurun_out/r2_18_prof_mf4 \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-solve --no-variants --matrix-free 2 > gpurun_out/r2_18_ncu_mf4.log 2>&1
ncu -i gpurun_out/r2_18_prof_mf4.ncu-rep --page raw --csv > gpurun_out/r2_18_prof_mf4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_18_prof_mf4.ncu-rep --page source --csv > gpurun_out/r2_18_prof_mf4_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_18_prof_mf4_raw.csv gpurun_out/r2_18_prof_mf4_src.csv 300 > gpurun_out/r2_18_prof_mf4_summary.txt 2>&1
for f in matvec2 mf4; do python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2_18_prof_${f}_raw.csv')))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread'): print('$f',h,u,v)
PY
done
rm -f gpurun_out/*.ncu-rep
