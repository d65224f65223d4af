#!/bin/bash
# ncu evidence (1 GPU): launch list of the bench command + one --set full capture of the top kernel
set -x
mkdir -p gpurun_out
# every launch with its device time; same command as the bench, fewer steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_n125.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_launch_bench.log 2>&1
tail -3 gpurun_out/ncu_launch_bench.log | cut -c1-400
# the mat-vec kernel, full set, on an 80^3 cube (storkm 14.7 GB >> L2) so the ~40 replays stay cheap
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 4 -c 2 -f -o gpurun_out/prof_matvec_n80 \
    python bench.py --cube 80 --steps 4 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_full_bench.log 2>&1
tail -3 gpurun_out/ncu_full_bench.log | cut -c1-400
ls -la gpurun_out
# other element types, device-resident bench only (hex8 200^3 = BASELINE config D on one GPU)
timeout 600 python bench.py --hex 8 --cube 200 --steps 50 --no-cpu > gpurun_out/bench_hex8_n200.json 2> gpurun_out/bench_hex8_n200.err; cut -c1-1500 gpurun_out/bench_hex8_n200.json; tail -3 gpurun_out/bench_hex8_n200.err
