#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30
timeout 600 python bench.py --program p123 --cube 100 --steps 100 --no-cpu > gpurun_out/bench_p123_n100.json 2> gpurun_out/bench_p123.err; cut -c1-1800 gpurun_out/bench_p123_n100.json; tail -3 gpurun_out/bench_p123.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 4 -c 1 -f -o gpurun_out/prof_matvec_n125 \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_full_n125.log 2>&1
tail -3 gpurun_out/ncu_full_n125.log | cut -c1-300
