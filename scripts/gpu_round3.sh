#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=10 2>&1 | tail -30
timeout 600 python bench.py --matrix-free --steps 100 --no-cpu > gpurun_out/bench_mf_n125.json 2> gpurun_out/bench_mf.err; cut -c1-2600 gpurun_out/bench_mf_n125.json; tail -3 gpurun_out/bench_mf.err
timeout 600 python bench.py --matrix-free --hex 8 --cube 200 --steps 100 --no-cpu --no-solve > gpurun_out/bench_mf_hex8_n200.json 2> gpurun_out/bench_mf8.err; cut -c1-600 gpurun_out/bench_mf_hex8_n200.json; tail -3 gpurun_out/bench_mf8.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_mf -s 4 -c 1 -f -o gpurun_out/prof_mf_n125 \
    python bench.py --matrix-free --steps 3 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_mf.log 2>&1
tail -2 gpurun_out/ncu_mf.log | cut -c1-200
