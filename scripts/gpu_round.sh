#!/bin/bash
# one GPU box visit: parity tests, bench at two sizes, ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30
timeout 600 python bench.py --cube 40 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n40.json 2> gpurun_out/bench_n40.err; tail -c 3000 gpurun_out/bench_n40.json; tail -5 gpurun_out/bench_n40.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n125.json 2> gpurun_out/bench_n125.err; tail -c 4000 gpurun_out/bench_n125.json; tail -5 gpurun_out/bench_n125.err
