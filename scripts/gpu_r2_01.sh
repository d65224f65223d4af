#!/bin/bash
# round 2, visit 1 (1 GPU): full GPU test suite, then the default bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_01_gpus.txt 2>&1
free -g > gpurun_out/r2_01_mem.txt; nproc >> gpurun_out/r2_01_mem.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_01_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_01_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
echo "bench rc=$?" >> gpurun_out/r2_01_bench.err
tail -5 gpurun_out/r2_01_pytest.log; tail -3 gpurun_out/r2_01_bench.err; head -c 600 gpurun_out/r2_01_bench.json
