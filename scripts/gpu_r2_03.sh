#!/bin/bash
# round 2, visit 3 (1 GPU): p122 on the device + the whole GPU suite after the kernel signature changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plastic.py -x -q --durations=5 > gpurun_out/r2_03_plastic.log 2>&1
echo "plastic rc=$?" >> gpurun_out/r2_03_plastic.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_plastic.py > gpurun_out/r2_03_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_03_pytest.log
tail -30 gpurun_out/r2_03_plastic.log; tail -12 gpurun_out/r2_03_pytest.log
