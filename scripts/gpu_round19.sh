#!/bin/bash
# r01 round 19: C++ scalar driver (p123/p124/p125) + xx11 golden + pf_sum on the GPU
set -x
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15
