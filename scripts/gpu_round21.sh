#!/bin/bash
# r01 round 21: the xx3-compatible boundary on the GPU
set -x
timeout 600 python -m pytest tests/test_gpu_xx3_compat.py -m gpu -q 2>&1 | tail -15
