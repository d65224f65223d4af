#!/bin/bash
# r01 round 18: 2 GPUs -- N-rank parity including the new drivers (p124, p124_fixed, p125, hex20_mat)
set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -30
