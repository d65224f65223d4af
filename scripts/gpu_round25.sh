#!/bin/bash
# r01 round 25: 4-node tetrahedra, shuffled numbering, line-based .lds reader -- whole 1-GPU suite
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
