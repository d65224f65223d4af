#!/bin/bash
# r01 round 23: scatter tables built on the device (stable radix sort), pf_make_ggl's scan in parallel
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/time_setup.py 2>&1 | tail -3
timeout 300 python scripts/time_setup.py 200 8 2>&1 | tail -3
