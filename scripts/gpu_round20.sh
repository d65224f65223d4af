#!/bin/bash
# r01 round 20: register-tiled km formation (k_form_km_tiled) -- parity, then old vs new timing
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_materials.py tests/test_gpu_symmetric.py -m gpu -q -x 2>&1 | tail -8
PF_FORM=old timeout 300 python scripts/time_form_km.py 2>&1 | tail -3
timeout 300 python scripts/time_form_km.py 2>&1 | tail -3
