#!/bin/bash
# ncu --set full of the mat-vec at BASELINE config B (p123, 100^3 8-node bricks) and config D (p121 hex8 200^3)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -f -o gpurun_out/prof_matvec_p123_n100 \
    python bench.py --program p123 --cube 100 --steps 4 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu3_a.log 2>&1
tail -1 gpurun_out/ncu3_a.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -f -o gpurun_out/prof_matvec_hex8_n200 \
    python bench.py --hex 8 --cube 200 --steps 4 --warmup 3 --no-cpu --no-solve --no-variants > gpurun_out/ncu3_b.log 2>&1
tail -1 gpurun_out/ncu3_b.log | cut -c1-200
