#!/bin/bash
# ncu --set full on the kernels beside the mat-vec at config C: k_form_km_tiled, then one iteration's
# k_scatter / k_dot / k_pcg_update / k_pupdate (the 3rd iteration)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_form_km_tiled -c 1 -f -o gpurun_out/prof_form_km_tiled_n125 \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-solve --no-variants > gpurun_out/ncu2_a.log 2>&1
tail -2 gpurun_out/ncu2_a.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_scatter|k_dot|k_pcg_update|k_pupdate" -s 9 -c 4 -f -o gpurun_out/prof_vector_n125 \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-solve --no-variants > gpurun_out/ncu2_b.log 2>&1
tail -2 gpurun_out/ncu2_b.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
