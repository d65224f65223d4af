#!/bin/bash
# r01 round 15: k_matvec with the ring slot refilled in column chunks (hex20), vs the read-only stream ceiling
set -x
mkdir -p gpurun_out
for tune in 3 4; do
  PF_TUNE=$tune timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hex20" 2>&1 | tail -2
done
for tune in 0 3 4 5 6; do
  PF_TUNE=$tune timeout 300 python bench.py --steps 50 --no-cpu --no-solve --no-variants > gpurun_out/r15_t${tune}.json 2> gpurun_out/r15.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r15_t${tune}.json') if l.startswith('{')][-1]); r=d['roofline']; print('chunk tune $tune', round(d['value']), d['ms_per_step'], round(r['avg_launch_ms'],4), round(r['frac'],4), round(r['frac_of_read_only_stream'],4), round(r['read_only_stream_gbs']))"; tail -2 gpurun_out/r15.err
done
