#!/bin/bash
# round 2, visit 22 (N GPUs, N = $1): the default bench line on the round's last build (scaling table for profiles/)
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 2983$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_22_bench_g$N.json 2> gpurun_out/r2_22_bench_g$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_22_bench_g$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['parity_check']['bit_equal'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), v['kernel_ms_per_step'])
for k,v in d['variants'].items(): print(k, round(v['value'],1), v['ms_per_step'], v['kernel_ms_per_step'])
print('weak', d['weak']['value'], 'tts', d['time_to_solution']['solve_s'])
PY
