#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
for cfg in "--cube 125 --hex 20" "--cube 200 --hex 8"; do
for mode in peer nccl; do
  tag=$(echo $cfg | tr -d ' -')
  PF_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N $cfg --steps 200 --warmup 5 --no-cpu > gpurun_out/p8_${mode}_${tag}_g$N.json 2> gpurun_out/p8_${mode}_${tag}_g$N.err
  grep '^{' gpurun_out/p8_${mode}_${tag}_g$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', '$cfg', d['n_gpus'], round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['gpu_launches'], d['time_to_solution'])"; tail -2 gpurun_out/p8_${mode}_${tag}_g$N.err | cut -c1-300
done
done
PF_HALO=peer timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
