#!/bin/bash
# N-GPU visit: N-rank parity in both transport modes, then bench nccl vs peer
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -15
PF_HALO=nccl timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
for mode in peer nccl; do
  PF_HALO=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 200 --warmup 5 --no-cpu --no-solve > gpurun_out/peer_${mode}_g$N.json 2> gpurun_out/peer_${mode}_g$N.err
  grep '^{' gpurun_out/peer_${mode}_g$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', d['n_gpus'], round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['gpu_launches'])"; tail -3 gpurun_out/peer_${mode}_g$N.err
done
