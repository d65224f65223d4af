#!/bin/bash
# N-GPU box visit: N-rank parity against the oracle, then the strong-scaling bench
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -25
for n in $(seq 1 8); do
  if [ $n -le $N ] && { [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; }; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu > gpurun_out/scale_n125_g$n.json 2> gpurun_out/scale_g$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 100 --warmup 5 --no-cpu > gpurun_out/scale_n125_g$n.json 2> gpurun_out/scale_g$n.err
    fi
    tail -c 2500 gpurun_out/scale_n125_g$n.json; tail -4 gpurun_out/scale_g$n.err
  fi
done
