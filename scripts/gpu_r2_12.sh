#!/bin/bash
# round 2, visit 12 (2 GPUs): N-rank parity at 2 in both transports with the tensor-core matrix-free kernel as default
# (its producer warps wait for the peers' forward-halo flags), bench at N = 2: stored and matrix-free modes 2 / 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q --durations=5 -k "2-peer or 2-nccl" > gpurun_out/r2_12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_12_pytest.log
tail -6 gpurun_out/r2_12_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29721 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_12_bench_g2_mf2.json 2> gpurun_out/r2_12_bench_g2_mf2.err
timeout 600 $TR --master-port 29722 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 1 > gpurun_out/r2_12_bench_g2_mf1.json 2> gpurun_out/r2_12_bench_g2_mf1.err
PF_MF=2lane timeout 600 $TR --master-port 29723 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_12_bench_g2_mf2_2lane.json 2> gpurun_out/r2_12_bench_g2_mf2_2lane.err
for f in gpurun_out/r2_12_bench_g2*.json; do echo $f; tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'], d['kernel_ms_per_step'], round(d['roofline']['frac'],4))"; done
