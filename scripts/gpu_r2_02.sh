#!/bin/bash
# round 2, visit 2 (2 GPUs): N-rank parity in both transports, bench at N = 2 fused / unfused
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multirank.py -x -q --durations=5 > gpurun_out/r2_02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_02_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_02_bench_g2.json 2> gpurun_out/r2_02_bench_g2.err
echo "bench rc=$?" >> gpurun_out/r2_02_bench_g2.err
PF_FUSE=0 timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-variants --no-solve > gpurun_out/r2_02_bench_g2_unfused.json 2> gpurun_out/r2_02_bench_g2_unfused.err
PF_GRAPH=0 timeout 600 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-variants --no-solve > gpurun_out/r2_02_bench_g2_nograph.json 2> gpurun_out/r2_02_bench_g2_nograph.err
timeout 600 $TR --master-port 29614 bench.py --gpus 2 --steps 200 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_02_bench_g2_mf2.json 2> gpurun_out/r2_02_bench_g2_mf2.err
PF_FUSE=0 timeout 600 $TR --master-port 29615 bench.py --gpus 2 --steps 200 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_02_bench_g2_mf2_unfused.json 2> gpurun_out/r2_02_bench_g2_mf2_unfused.err
tail -8 gpurun_out/r2_02_pytest.log; tail -3 gpurun_out/r2_02_bench_g2.err; for f in gpurun_out/r2_02_bench_g2*.json; do echo $f; head -c 300 $f; echo; done
