#!/bin/bash
# round 2, visit 4 (8 GPUs): N-rank parity at 8 (both transports) and 4, bench at N = 8 fused / unfused, matrix-free
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q --durations=5 -k "8-peer or 8-nccl or 4-peer" > gpurun_out/r2_04_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_04_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_04_bench_g8.json 2> gpurun_out/r2_04_bench_g8.err
echo "bench rc=$?" >> gpurun_out/r2_04_bench_g8.err
PF_FUSE=0 timeout 600 $TR --master-port 29712 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve > gpurun_out/r2_04_bench_g8_unfused.json 2> gpurun_out/r2_04_bench_g8_unfused.err
timeout 600 $TR --master-port 29713 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve > gpurun_out/r2_04_bench_g8_fused200.json 2> gpurun_out/r2_04_bench_g8_fused200.err
PF_GRAPH=0 timeout 600 $TR --master-port 29714 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve > gpurun_out/r2_04_bench_g8_nograph.json 2> gpurun_out/r2_04_bench_g8_nograph.err
timeout 600 $TR --master-port 29715 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_04_bench_g8_mf2.json 2> gpurun_out/r2_04_bench_g8_mf2.err
PF_FUSE=0 timeout 600 $TR --master-port 29716 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_04_bench_g8_mf2_unfused.json 2> gpurun_out/r2_04_bench_g8_mf2_unfused.err
tail -8 gpurun_out/r2_04_pytest.log; tail -3 gpurun_out/r2_04_bench_g8.err; for f in gpurun_out/r2_04_bench_g8*.json; do echo $f; tail -1 $f | head -c 250; echo; done
