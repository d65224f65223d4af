#!/bin/bash
# round 2, visit 17 (8 GPUs): N-rank parity at 8 in both transports (incl. the tensor-core matrix-free kernel and p1210),
# the default bench line at N = 8 (parity_check, config D, weak, variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q --durations=3 -k "8-peer or 8-nccl" > gpurun_out/r2_17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_17_pytest.log
tail -6 gpurun_out/r2_17_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_17_bench_g8.json 2> gpurun_out/r2_17_bench_g8.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2_17_bench_g8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_17_bench_g8.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['parity_check']['bit_equal'], d['parity_check']['transport'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), v['kernel_ms_per_step'], round(v['roofline']['frac'],4))
for k,v in d['variants'].items(): print(k, round(v['value'],1), v['ms_per_step'], v['kernel_ms_per_step'], round(v['roofline']['frac'],4), (v.get('time_to_solution') or {}).get('solve_s'))
print('weak', d['weak']['value'], 'tts', d['time_to_solution'])
PY
timeout 600 $TR --master-port 29812 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-variants --no-solve --matrix-free 2 > gpurun_out/r2_17_bench_g8_mf2.json 2> gpurun_out/r2_17_bench_g8_mf2.err
tail -1 gpurun_out/r2_17_bench_g8_mf2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mf2 200 steps', round(d['value'],1), d['ms_per_step'], d['kernel_ms_per_step'], round(d['roofline']['frac'],4))"
