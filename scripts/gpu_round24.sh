#!/bin/bash
# r01 round 24: 4 GPUs -- N-rank parity (2 and 4 ranks, all specs incl. p124 / p125 / materials) and the
# strong-scaling bench line at 4 GPUs after this session's changes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu --no-solve > gpurun_out/r24_scale_g4.json 2> gpurun_out/r24_g4.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r24_scale_g4.json') if l.startswith('{')][-1]); print('C 4 GPUs', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], 'e2e', round(d['e2e']['value'])); print({k: round(v['value']) for k, v in (d.get('variants') or {}).items()})"; tail -3 gpurun_out/r24_g4.err
