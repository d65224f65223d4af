#!/bin/bash
# r01: 8 x B200, the new kernels at scale (packed storkm layout, matrix-free v4), peer-memory transport
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { # tag args...
  tag=$1; shift 1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N "$@" --steps 200 --warmup 5 --no-cpu --no-variants > gpurun_out/p8b_${tag}_g$N.json 2> gpurun_out/p8b_${tag}_g$N.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/p8b_${tag}_g$N.json') if l.startswith('{')][-1]); print('$tag', d['n_gpus'], round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['gpu_launches'], d['time_to_solution'])"; tail -2 gpurun_out/p8b_${tag}_g$N.err | cut -c1-300
}
run C_hex20_125_full --cube 125 --hex 20
run C_hex20_125_packed --cube 125 --hex 20 --layout 1
run E_mf2_125 --cube 125 --hex 20 --matrix-free 2
run D_hex8_200_packed --cube 200 --hex 8 --layout 1
run E_mf2_hex8_200 --cube 200 --hex 8 --matrix-free 2
