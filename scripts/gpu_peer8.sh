#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { # mode tag args...
  mode=$1; tag=$2; shift 2
  PF_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N "$@" --steps 200 --warmup 5 --no-cpu > gpurun_out/p8_${mode}_${tag}_g$N.json 2> gpurun_out/p8_${mode}_${tag}_g$N.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/p8_${mode}_${tag}_g$N.json') if l.startswith('{')][-1]); print('$mode $tag', d['n_gpus'], round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'], d['gpu_launches'], d['time_to_solution'])"; tail -2 gpurun_out/p8_${mode}_${tag}_g$N.err | cut -c1-300
}
run peer C_hex20_125 --cube 125 --hex 20
run nccl C_hex20_125 --cube 125 --hex 20
run peer D_hex8_200 --cube 200 --hex 8
run nccl D_hex8_200 --cube 200 --hex 8
run peer E_mf2_125 --cube 125 --hex 20 --matrix-free 2
run peer weak_hex20_100 --cube 100 --hex 20 --weak --no-solve
