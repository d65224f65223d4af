#!/bin/bash
# round 2, visit 11 (1 GPU): full GPU suite + smoke with the tensor-core matrix-free kernels as default; compute-sanitizer
# (memcheck, racecheck, synccheck) on k_apply_mf4 / k_apply_mf3 in both modes, both bricks, ragged sizes; default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_11_pytest.log
tail -9 gpurun_out/r2_11_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cat > /tmp/san4.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from parafem_b200 import host, solver
with solver.Solver(0, 1, 0) as s:
    for nod, dims in ((20, (5, 3, 4)), (8, (10, 7, 5)), (20, (4, 4, 4))):
        for mode in (2, 1):
            p = host.cube_p121(*dims, nod, aa=1., bb=2., cc=.5, limit=12)
            solver.setup_problem(s, p, matrix_free=mode)
            pm = np.random.RandomState(2).randn(p.nels, p.ntot)
            ut = s.matvec(pm)
            x, it, cv = s.pcg_solve(p.r_pp, p.tol, p.limit)
            x, it, cv = s.pcg_solve(p.r_pp, p.tol, p.limit)      # second solve: graph replay
            print("mf", nod, dims, mode, it, float(np.abs(ut).max()), float(np.abs(x).max()))
PY
for sel in 4 3; do
for tool in memcheck racecheck "synccheck --num-cuda-barriers 65536"; do
  PF_MF=$sel timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san4.py > gpurun_out/sanitizer4_mf${sel}_${tool%% *}.log 2>&1
  tail -3 gpurun_out/sanitizer4_mf${sel}_${tool%% *}.log
done
done
timeout 900 python bench.py > gpurun_out/r2_11_bench_default.json 2> gpurun_out/r2_11_bench_default.err
tail -c 600 gpurun_out/r2_11_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_11_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_check'])
for k,v in d['variants'].items(): print(k, round(v['value'],1), v['kernel_ms_per_step'], round(v['roofline']['frac'],4), v['roofline'].get('frac_of_dfma_peak'), v.get('time_to_solution'))
PY
