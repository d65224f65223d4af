#!/bin/bash
# compute-sanitizer on the smoke problem (the reference has no race detection at all)
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from parafem_b200 import host, solver
for prob, mf, lay in ((host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=30), 0, 0),
                      (host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=30), 1, 0),
                      (host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=30), 2, 0),
                      (host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=30), 0, 1),
                      (host.cube_p121(6, 5, 4, 8, aa=1., bb=1., cc=1., limit=30), 0, 0),
                      (host.cube_p121(6, 5, 4, 8, aa=1., bb=1., cc=1., limit=30), 2, 0),
                      (host.cube_p121(6, 5, 4, 8, aa=1., bb=1., cc=1., limit=30), 0, 1),
                      (host.cube_p123(7, 6, 5, limit=30), 0, 0),
                      (host.cube_p123(7, 6, 5, limit=30), 0, 1)):
    with solver.Solver(0, 1, 0) as s:
        solver.setup_problem(s, prob, matrix_free=mf, layout=lay)
        x, it, cv = s.pcg_solve(prob.r_pp, prob.tol, prob.limit)
        s.apply(np.ones(prob.neq)); s.dot(x, x)
        print(prob.program, prob.nod, mf, lay, it, cv, float(np.abs(x).max()))
PY
for tool in memcheck racecheck "synccheck --num-cuda-barriers 65536"; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_${tool%% *}.log 2>&1
  tail -12 gpurun_out/sanitizer_${tool%% *}.log
done
