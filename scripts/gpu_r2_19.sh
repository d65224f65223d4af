#!/bin/bash
# round 2, visit 19 (1 GPU): full GPU suite + smoke on the round's last build; compute-sanitizer (memcheck, racecheck,
# synccheck) on the kernels added since visit 11: k_matvec2 (p123 / p124 / p125), p1210 in both forms, the tensor-core
# matrix-free kernels on meshes smaller than one warp pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_19_pytest.log
tail -9 gpurun_out/r2_19_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cat > /tmp/san5.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from parafem_b200 import host, solver, driver
from p1210_util import synthetic
with solver.Solver(0, 1, 0) as s:
    for q in (host.cube_p123(7, 6, 5, limit=25), host.cube_p123(3, 2, 2, limit=25)):
        solver.setup_problem(s, q)
        x, it, cv = s.pcg_solve(q.r_pp, q.tol, q.limit)
        x, it, cv = s.pcg_solve(q.r_pp, q.tol, q.limit)
        print("p123", q.nels, it, float(np.abs(x).max()))
    p = host.cube_p125(6, 5, 4, nstep=10)
    solver.setup_problem(s, p); s.explicit_start(p.val0); s.explicit_steps(10)
    print("p125", float(s.pcg_get_x().max()))
    for form in (0, 1):
        for shape in ((4, 5, 3), (1, 1, 2)):
            q = synthetic(host, *shape, nstep=12, npri=6)
            q.form = form
            res = driver.run_p1210(q, s)
            print("p1210", form, shape, float(np.abs(res["x"]).max()))
    for dims, nod in (((1, 1, 2), 20), ((2, 1, 1), 8)):
        for mode in (1, 2):
            q = host.cube_p121(*dims, nod, aa=1., bb=.5, cc=2., limit=10)
            solver.setup_problem(s, q, matrix_free=mode)
            x, it, cv = s.pcg_solve(q.r_pp, q.tol, q.limit)
            print("mf tiny", dims, nod, mode, it)
PY
for tool in memcheck racecheck "synccheck --num-cuda-barriers 65536"; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python /tmp/san5.py > gpurun_out/sanitizer5_${tool%% *}.log 2>&1
  tail -2 gpurun_out/sanitizer5_${tool%% *}.log
done
