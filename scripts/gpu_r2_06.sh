#!/bin/bash
# round 2, visit 6 (1 GPU): p129, p122, tetrahedron rules, matrix-free dispatch, pcg_km; compute-sanitizer on the kernels
# added this round; the reference arm on the GPU box's host cores at the GPU arm's own size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dynamic.py tests/test_gpu_plastic.py tests/test_gpu_tetrahedra.py tests/test_gpu_matrix_free.py tests/test_gpu_parity.py -q --durations=6 > gpurun_out/r2_06_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_06_pytest.log
tail -30 gpurun_out/r2_06_pytest.log
cat > /tmp/san3.py <<'PY'
import sys, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from parafem_b200 import host, solver, driver
with solver.Solver(0, 1, 0) as s:
    # p122: fixed-freedom branch on a small box, loaded branch on hex20
    p = host.cube_p121(4, 4, 3, 8, aa=1., bb=1., cc=1., e=100.0, v=0.3)
    p.program, p.phi, p.c, p.psi = 122, 20.0, 4.0, 0.0
    p.qinc, p.plasits, p.cjits, p.plastol, p.cjtol, p.loaded_nodes = [0.5, 0.3], 12, 60, 1e-4, 1e-6, 1
    print("p122 hex8", [(r[4], r[5]) for r in driver.run_p122(p, s)["rows"]])
    p = host.cube_p121(5, 5, 3, 20, aa=2., bb=2., cc=2., e=100.0, v=0.3)
    p.program, p.phi, p.c, p.psi = 122, 20.0, 4.0, 0.0
    p.qinc, p.plasits, p.cjits, p.plastol, p.cjtol, p.loaded_nodes = [0.5, 0.3], 10, 60, 1e-4, 1e-6, 1
    print("p122 hex20", [(r[4], r[5]) for r in driver.run_p122(p, s)["rows"]])
    # p129: 27-point rule and 8-point rule
    for nip in (27, 8):
        p = host.cube_p129(2, 4, 2, .25, .25, .25, e=1.0e4, nip=nip, nstep=2, limit=40)
        print("p129", nip, [r[3] for r in driver.run_p129(p, s)["rows"]])
    # pcg_km on three element types; tetrahedron rules; two-lane matrix-free kernel (mode 2) and mode 1
    for q in (host.cube_p121(5, 4, 3, 20, aa=1., bb=1., cc=1., limit=25), host.cube_p121(6, 5, 4, 8, aa=1., bb=1., cc=1., limit=25),
              host.cube_p123(7, 6, 5, limit=25)):
        solver.setup_problem(s, q)
        km = s.get_storkm(0, 1)[0]
        dg = s.diag_precon()
        s.setup_mesh(q)
        x, it, cv = s.pcg_km(km, dg, q.r_pp, q.tol, q.limit)
        print("pcg_km", q.ntot, it, float(np.abs(x).max()))
    from tet_util import tet_problem
    for nip in (4, 5):
        q = tet_problem(host.cube_p121(3, 3, 3, 8, aa=1., bb=1., cc=1., limit=30)); q.nip = nip
        solver.setup_problem(s, q)
        print("tet", nip, s.pcg_solve(q.r_pp, q.tol, q.limit)[1])
    for nod in (20, 8):
        for mode in (2, 1):
            q = host.cube_p121(5, 4, 3, nod, aa=1., bb=1., cc=1., limit=20)
            solver.setup_problem(s, q, matrix_free=mode)
            print("mf", nod, mode, s.pcg_solve(q.r_pp, q.tol, q.limit)[1])
PY
for tool in memcheck racecheck "synccheck --num-cuda-barriers 65536"; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san3.py > gpurun_out/r2_06_sanitizer_${tool%% *}.log 2>&1
  tail -5 gpurun_out/r2_06_sanitizer_${tool%% *}.log
done
timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_06_bench_reference.json 2> gpurun_out/r2_06_bench_reference.err
tail -1 gpurun_out/r2_06_bench_reference.json | cut -c1-900
