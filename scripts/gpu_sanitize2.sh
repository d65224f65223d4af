#!/bin/bash
# compute-sanitizer on the kernels added after the first sanitizer visit: tiled km formation (+ materials),
# p124 / p125 element matrices and time stepping, the xx3-compatible mat-vec, graph replay of the iteration
set -x
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import sys; sys.path.insert(0, '.')
import ctypes as C
import numpy as np
from parafem_b200 import host, solver
from parafem_b200._lib import lib, ptr
with solver.Solver(0, 1, 0) as s:
    for nod, lay in ((20, 0), (8, 0), (20, 1)):
        p = host.cube_p121(5, 4, 3, nod, aa=1., bb=1., cc=1., limit=25)
        rng = np.random.RandomState(1)
        p.prop = np.column_stack([rng.uniform(50., 500., 3), rng.uniform(.1, .4, 3)])
        p.etype_pp = rng.randint(1, 4, p.nels).astype(np.int32)
        solver.setup_problem(s, p, layout=lay)
        x, it, cv = s.pcg_solve(p.r_pp, p.tol, p.limit)
        print("materials", nod, lay, it, float(np.abs(x).max()))
        q = host.cube_p121(5, 4, 3, nod, aa=1., bb=1., cc=1., limit=25)
        solver.setup_problem(s, q, layout=lay)
        x, it, cv = s.pcg_solve(q.r_pp, q.tol, q.limit)
        x, it, cv = s.pcg_solve(q.r_pp, q.tol, q.limit)          # second solve: graph replay
        print("tiled", nod, lay, it, float(np.abs(x).max()))
    for fixed in (False, True):
        p = host.cube_p124(6, 5, 4, nstep=4, fixed=fixed)
        solver.setup_problem(s, p)
        s.transient_start(p.val0, p.val_f if fixed else None)
        loads = np.zeros(p.neq); loads[p.nres - 1] = 0.1
        its = [s.transient_step(p.tol, p.limit, loads)[0] for _ in range(p.nstep)]
        print("p124", fixed, its, float(s.pcg_get_x().max()))
    p = host.cube_p125(6, 5, 4, nstep=20)
    solver.setup_problem(s, p)
    s.explicit_start(p.val0); s.explicit_steps(20)
    print("p125", float(s.pcg_get_x().max()), s.sum(s.pcg_get_x()))
L = lib()
ci = lambda v: C.byref(C.c_int(v))
for n_mat, nr, nc in ((37, 60, 60), (41, 24, 24), (67, 8, 8), (19, 12, 7)):
    dk, dr, dl = C.c_void_p(), C.c_void_p(), C.c_void_p()
    km, pm, out = np.random.rand(n_mat, nc, nr), np.random.rand(n_mat, nc), np.empty((n_mat, nr))
    assert L.allocate_memory_on_gpu(ci(km.size), ci(8), C.byref(dk)) == 0
    assert L.allocate_memory_on_gpu(ci(pm.size), ci(8), C.byref(dl)) == 0
    assert L.allocate_memory_on_gpu(ci(out.size), ci(8), C.byref(dr)) == 0
    assert L.copy_data_to_gpu(ci(km.size), ci(8), ptr(km), C.byref(dk)) == 0
    assert L.copy_data_to_gpu(ci(pm.size), ci(8), ptr(pm), C.byref(dl)) == 0
    assert L.matrix_vector_multiplies(ci(n_mat), ci(nr), ci(nc), C.byref(dl), C.byref(dk), C.byref(dr)) == 0
    assert L.copy_data_from_gpu(ci(out.size), ci(8), ptr(out), C.byref(dr)) == 0
    for d in (dk, dr, dl):
        assert L.free_memory_on_gpu(C.byref(d)) == 0
    print("xx3", n_mat, nr, nc, float(out.sum()))
PY
for tool in memcheck racecheck "synccheck --num-cuda-barriers 65536"; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san2.py > gpurun_out/sanitizer2_${tool%% *}.log 2>&1
  tail -6 gpurun_out/sanitizer2_${tool%% *}.log
done
