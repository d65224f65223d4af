"""Wall-clock split of the device setup at config C (or `n nod`): host mesh generation, pf_setup_mesh (tables +
uploads), element matrices, preconditioner."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parafem_b200 import host, solver  # noqa: E402

n, nod = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (125, 20)
t = time.perf_counter()
p = host.cube_p121(n, n, n, nod)
t_mesh = time.perf_counter() - t
with solver.Solver(0, 1, 0) as s:
    out = []
    for rep in range(2):
        t0 = time.perf_counter(); s.setup_mesh(p)
        t1 = time.perf_counter(); s.form_km_elastic(p.e, p.v)
        t2 = time.perf_counter(); s.build_precon()
        t3 = time.perf_counter()
        out.append((t1 - t0, t2 - t1, t3 - t2))
    for rep, (a, b, c) in enumerate(out):
        print(f"setup {n}^3 hex{nod} rep {rep}: host mesh {t_mesh:.3f} s, pf_setup_mesh {a:.3f} s, form_km {b:.3f} s, "
              f"build_precon {c:.3f} s", flush=True)
