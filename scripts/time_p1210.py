"""p1210 (explicit elasto-plastic dynamics) at scale on one GPU: K explicit steps of an n^3 cube of 20-node bricks,
ms per step and the kernel shares (CUDA events around every launch).  python scripts/time_p1210.py [n] [steps] [form]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parafem_b200 import host, solver   # noqa: E402
from p1210_util import synthetic         # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
form = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = synthetic(host, n, n, n, nstep=K, npri=K)
p.form = form
p.dtim = 2.0e-3 * 4.0 / n                      # inside the stability limit of the finer mesh
with solver.Solver(0, 1, 0) as s:
    solver.setup_problem(s, p)
    s.vm_explicit_steps(5)
    ms = s.vm_explicit_steps(K)
    s.set_profile(True); s.reset_profile()
    s.vm_explicit_steps(K)
    km = {name: s.kernel_ms(i) for i, name in enumerate(("elements", "scatter", "vector", "halo"))}
    s.set_profile(False)
    x1, _, _ = s.vm_explicit_get()
    print(json.dumps({"program": "p1210", "form": form, "workload": f"{n}^3 20-node bricks, {p.nels} elements, {p.neq} equations", "steps": K,
                      "ms_per_step": ms / K, "MDOF_steps_per_s": p.neq * K / (ms / 1e3) / 1e6,
                      "kernel_ms_per_step": {k: v[0] / K for k, v in km.items()}, "max_abs_x1": float(abs(x1).max())}))
