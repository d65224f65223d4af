"""Times pf_form_km_elastic (wall clock around the blocking C-ABI call, best of 3) on the BASELINE cubes.
PF_FORM=old selects the first, untiled build of the kernel.  usage: time_form_km.py [n nod]..."""
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parafem_b200 import host, solver  # noqa: E402

args = [int(v) for v in sys.argv[1:]] or [125, 20, 200, 8]
with solver.Solver(0, 1, 0) as s:
    for n, nod in zip(args[0::2], args[1::2]):
        p = host.cube_p121(n, n, n, nod)
        s.setup_mesh(p)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            s.form_km_elastic(p.e, p.v)
            best = min(best, time.perf_counter() - t)
        km = s.get_storkm(0, 2000)
        flops = p.nels * 8 * (nod * nod) * 60          # FP64 instructions of the tiled build per element
        print(f"form_km {os.environ.get('PF_FORM', 'tiled')}: {n}^3 hex{nod} {p.nels} elements {best * 1e3:.1f} ms "
              f"= {p.nels * (3 * nod) ** 2 * 8 / best / 1e9:.0f} GB/s written, {flops / best / 1e12:.2f} T FP64 instr/s; "
              f"crc of the first 2000 matrices {zlib.crc32(km.tobytes()):08x}", flush=True)
