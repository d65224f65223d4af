"""Program p125 (explicit transient conduction) through the C-ABI: the gather / mat-vec /
scatter kernels driven 5000 times without a solver.  Bar: bit-equal to the oracle's restatement of p125.f90 and
the reference's golden log / nodal files of examples/5th_ed/p125/demo."""
import os
import re

import numpy as np
import pytest

import oracle
from parafem_b200 import driver, host, solver

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.close()


@pytest.mark.parametrize("shape,nip", [((7, 5, 6), 8), ((3, 3, 3), 8), ((5, 4, 4), 1)])
def test_explicit_matrices_and_recursion_equal_oracle(gpu, shape, nip):
    """store_pm_pp, globma_pp (p125.f90:66-82) and the field after 1, 2, 10 and 60 steps (:94-99), bit for bit."""
    p = host.cube_p125(*shape, aa=.2, bb=.25, cc=.1, kx=1.5, ky=2.0, kz=0.5, dtim=1e-4, nstep=60, nip=nip)
    solver.setup_problem(gpu, p)
    store, mass = oracle.form_k_explicit(p.g_coord_pp, nip, p.kx, p.ky, p.kz, p.dtim)
    assert np.array_equal(gpu.get_storkm(), store)
    ref = oracle.p125(store, mass, p.g_g_pp, p.neq, p.val0, 60, npes=1, keep=(1, 2, 10, 60))
    assert np.array_equal(gpu.diag_precon(), ref["globma"])
    gpu.explicit_start(p.val0)
    assert np.array_equal(gpu.pcg_get_x(), np.full(p.neq, p.val0))
    done = 0
    for j in (1, 2, 10, 60):
        ms = gpu.explicit_steps(j - done)
        done = j
        assert np.array_equal(gpu.pcg_get_x(), ref["fields"][j]), j
        assert ms > 0.0
    assert gpu.explicit_steps(0) >= 0.0 and np.array_equal(gpu.pcg_get_x(), ref["fields"][60])


def test_p125_demo_golden_log_and_fields(gpu, golden, tmp_path):
    """examples/5th_ed/p125/demo: the eleven '  Time  Pressure' lines of p125_demo.res reproduced as TEXT and
    the nodal pressure files of steps 500 and 5000 to the 5 digits they print."""
    p = host.cube_p125(25, 25, 25, aa=.04, bb=.04, cc=.04, round_mode=1)
    base = str(tmp_path / "p125_demo")
    res = driver.run_p125(p, gpu, out_base=base)
    driver.write_res_p125(base + ".res", p, res)
    out = open(base + ".res").read().splitlines()
    gold = open(os.path.join(golden, "p125_demo.res")).read().splitlines()
    assert gold[1] in out                                   # There are 17576 nodes 1951 restrained and 15625 equations
    rows = [l for l in gold if re.match(r"^\s+0\.\d+E[+-]\d+\s+0\.\d+E[+-]\d+\s*$", l)]
    assert len(rows) == 11
    for line in rows:
        assert line in out, (line, out)
    arr = np.load(os.path.join(GOLD, "arrays.npz"))
    for j in (500, 5000):
        field = np.loadtxt(f"{base}.ensi.NDPRE-{j:06d}", skiprows=4)
        g = arr[f"p125_ndpre_{j:04d}"].astype(np.float64)
        assert np.abs(field - g).max() <= 1.2e-5 * np.abs(g).max()


def test_explicit_properties_at_64k_elements(gpu):
    """Size-independent properties at 64 000 elements: linear in val0 (exact for a power of two) and inside
    [0, val0] (forward Euler is inside its stability limit at the shipped dtim; oracle: min 0.0348, max 99.763)."""
    p = host.cube_p125(40, 40, 40, nstep=200)
    solver.setup_problem(gpu, p)
    gpu.explicit_start(100.0)
    gpu.explicit_steps(200)
    a = gpu.pcg_get_x()
    gpu.explicit_start(400.0)
    gpu.explicit_steps(200)
    assert np.array_equal(gpu.pcg_get_x(), 4.0 * a)
    assert 0.0 < a.min() < 0.05 and 99.7 < a.max() < 100.0
