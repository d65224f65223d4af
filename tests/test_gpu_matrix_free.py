"""BASELINE config E: the matrix-free variant (km never stored; the element operator
sum_gp B^T D B p det w is recomputed every iteration) against its own oracle mirror
(orc_apply_mf, same fma chains) and against the stored-km path."""
import numpy as np
import pytest

import oracle
from parafem_b200 import PfError, host, solver

pytestmark = pytest.mark.gpu

CASES = {
    "hex20": lambda: host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=200),
    "hex20_distorted_ragged": lambda: host.cube_p121(5, 3, 4, 20, aa=1., bb=2., cc=.5, limit=400, distort=0.2),
    "hex8": lambda: host.cube_p121(10, 7, 5, 8, aa=1., bb=1., cc=1., limit=400),
}


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.close()


@pytest.mark.parametrize("mode", [1, 2])      # 1: rebuild from coordinates; 2: stored geometric factors
@pytest.mark.parametrize("name", list(CASES))
def test_matrix_free_products_equal_oracle(gpu, name, mode):
    p = CASES[name]()
    solver.setup_problem(gpu, p, matrix_free=mode)
    rng = np.random.RandomState(2)
    pm = rng.randn(p.nels, p.ntot)
    ut = gpu.matvec(pm)
    ref = oracle.apply_mf(p.g_coord_pp, p.nod, p.nip, p.e, p.v, pm, mode=mode)
    assert np.array_equal(ut, ref)
    # and it is the same operator as the stored matrices, to rounding
    km = oracle.form_km_elastic(p.g_coord_pp, p.nod, p.nip, p.e, p.v)
    stored = oracle.matvec(km, pm)
    assert np.abs(ut - stored).max() <= 1e-13 * np.abs(stored).max()
    # the preconditioner diagonal is km's diagonal, bit for bit
    r = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, 1.0, 1, npes=1, red_mode=1)
    assert np.array_equal(gpu.diag_precon(), r["diag"])
    with pytest.raises(PfError):
        gpu.get_storkm()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", list(CASES))
def test_matrix_free_pcg_equals_oracle(gpu, name, mode):
    p = CASES[name]()
    solver.setup_problem(gpu, p, matrix_free=mode)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    km = oracle.form_km_elastic(p.g_coord_pp, p.nod, p.nip, p.e, p.v)
    mf = dict(g_coord_pp=p.g_coord_pp, nod=p.nod, nip=p.nip, e=p.e, v=p.v, mode=mode)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1, mf=mf)
    assert conv and iters == ref["iters"]
    assert np.array_equal(x, ref["x"])
    assert np.array_equal(gpu.ratio_history(), ref["ratio"])
    # against the stored-km solve: the same operator in another rounding.  The stopping ratio hovers around tol for a
    # few iterations, so the count may move by one or two (78-80 on the tiny deck between legal
    # summation orders; here 90 against 88 on the distorted mesh with the tensor-core kernel's order); the field is
    # within the stopping tolerance, and within 1e-9 when driven to 1e-13 (next test)
    stored = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1)
    assert abs(iters - stored["iters"]) <= 2
    assert np.linalg.norm(x - stored["x"]) <= 1e-4 * np.linalg.norm(stored["x"])


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("dims,nod", [((1, 1, 2), 20), ((2, 1, 1), 8), ((1, 3, 3), 20)])
def test_fewer_elements_than_one_warp_pass(gpu, dims, nod, mode):
    """2, 2 and 9 elements: the tensor-core kernel's passes of 8 elements are ragged or nearly empty, most consumer warps
    have nothing to do and the producers' index / ring prologues are longer than the work."""
    p = host.cube_p121(*dims, nod, aa=1., bb=.5, cc=2., limit=100)
    solver.setup_problem(gpu, p, matrix_free=mode)
    pm = np.random.RandomState(5).randn(p.nels, p.ntot)
    assert np.array_equal(gpu.matvec(pm), oracle.apply_mf(p.g_coord_pp, p.nod, p.nip, p.e, p.v, pm, mode=mode))
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    km = oracle.form_km_elastic(p.g_coord_pp, p.nod, p.nip, p.e, p.v)
    mf = dict(g_coord_pp=p.g_coord_pp, nod=p.nod, nip=p.nip, e=p.e, v=p.v, mode=mode)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1, mf=mf)
    assert iters == ref["iters"] and np.array_equal(x, ref["x"])


def test_matrix_free_tight_tolerance_matches_stored_path(gpu):
    """Driven to tol 1e-13 the two operator roundings land on the same field within 1e-9."""
    p = CASES["hex20"]()
    solver.setup_problem(gpu, p, matrix_free=True)
    x_mf, _, c1 = gpu.pcg_solve(p.r_pp, 1e-13, 2000)
    solver.setup_problem(gpu, p, matrix_free=False)
    x_st, _, c2 = gpu.pcg_solve(p.r_pp, 1e-13, 2000)
    assert c1 and c2
    assert np.linalg.norm(x_mf - x_st) <= 1e-9 * np.linalg.norm(x_st)


def test_matrix_free_rejects_p123(gpu):
    p = host.cube_p123(5, 5, 5)
    gpu.setup_mesh(p)
    with pytest.raises(PfError):
        gpu.set_matrix_free(True)
    gpu.set_matrix_free(False)
