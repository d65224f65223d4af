"""The packed-lower-triangle storkm layout (pf_set_storkm_layout(h, 1)): half the storkm stream.

The kernel keeps the reference's summation order (p121.f90:93-97: u_i = sum_j K(i,j) p_j, j ascending,
separate multiply and add) with K(i,j) = L(max(i,j), min(i,j)), so the oracle for it is the ordinary
oracle fed the symmetrised element matrices: bit-exact.  Against the reference's own (unsymmetrised)
storkm_pp the difference is the rounding asymmetry of BtDB (1e-16 relative): iteration count +-1 and,
driven to tol 1e-13, the same field within north_star's 1e-9 relative L2."""
import numpy as np
import pytest

import oracle
from parafem_b200 import host, solver

pytestmark = pytest.mark.gpu
TOL_L2 = 1e-9


def symmetrise(km):
    """K(i,j) := L(max,min).  km is (nels, ntot, ntot) with km[e, j, i] = K(i,j) (Fortran columns)."""
    up = np.triu(km)                       # entries with i >= j: the lower triangle in (i,j) terms
    return up + np.triu(km, 1).transpose(0, 2, 1)


def km_oracle(p):
    if p.program == 121:
        return oracle.form_km_elastic(p.g_coord_pp, p.nod, p.nip, p.e, p.v)
    return oracle.form_kc_laplace(p.g_coord_pp, p.nip, p.kx, p.ky, p.kz)


PROBLEMS = {
    "tiny_hex20": lambda: host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=200),
    "distorted_ragged_hex20": lambda: host.cube_p121(5, 3, 4, 20, aa=1., bb=2., cc=.5, limit=400, distort=0.2),
    "hex8_elastic_odd_count": lambda: host.cube_p121(9, 7, 5, 8, aa=1., bb=1., cc=1., limit=400, distort=0.1),
    "p123_box": lambda: host.cube_p123(12, 9, 10, limit=500),
    "p123_tile_tail": lambda: host.cube_p123(5, 5, 5, limit=200),
}


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.set_storkm_layout(0)
    s.close()


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_packed_matrices_products_and_solve_equal_oracle_on_symmetrised_km(gpu, name):
    p = PROBLEMS[name]()
    solver.setup_problem(gpu, p, layout=1)
    ks = symmetrise(km_oracle(p))
    assert np.array_equal(gpu.get_storkm(), ks)                    # formed on the device, lower triangle kept
    rng = np.random.RandomState(11)
    pm = rng.randn(p.nels, p.ntot)
    assert np.array_equal(gpu.matvec(pm), oracle.matvec(ks, pm))   # same order as MATMUL, K(i,j)=L(max,min)
    pv = rng.randn(p.neq)
    assert np.array_equal(gpu.apply(pv), oracle.scatter(p.g_g_pp, oracle.matvec(ks, oracle.gather(p.g_g_pp, pv)), p.neq))
    kw = dict(no_f=p.no_f, val_f=p.val_f) if p.program == 123 and p.no_f.size else {}
    ref = oracle.pcg(ks, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1, **kw)
    assert np.array_equal(gpu.diag_precon(), ref["diag"])
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    assert conv and iters == ref["iters"]
    assert np.array_equal(x, ref["x"])
    # against the reference's unsymmetrised storkm_pp: count +-1, field within the stopping tolerance
    full = oracle.pcg(km_oracle(p), p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1, **kw)
    assert abs(iters - full["iters"]) <= 1
    assert np.linalg.norm(x - full["x"]) <= 1e-4 * np.linalg.norm(full["x"])


def test_uploaded_storkm_is_packed_and_read_back_symmetrised(gpu):
    p = PROBLEMS["distorted_ragged_hex20"]()
    km = km_oracle(p)
    gpu.setup_mesh(p)
    gpu.set_matrix_free(False)
    gpu.set_storkm_layout(1)
    gpu.set_storkm(km)                                             # xx3-style upload of a host storkm_pp
    assert np.array_equal(gpu.get_storkm(), symmetrise(km))
    assert np.array_equal(gpu.get_storkm(3, 5), symmetrise(km)[3:8])
    gpu.build_precon()
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    ref = oracle.pcg(symmetrise(km), p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1)
    assert conv and iters == ref["iters"] and np.array_equal(x, ref["x"])


def test_tight_tolerance_matches_reference_layout(gpu):
    """tol 1e-13: packed and reference layouts land on the same field within 1e-9 relative L2."""
    p = PROBLEMS["tiny_hex20"]()
    solver.setup_problem(gpu, p, layout=1)
    x_sym, _, c1 = gpu.pcg_solve(p.r_pp, 1e-13, 2000)
    solver.setup_problem(gpu, p, layout=0)
    x_ref, _, c2 = gpu.pcg_solve(p.r_pp, 1e-13, 2000)
    assert c1 and c2
    assert np.linalg.norm(x_sym - x_ref) <= TOL_L2 * np.linalg.norm(x_ref)
    # and the reference layout is still the reference's storkm_pp, bit for bit
    assert np.array_equal(gpu.get_storkm(), km_oracle(p))


def test_asymmetry_of_the_reference_matrices_is_rounding_only():
    """How far storkm_pp is from symmetric (what the packed layout discards): <= a few ulp."""
    p = PROBLEMS["distorted_ragged_hex20"]()
    km = km_oracle(p)
    asym = np.abs(km - km.transpose(0, 2, 1)).max()
    assert asym <= 1e-13 * np.abs(km).max()
