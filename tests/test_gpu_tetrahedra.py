"""4-node tetrahedra through the C-ABI: shape_der nod = 4 / sample('tetrahedron') nip = 1 in the
p121 and p123 element loops, 12x12 and 4x4 element matrices through the same bulk-copy-ring mat-vec, scatter and
PCG kernels.  Bar: bit-equal to the oracle; the patch test and the reference's tetrahedron deck as size-independent
checks."""
import os

import numpy as np
import pytest

import oracle
from parafem_b200 import host, solver
from parafem_b200._lib import PfError
from tet_util import patch_problem, tet_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.close()


def test_scalar_tetrahedra_equal_oracle_and_pass_the_patch_test(gpu):
    p, field = patch_problem()
    solver.setup_problem(gpu, p)
    kc = oracle.form_kc_laplace(p.g_coord_pp, 1, 1., 1., 1.)
    assert np.array_equal(gpu.get_storkm(), kc)
    rng = np.random.RandomState(2)
    pv = rng.randn(p.neq)
    pm = gpu.gather(pv)
    assert np.array_equal(pm, oracle.gather(p.g_g_pp, pv))
    assert np.array_equal(gpu.matvec(pm), oracle.matvec(kc, pm))
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    ref = oracle.pcg(kc, p.g_g_pp, p.neq, np.zeros(p.neq), p.tol, p.limit, npes=1, red_mode=1, no_f=p.no_f, val_f=p.val_f)
    assert conv and iters == ref["iters"] and np.array_equal(x, ref["x"])
    assert np.abs(x - field).max() <= 1e-10 * np.abs(field).max()


@pytest.mark.parametrize("shape", [(4, 3, 3), (7, 5, 2)])
def test_elastic_tetrahedra_equal_oracle(gpu, shape):
    base = host.cube_p121(*shape, 8, aa=1., bb=.8, cc=1.25, limit=3000)
    p = tet_problem(base)
    solver.setup_problem(gpu, p)
    km = oracle.form_km_elastic(p.g_coord_pp, 4, 1, p.e, p.v)
    assert np.array_equal(gpu.get_storkm(), km)
    rng = np.random.RandomState(3)
    pv = rng.randn(p.neq)
    u_ref = oracle.scatter(p.g_g_pp, oracle.matvec(km, oracle.gather(p.g_g_pp, pv)), p.neq)
    assert np.array_equal(gpu.apply(pv), u_ref)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1)
    assert conv and iters == ref["iters"] and np.array_equal(x, ref["x"])
    assert np.array_equal(gpu.ratio_history(), ref["ratio"])
    sig = gpu.centroid_stress(0, p.e, p.v)
    g0 = p.g_g_pp[0]
    eld = np.where(g0 > 0, x[np.maximum(g0, 1) - 1], 0.0)
    assert np.allclose(sig, oracle.centroid_stress(4, p.g_coord_pp[0], eld, p.e, p.v), rtol=0, atol=1e-12)
    # per-element materials on tetrahedra
    p.prop = np.array([[50., .2], [500., .35], [5000., .1]])
    p.etype_pp = (np.arange(p.nels_pp) % 3 + 1).astype(np.int32)
    solver.setup_problem(gpu, p)
    assert np.array_equal(gpu.get_storkm(), oracle.form_km_elastic_mat(p.g_coord_pp, 4, 1, p.prop, p.etype_pp))


def test_xx11_tetrahedron_deck(gpu, golden):
    """The reference's tetrahedron deck (no shipped output): device == oracle, and the loaded face's mean temperature
    within 2 % of the brick deck's golden."""
    p = host.read_deck_p123(os.path.join(golden, "xx11_tetcube"))
    r0 = p.r_pp.copy()
    solver.setup_problem(gpu, p)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    kc = oracle.form_kc_laplace(p.g_coord_pp, p.nip, p.kx, p.ky, p.kz)
    ref = oracle.pcg(kc, p.g_g_pp, p.neq, r0, p.tol, p.limit, npes=1, red_mode=1, no_f=p.no_f, val_f=p.val_f)
    assert conv and iters == ref["iters"] and np.array_equal(x, ref["x"])
    gold = np.loadtxt(os.path.join(golden, "xx11.ttr"), skiprows=2)[:, 1]
    brick = host.read_deck_p123(os.path.join(golden, "xx11"))
    assert abs(x[np.flatnonzero(r0)].mean() - gold[np.flatnonzero(brick.r_pp)].mean()) < 0.02 * gold.max()


def test_tetrahedra_limits_are_reported(gpu):
    p = tet_problem(host.cube_p121(3, 3, 3, 8, aa=1., bb=1., cc=1.))
    gpu.setup_mesh(p)
    gpu.set_matrix_free(1)
    with pytest.raises(PfError, match="tetrahedra"):
        gpu.form_km_elastic(p.e, p.v)
    gpu.set_matrix_free(0)
    gpu.set_storkm_layout(1)
    with pytest.raises(PfError, match="tetrahedra"):
        gpu.form_km_elastic(p.e, p.v)
    gpu.set_storkm_layout(0)
    bad = tet_problem(host.cube_p121(3, 3, 3, 8, aa=1., bb=1., cc=1.))
    bad.nip = 8
    with pytest.raises(PfError, match="nip = 1, 4 or 5"):
        gpu.setup_mesh(bad)


@pytest.mark.parametrize("nip", [4, 5])
def test_tetrahedron_four_and_five_point_rules(gpu, nip):
    """sample('tetrahedron') nip = 4 and 5 (new_library.f90:1343-1378; written there with single-precision literals,
    restated as floats widened to double): element matrices bit-equal to the oracle for the elastic and the scalar
    element, and -- the derivatives of a 4-node tetrahedron being constant -- equal to the one-point rule's to the
    single-precision accuracy of the rule's weights; the solve converges to the same field."""
    for base in (host.cube_p121(4, 3, 3, 8, aa=1., bb=1., cc=1., limit=3000), host.cube_p123(5, 4, 4, limit=800)):
        p1 = tet_problem(base)
        p = tet_problem(base)
        p.nip = nip
        solver.setup_problem(gpu, p)
        km = gpu.get_storkm()
        ref = (oracle.form_km_elastic(p.g_coord_pp, 4, nip, p.e, p.v) if p.program == 121
               else oracle.form_kc_laplace(p.g_coord_pp, nip, p.kx, p.ky, p.kz))
        assert np.array_equal(km, ref)
        x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
        solver.setup_problem(gpu, p1)
        km1 = gpu.get_storkm()
        x1, it1, conv1 = gpu.pcg_solve(p1.r_pp, p1.tol, p1.limit)
        assert conv and conv1 and abs(iters - it1) <= 1
        assert np.abs(km - km1).max() <= 2e-7 * np.abs(km1).max()
        assert np.linalg.norm(x - x1) <= 1e-5 * np.linalg.norm(x1)
