"""Per-element materials and displacement control for the elastic elements (xx2 / xx1 /
rfemsolve reuse p121's three kernels with a non-uniform storkm): pf_form_km_elastic_mat through the C-ABI
against the oracle, and the reference's multi-material golden (examples/dev/xx2/xx2-tiny: 59 iterations,
756 x 3 displacements)."""
import os
import re

import numpy as np
import pytest

import oracle
from parafem_b200 import host, solver
from parafem_b200._lib import PfError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.close()


def with_materials(p, np_types, seed=3):
    rng = np.random.RandomState(seed)
    p.prop = np.column_stack([rng.uniform(50., 5000., np_types), rng.uniform(0.05, 0.45, np_types)])
    p.etype_pp = rng.randint(1, np_types + 1, p.nels_pp).astype(np.int32)
    return p


@pytest.mark.parametrize("nod,layout", [(20, 0), (8, 0), (20, 1), (8, 1)])
def test_material_matrices_and_solve_equal_oracle(gpu, nod, layout):
    """storkm_pp with e, v = prop(:,etype_pp(iel)) (xx2.f90:169-193) == the oracle, bit for bit, and so is
    the whole solve; layout 1 = packed lower triangles against the symmetrised oracle matrices."""
    p = with_materials(host.cube_p121(6, 5, 4, nod, aa=1., bb=.8, cc=1.25, limit=1000, distort=0.15 if nod == 20 else 0.0), 7)
    solver.setup_problem(gpu, p, layout=layout)
    km = oracle.form_km_elastic_mat(p.g_coord_pp, nod, p.nip, p.prop, p.etype_pp)
    if layout == 1:
        km = np.triu(km) + np.triu(km, 1).transpose(0, 2, 1)
    assert np.array_equal(gpu.get_storkm(), km)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1)
    assert (iters, conv) == (ref["iters"], ref["converged"]) and conv
    assert np.array_equal(x, ref["x"])
    one = host.cube_p121(6, 5, 4, nod, aa=1., bb=.8, cc=1.25)
    one.prop, one.etype_pp = np.array([[one.e, one.v]]), np.ones(one.nels_pp, np.int32)
    solver.setup_problem(gpu, one)                 # one material through the table == pf_form_km_elastic
    assert np.array_equal(gpu.get_storkm(), oracle.form_km_elastic(one.g_coord_pp, nod, one.nip, one.e, one.v))


def test_xx2_tiny_deck_golden(gpu, tiny_xx2, golden):
    """The reference's multi-material deck through the device path: 59 iterations (xx2-tiny.res; the blocked
    reduction order gives 58, the oracle in the same order too) and the golden displacements to 5 digits."""
    p = tiny_xx2
    solver.setup_problem(gpu, p)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    gold = int(re.search(r"Number of PCG iterations\s+(\d+)", open(os.path.join(golden, "xx2-tiny.res")).read()).group(1))
    assert conv and gold == 59 and abs(iters - gold) <= 2
    dis = np.loadtxt(os.path.join(golden, "xx2-tiny.dis"), skiprows=2)[:, 1:]
    u = np.zeros((p.nn, 3))
    m = p.nf > 0
    u[m] = x[p.nf[m] - 1]
    assert np.abs(u - dis).max() < 1e-5
    km = oracle.form_km_elastic_mat(p.g_coord_pp, p.nod, p.nip, p.prop, p.etype_pp)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, p.r_pp, p.tol, p.limit, npes=1, red_mode=1)
    assert iters == ref["iters"] and np.array_equal(x, ref["x"])


def test_displacement_control_equals_oracle(gpu):
    """xx1 / xx2 fixed freedoms (xx2.f90:250-300, 322-327): penalty on the preconditioner, r = store*valf,
    u = p*store inside the loop -- on the elastic elements."""
    p = host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., limit=500)
    top = np.unique(p.nf[:30, 2])
    p.no_f = top[top > 0][:6].astype(np.int32)          # z-freedoms of some top-face nodes pushed down by 0.01
    p.val_f = np.full(p.no_f.size, -0.01)
    p.r_pp[:] = 0.0
    solver.setup_problem(gpu, p)
    x, iters, conv = gpu.pcg_solve(p.r_pp, p.tol, p.limit)
    km = oracle.form_km_elastic(p.g_coord_pp, 20, 8, p.e, p.v)
    ref = oracle.pcg(km, p.g_g_pp, p.neq, np.zeros(p.neq), p.tol, p.limit, npes=1, red_mode=1, no_f=p.no_f, val_f=p.val_f)
    assert conv and iters == ref["iters"] and np.array_equal(x, ref["x"])
    assert np.abs(x[p.no_f - 1] + 0.01).max() < 1e-8


def test_material_argument_errors(gpu):
    p = with_materials(host.cube_p121(3, 3, 3, 8, aa=1., bb=1., cc=1.), 3)
    gpu.setup_mesh(p)
    bad = p.etype_pp.copy()
    bad[5] = 4
    with pytest.raises(PfError, match="outside 1..3"):
        gpu.form_km_elastic_mat(p.prop, bad)
    gpu.set_matrix_free(1)
    with pytest.raises(PfError, match="one material"):
        gpu.form_km_elastic_mat(p.prop, p.etype_pp)
    gpu.set_matrix_free(0)
