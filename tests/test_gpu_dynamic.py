"""Program p129 (forced vibration of an elastic solid: implicit theta method, consistent mass, Rayleigh damping; one PCG
solve per time step on store_mm*c3 + store_km*c4) on the device, through the C-ABI (pf_form_dynamic / pf_dynamic_start /
pf_dynamic_step / pf_dynamic_get), against the oracle.  The reference ships a deck for this program
(examples/5th_ed/p129/p129_tiny.*, reproduced by the in-memory generators: tests/test_oracle_mesh.py) but no output,
so the pin is GPU == oracle, bit for bit: nothing transcendental runs on the device (the cosines of the harmonic
load are formed on the host by the same libm the oracle uses)."""
import numpy as np
import pytest

import oracle
from parafem_b200 import driver, host, solver

pytestmark = pytest.mark.gpu


def oracle_run(m, nstep, nip):
    km = oracle.form_km_elastic(m.g_coord_pp, 20, nip, m.e, m.v)
    mm = oracle.form_mass(m.g_coord_pp, 20, nip, m.rho)
    return km, mm, oracle.p129(km, mm, m.g_g_pp, m.neq, m.r_pp, m.theta, m.omega, m.alpha1, m.beta1, nstep, m.tol, m.limit,
                               npes=1, red_mode=1, keep=(1, nstep))


@pytest.mark.parametrize("nip,theta", [(27, 1.0), (8, 0.5)])
def test_p129_steps_equal_oracle(nip, theta):
    """A short cantilever (3 x 6 x 2 bricks) with the deck's material and load data: the three matrix sets, the
    preconditioner and eight time steps (iteration counts, displacement, velocity, acceleration) equal the oracle's."""
    nstep = 8
    p = host.cube_p129(3, 6, 2, .25, .25, .25, e=1.0e4, theta=theta, nip=nip, nstep=nstep)
    m = oracle.cube_p129(3, 6, 2, .25, .25, .25, e=1.0e4, theta=theta, nip=nip, nstep=nstep)
    assert np.array_equal(p.g_g_pp, m.g_g_pp) and np.array_equal(p.r_pp, m.r_pp) and p.nres == m.nres
    km, mm, ref = oracle_run(m, nstep, nip)
    with solver.Solver(0, 1, 0) as s:
        out = driver.run_p129(p, s)
        c1 = (1.0 - theta) * p.dtim
        c3, c4 = p.alpha1 + 1.0 / (theta * p.dtim), p.beta1 + theta * p.dtim
        assert np.array_equal(s.get_storkm(0, 4), (mm * c3 + km * c4)[:4])          # the PCG matrix, element by element
        assert np.array_equal(s.get_storkb(0, 4), (km * (p.beta1 - c1) + mm * c3)[:4])
    assert p.dtim == ref["dtim"]
    assert [r[3] for r in out["rows"]] == [r[2] for r in ref["rows"]]
    assert [r[:2] for r in out["rows"]] == [r[:2] for r in ref["rows"]]
    assert np.array_equal(out["x"], ref["x"]) and np.array_equal(out["d1x"], ref["d1x"]) and np.array_equal(out["d2x"], ref["d2x"])
    assert out["rows"][-1][2] == ref["x"][p.nres - 1] and abs(out["rows"][-1][2]) > 0


def test_p129_shipped_deck_size_runs_and_writes_res(tmp_path):
    """The shipped deck's mesh (8 x 40 x 8 bricks, 36 720 equations, the 27-point rule) from the in-memory generator with
    the deck's control data: two steps (422 + 418 PCG iterations) equal the oracle's, and the .res lines have the reference's layout."""
    p = host.cube_p129(8, 40, 8, .125, .125, .125, e=1.0e4, nstep=2)
    m = oracle.cube_p129(8, 40, 8, .125, .125, .125, e=1.0e4, nstep=2)
    oracle.use_all_cores()
    km, mm, ref = oracle_run(m, 2, 27)
    with solver.Solver(0, 1, 0) as s:
        out = driver.run_p129(p, s, out_base=str(tmp_path / "p129_tiny"))
    assert (p.neq, p.nres) == (36720, 36072)
    assert [r[3] for r in out["rows"]] == [r[2] for r in ref["rows"]] == [422, 418] and np.array_equal(out["x"], ref["x"])
    driver.write_res_p129(str(tmp_path / "p129_tiny.res"), p, out)
    lines = open(tmp_path / "p129_tiny.res").read().splitlines()
    assert lines[3] == "   Time t  cos(omega*t) Displacement Iterations" and len(lines) == 4 + 2 + 1
    t, c, x, it = lines[4].split()
    assert float(t) == pytest.approx(p.dtim, rel=1e-3) and int(it) == ref["rows"][0][2]
    assert (tmp_path / "p129_tiny.ensi.DISPL-000002").exists()
