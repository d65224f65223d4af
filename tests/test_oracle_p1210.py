"""p1210 (explicit elasto-plastic von Mises dynamics, programs/5th_ed/p1210/p1210.f90) on CPU: the deck reader on the
reference's one deck (both .dat layouts) and the oracle against p1210_tiny.dis -- the displacement fields of a
300 000-step run, five digits each."""
import numpy as np
import pytest

import oracle
from parafem_b200 import host
from p1210_util import GOLDEN_PLOAD, equal_to_printed_digits, golden_fields, nodal, synthetic, write_tiny_deck


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    return host.read_deck_p1210(write_tiny_deck(tmp_path_factory.mktemp("p1210")))


def test_deck_reader_takes_both_layouts(tiny, tmp_path):
    p = tiny
    assert (p.nels, p.nn, p.nr, p.nod, p.nip, p.neq) == (5, 68, 8, 20, 8, 180)        # p1210_tiny.res: 180 equations
    assert (p.rho, p.e, p.v, p.sbary, p.dtim, p.nstep, p.npri, p.nres) == (1e-2, 4e4, .3, 350.0, 1e-6, 300000, 3000, 1)
    assert p.pload == -0.003                                                          # as written; see GOLDEN_PLOAD
    assert abs(p.total_load - 10.0) < 1e-4                                            # "Total load applied 0.1000E+02"
    q = host.read_deck_p1210(write_tiny_deck(tmp_path, current_layout=True))
    assert (q.rho, q.e, q.v, q.sbary, q.dtim, q.nstep, q.npri, q.nres, q.pload) == (1e-2, 4e4, .3, 350.0, 1e-6, 300000, 3000, 1, 2.0)
    assert np.array_equal(q.g_g_pp, p.g_g_pp) and np.array_equal(q.r_pp, p.r_pp)


def test_lumped_mass_sums_to_the_mass_of_the_solid(tiny):
    p = tiny
    out = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, GOLDEN_PLOAD, 0, 1)
    mm_tmp = np.zeros((p.nels, 60))
    assert oracle.lib().orc_p1210_mass(oracle.C.c_int64(p.nels), 20, 8, oracle._p(oracle._f64(p.g_coord_pp)), oracle.C.c_double(p.rho),
                                       oracle._p(mm_tmp)) == 0
    # every element: 12 mid-side nodes carry volume*rho/13, 8 corners an eighth of that, in each of 3 directions
    vol = 1.0 * 1.0 * 1.0 * p.rho
    assert np.allclose(mm_tmp.reshape(p.nels, 20, 3).sum(axis=1), vol, rtol=1e-13)
    assert np.all(out["mm"] > 0) and out["mm"].sum() < mm_tmp.sum()                   # restrained freedoms carry no equation


def test_all_golden_fields_of_the_300000_step_run(tiny):
    """Every kept output step of p1210_tiny.dis (elastic start, first yield, plastic cycling, the last step) to the five
    digits printed -- on one emulated rank and on the golden's own four (the partition only orders the scatter's sums)."""
    p = tiny
    gold = golden_fields()
    out = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, GOLDEN_PLOAD, 300000, 3000)
    by_step = {s: x for s, x, _, _ in out["snaps"]}
    assert len(by_step) == 100
    for step, g in gold.items():
        assert equal_to_printed_digits(nodal(p, by_step[step]), g), step
    assert np.abs(gold[300000]).max() > 50 * np.abs(gold[3000]).max()                 # it moved, and it is not elastic scaling:
    lin = nodal(p, by_step[3000]) * (np.abs(gold[300000]).max() / np.abs(gold[3000]).max())
    assert not equal_to_printed_digits(lin, gold[300000], digits=2)
    four = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, GOLDEN_PLOAD, 9000, 3000, npes=4)
    for step, x, _, _ in four["snaps"]:
        assert equal_to_printed_digits(nodal(p, x), gold[step]), step
        assert np.abs(x - by_step[step]).max() <= 1e-12 * np.abs(x).max()


def test_operator_form_is_another_rounding_of_the_same_update(tiny):
    """orc_p1210_elements_mf (what the tensor-core kernel computes) against elements_2 as written, and against the golden
    fields through first yield (the full 300 000 steps were checked once: all 100 fields, 23 s)."""
    p = tiny
    a = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, GOLDEN_PLOAD, 30000, 3000, form=1)
    b = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, GOLDEN_PLOAD, 30000, 3000, form=0)
    gold = golden_fields()
    for (step, x, _, _), (_, y, _, _) in zip(a["snaps"], b["snaps"]):
        assert not np.array_equal(x, y) and np.abs(x - y).max() <= 1e-12 * np.abs(y).max()
        if step in gold:
            assert equal_to_printed_digits(nodal(p, x), gold[step]), step


def test_the_written_load_factor_is_not_the_golden_one(tiny):
    p = tiny
    out = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, 3000, 3000)
    ours, g = nodal(p, out["snaps"][0][1]), golden_fields()[3000]
    # elastic so far: the field is the golden one times pload / 2.0
    assert np.abs(ours - g * (p.pload / GOLDEN_PLOAD)).max() <= 1e-4 * np.abs(ours).max()


def test_synthetic_case_yields():
    p = synthetic(host)
    el = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, 1e30, p.rho, p.dtim, p.pload, p.nstep, p.npri)
    pl = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, p.nstep, p.npri)
    a, b = el["snaps"][-1][1], pl["snaps"][-1][1]
    assert np.all(np.isfinite(b)) and np.abs(a - b).max() > 1e-3 * np.abs(a).max()   # the yield branch changed the answer
