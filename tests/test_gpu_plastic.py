"""Program p122 (3-D elasto-plasticity: Mohr-Coulomb, viscoplastic strain method, PCG restarted from the current x) on
the device, through the C-ABI (pf_plastic_begin / pf_plastic_increment / pf_plastic_get), against the reference's own
golden log examples/5th_ed/p122/demo/p122_demo.res and displacement field, and against the oracle.

Tolerances, and why: the Gauss-point update evaluates asin / sin / cos / tan of the Lode angle (invar, mocouf, mocouq).
CUDA's double-precision functions are accurate to 1-2 ulp but are not glibc's, so bit equality with the oracle is not
available here as it is for the elastic path; every other operation keeps the oracle's order (orc_p122_elements,
blocked reductions).  Measured: fields agree to ~1e-13 relative; asserted: 1e-9 relative L2 (north_star's bound), equal
iteration counts, golden values to the digits printed."""
import os
import re

import numpy as np
import pytest

import oracle
from parafem_b200 import driver, host, solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def deck(golden, tmp_path_factory):
    """p122_demo.{dat,d,bnd,fix} rebuilt from the packed fixtures (tests/golden/make_golden.py)."""
    a = np.load(os.path.join(os.path.dirname(__file__), "golden", "arrays.npz"))
    d = str(tmp_path_factory.mktemp("p122"))
    base = os.path.join(d, "p122_demo")
    open(base + ".dat", "w").write(open(os.path.join(golden, "p122_demo.dat")).read())
    with open(base + ".d", "w") as f:
        f.write("*THREE_DIMENSIONAL\n*NODES\n")
        for i, c in enumerate(a["p122_coord"]):
            f.write(f"{i + 1}  {float(c[0])!r}  {float(c[1])!r}  {float(c[2])!r}\n")
        f.write("*ELEMENTS\n")
        for e, g in enumerate(a["p122_gnum_sg"]):
            f.write(f"{e + 1}  3  8  1  " + "  ".join(str(int(v)) for v in g) + "  1\n")
    with open(base + ".bnd", "w") as f:
        for row in a["p122_rest"].T:
            f.write(" ".join(str(int(v)) for v in row) + "\n")
    with open(base + ".fix", "w") as f:
        for n, s, v in zip(a["p122_fix_node"], a["p122_fix_sense"], a["p122_fix_val"]):
            f.write(f"{int(n)} {int(s)} {float(v)!r}\n")
    return base, a


def golden_rows(golden):
    res = open(os.path.join(golden, "p122_demo.res")).read()
    d = [float(x) for x in re.findall(r"The displacement is\s+(\S+)", res)]
    s = [[float(x) for x in m] for m in re.findall(r"sigma y\s*\n\s*(\S+)\s+(\S+)\s+(\S+)", res)]
    cj = [int(x) for x in re.findall(r"total number of cj iterations was\s+(\d+)", res)]
    pl = [int(x) for x in re.findall(r"number of plastic iterations was\s+(\d+)", res)]
    return res, d, s, cj, pl


def test_p122_demo_golden_log_and_field(deck, golden, tmp_path):
    """All ten load increments of p122_demo.res: displacement and the three stresses of the first Gauss point to the 4
    digits printed, total cj iterations and plastic iterations EQUAL; final field to the 4 digits of
    p122_demo.ensi.DISPL-000010; the .res file the driver writes carries the golden's lines."""
    base, a = deck
    p = host.read_deck_p122(base)
    res, gold_d, gold_s, gold_cj, gold_pl = golden_rows(golden)
    assert (p.nels, p.nn, p.nr, p.neq, p.no_f.size) == (1152, 1469, 497, 3636, 19) and f"{p.neq} equations" in res
    with solver.Solver(0, 1, 0) as s:
        out = driver.run_p122(p, s, out_base=str(tmp_path / "p122_demo"))
    rows = out["rows"]
    assert len(rows) == len(gold_d) == 10
    assert f"{out['dt']:.3E}" == "5.200E-04"                         # "The critical timestep is     0.5200E-03"
    for (d1, sz, sx, sy, cjtot, plasiters), d, sg, cj, pl in zip(rows, gold_d, gold_s, gold_cj, gold_pl):
        assert (cjtot, plasiters) == (cj, pl)
        assert abs(d1 - d) <= 6e-4 * abs(d)
        assert all(abs(m - g) <= 6e-4 * abs(g) for m, g in zip((sz, sx, sy), sg))
    field = np.where(p.nf > 0, out["totd"][np.maximum(p.nf, 1) - 1], 0.0)
    gold_f = a["p122_displ_010"].astype(np.float64).reshape(3, p.nn).T
    assert np.abs(field - gold_f).max() <= 6e-4 * np.abs(gold_f).max()
    driver.write_res_p122(str(tmp_path / "p122_demo.res"), p, out)
    mine = open(tmp_path / "p122_demo.res").read().splitlines()
    gold_lines = [l for l in res.splitlines() if l.strip()]
    keep = lambda ls: [l.split() for l in ls if l.strip() and not l.startswith(("Time", "This analysis", "This job"))]
    assert keep(mine) == keep(gold_lines)                             # every printed number, as text
    ens = np.loadtxt(str(tmp_path / "p122_demo.ensi.DISPL-000010"), skiprows=4)
    assert np.abs(ens - a["p122_displ_010"]).max() <= 6e-4 * np.abs(gold_f).max()


def test_p122_matches_oracle(deck):
    """The same run against oracle.p122_oracle with the blocked reductions (red_mode 1): equal iteration counts in every
    increment, totd within 1e-9 relative L2, first-point stresses within 1e-9 relative."""
    from oracle import p122_oracle
    base, a = deck
    p = host.read_deck_p122(base)
    ref_rows, ref_totd = p122_oracle.p122(p.g_coord_pp, p.g_g_pp, p.neq, p.phi, p.c, p.psi, p.e, p.v, p.qinc, p.plasits,
                                          p.cjits, p.plastol, p.cjtol, no_f=p.no_f, valf=p.val_f, red_mode=1)
    with solver.Solver(0, 1, 0) as s:
        out = driver.run_p122(p, s)
    for (d1, sz, sx, sy, cjtot, plasiters), o in zip(out["rows"], ref_rows):
        assert (cjtot, plasiters) == (o["cjtot"], o["plasiters"])
        assert abs(d1 - o["disp1"]) <= 1e-9 * abs(o["disp1"])
        assert np.allclose([sz, sx, sy], [o["sigma"][2], o["sigma"][0], o["sigma"][1]], rtol=1e-9, atol=0)
    assert np.linalg.norm(out["totd"] - ref_totd) <= 1e-9 * np.linalg.norm(ref_totd)


def test_p122_loaded_nodes_branch_matches_oracle():
    """The other loading branch of p122.f90:119-138 (loaded nodes, ld0_pp*qinc + bdylds_pp, no fixed freedoms) on a
    p12meshgen cube of 20-node bricks pressed by the p121 load patch until it yields."""
    from oracle import p122_oracle
    p = host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2., e=100.0, v=0.3)
    p.program, p.phi, p.c, p.psi = 122, 20.0, 4.0, 0.0
    p.qinc, p.plasits, p.cjits, p.plastol, p.cjtol, p.loaded_nodes = [0.5, 0.25, 0.25], 60, 400, 1e-4, 1e-6, 1
    ref_rows, ref_totd = p122_oracle.p122(p.g_coord_pp, p.g_g_pp, p.neq, p.phi, p.c, p.psi, p.e, p.v, p.qinc, p.plasits,
                                          p.cjits, p.plastol, p.cjtol, ld0=p.r_pp, red_mode=1)
    with solver.Solver(0, 1, 0) as s:
        out = driver.run_p122(p, s)
    assert len(out["rows"]) == len(ref_rows) == 3 and [o["plasiters"] for o in ref_rows] == [2, 8, 15]   # it yields
    # Here the plastic iterations are many and each PCG solve stops at cjtol = 1e-6 on a stopping ratio that a last-ulp
    # difference in asin / sin / cos of the Lode angle can move across the threshold (measured: 322 against 324 cj
    # iterations in the second increment): plastic iteration counts equal, cj totals within 1 %, fields within 1e-6
    # (the demo deck above, the reference's own case, gives equal counts and 1e-9).
    for (d1, sz, sx, sy, cjtot, plasiters), o in zip(out["rows"], ref_rows):
        assert plasiters == o["plasiters"] and abs(cjtot - o["cjtot"]) <= max(2, 0.01 * o["cjtot"])
    assert np.linalg.norm(out["totd"] - ref_totd) <= 1e-6 * np.linalg.norm(ref_totd)
