"""p1210 on the device (pf_vm_explicit_begin / _steps / _get): explicit elasto-plastic (von Mises) dynamics with a lumped
mass.  No transcendental function but sqrt, no reduction: the device is compared with the oracle BIT FOR BIT, and with the
reference's golden displacement fields (p1210_tiny.dis) to the digits printed."""
import numpy as np
import pytest

import oracle
from parafem_b200 import driver, host, solver
from p1210_util import GOLDEN_PLOAD, equal_to_printed_digits, golden_fields, nodal, synthetic, write_tiny_deck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    s = solver.Solver(0, 1, 0)
    yield s
    s.close()


def test_tiny_deck_equals_oracle_and_golden(gpu, tmp_path):
    p = host.read_deck_p1210(write_tiny_deck(tmp_path))
    p.pload = GOLDEN_PLOAD                       # tests/p1210_util.py: what the deck's last line does not say
    nstep = 60000                                # 20 of the golden's 100 output steps; yield starts before step 30000
    res = driver.run_p1210(p, gpu, nstep=nstep)
    ref = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, nstep, p.npri)
    assert np.array_equal(gpu.vm_explicit_get(mass=True)[3], ref["mm"])
    assert len(res["fields"]) == 20
    for step, x1, d1, d2 in ref["snaps"]:
        assert np.array_equal(res["fields"][step], x1), step
    assert np.array_equal(res["x"], ref["snaps"][-1][1]) and np.array_equal(res["d1x"], ref["snaps"][-1][2])
    assert np.array_equal(res["d2x"], ref["snaps"][-1][3])
    gold = golden_fields()
    for step in (3000, 6000, 9000, 30000, 60000):
        assert equal_to_printed_digits(nodal(p, res["fields"][step]), gold[step]), step
    # the log: header, the t = 0 row and one row per output step, in the reference's E12.4
    out = tmp_path / "p1210_tiny.res"
    driver.write_res_p1210(str(out), p, res)
    lines = out.read_text().splitlines()
    assert lines[1] == "There are           68 nodes            8 restrained and          180 equations"
    assert lines[3] == "  Time      Displacement  Velocity   Acceleration " and lines[4].split() == ["0.0000E+00"] * 4
    assert len(lines) == 4 + 1 + 20 + 1 and lines[5].startswith("  0.3000E-02")


@pytest.mark.parametrize("shape", [(4, 5, 3), (3, 7, 2)])
def test_synthetic_yielding_case_equals_oracle(gpu, shape):
    p = synthetic(host, *shape)
    res = driver.run_p1210(p, gpu)
    ref = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, p.nstep, p.npri)
    for step, x1, d1, d2 in ref["snaps"]:
        assert np.array_equal(res["fields"][step], x1), step
    assert np.array_equal(res["d1x"], ref["snaps"][-1][2]) and np.array_equal(res["d2x"], ref["snaps"][-1][3])
    el = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, 1e30, p.rho, p.dtim, p.pload, p.nstep, p.npri)
    assert np.abs(el["snaps"][-1][1] - res["x"]).max() > 1e-3 * np.abs(res["x"]).max()      # it did yield


def test_operator_form_on_the_tensor_cores(gpu, tmp_path):
    """pf_vm_explicit_set_form(1): the Gauss-point update inside the matrix-free tensor-core pipeline (k_apply_mf4<MID 1>).
    Another rounding of elements_2 -- bit-equal to ITS oracle mirror (orc_p1210_elements_mf), 1e-12 from the form the
    reference writes, and the golden fields to the digits printed."""
    p = host.read_deck_p1210(write_tiny_deck(tmp_path))
    p.pload, p.form = GOLDEN_PLOAD, 1
    nstep = 60000
    res = driver.run_p1210(p, gpu, nstep=nstep)
    ref = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, nstep, p.npri, form=1)
    for step, x1, d1, d2 in ref["snaps"]:
        assert np.array_equal(res["fields"][step], x1), step
    assert np.array_equal(res["d1x"], ref["snaps"][-1][2]) and np.array_equal(res["d2x"], ref["snaps"][-1][3])
    gold = golden_fields()
    for step in (3000, 6000, 9000, 30000, 60000):
        assert equal_to_printed_digits(nodal(p, res["fields"][step]), gold[step]), step
    ref0 = oracle.p1210(p.g_coord_pp, p.g_g_pp, p.neq, p.r_pp, p.e, p.v, p.sbary, p.rho, p.dtim, p.pload, 9000, p.npri, form=0)
    for step, x1, _, _ in ref0["snaps"]:
        assert np.abs(res["fields"][step] - x1).max() <= 1e-12 * np.abs(x1).max()
    # a mesh with ragged passes (60 and 42 elements: not multiples of 8) that yields
    for shape in ((4, 5, 3), (3, 7, 2)):
        q = synthetic(host, *shape)
        q.form = 1
        res = driver.run_p1210(q, gpu)
        ref = oracle.p1210(q.g_coord_pp, q.g_g_pp, q.neq, q.r_pp, q.e, q.v, q.sbary, q.rho, q.dtim, q.pload, q.nstep, q.npri, form=1)
        for step, x1, d1, d2 in ref["snaps"]:
            assert np.array_equal(res["fields"][step], x1), (shape, step)
        el = oracle.p1210(q.g_coord_pp, q.g_g_pp, q.neq, q.r_pp, q.e, q.v, 1e30, q.rho, q.dtim, q.pload, q.nstep, q.npri, form=1)
        assert np.abs(el["snaps"][-1][1] - res["x"]).max() > 1e-3 * np.abs(res["x"]).max()


def test_all_300000_steps_on_the_device_in_the_tensor_core_form(gpu, tmp_path):
    """The whole golden run on the device (form 1): every kept output step of p1210_tiny.dis, first yield, plastic cycling
    and the last step, to the digits printed.  (Form 0 equals its oracle bit for bit, and that oracle is checked over the
    300 000 steps on CPU in tests/test_oracle_p1210.py.)"""
    p = host.read_deck_p1210(write_tiny_deck(tmp_path))
    p.pload, p.form = GOLDEN_PLOAD, 1
    res = driver.run_p1210(p, gpu)
    gold = golden_fields()
    assert len(res["fields"]) == 100
    for step, g in gold.items():
        assert equal_to_printed_digits(nodal(p, res["fields"][step]), g), step


@pytest.mark.parametrize("form", [0, 1])
def test_cpp_driver_writes_the_golden_displacement_file(tmp_path, form):
    """p1210_b200 (C++ host code above the C-ABI) on the shipped deck: its <job>.b200.dis, in the layout of the reference's
    p1210_tiny.dis, holds the golden fields to the digits printed; the log has the reference's lines."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    job = write_tiny_deck(tmp_path)
    run = subprocess.run([os.path.join(root, "parafem_b200", "p1210_b200"), job, str(GOLDEN_PLOAD), str(form), "30000"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = open(job + ".b200.dis").read().splitlines()
    assert len(lines) == 10 * 70 and lines[0].startswith("*DISPLACEMENT") and lines[1].split() == ["3000"]
    gold = golden_fields()
    for k, step in enumerate(range(3000, 30001, 3000)):
        blk = lines[70 * k:70 * (k + 1)]
        assert int(blk[1]) == step
        if step in gold:
            ours = np.array([[float(x) for x in l.split()[1:]] for l in blk[2:]])
            assert equal_to_printed_digits(ours, gold[step]), step
    res = open(job + ".b200.res").read().splitlines()
    assert res[1] == "There are           68 nodes            8 restrained and          180 equations"
    assert res[3] == "  Time      Displacement  Velocity   Acceleration " and len(res) == 4 + 1 + 10 + 1


def test_needs_twenty_node_bricks(gpu):
    from parafem_b200 import PfError
    p = host.cube_p121(3, 3, 3, 8, aa=1., bb=1., cc=1.)
    p.program, p.rho, p.sbary, p.dtim, p.pload = 1210, 1.0, 4.0, 1e-3, 1.0
    with pytest.raises(PfError):
        solver.setup_problem(gpu, p)
