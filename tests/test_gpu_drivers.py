"""The host drivers above the C-ABI (C++ p121_b200, Python driver.py) reproduce the text lines
the reference's own regression test greps from <job>.res
(programs/5th_ed/p121/test.sh:7-35 with the patterns in programs/5th_ed/p121/tests:24-29)."""
import os
import re
import subprocess

import pytest

from parafem_b200 import driver, host, solver

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_lines(golden):
    return open(os.path.join(golden, "p121_demo.res")).read().splitlines()


def test_cpp_driver_on_demo_cube(golden, tmp_path):
    exe = os.path.join(ROOT, "parafem_b200", "p121_b200")
    res = subprocess.run([exe, "--cube", "20", "20"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert res.returncode == 0, res.stderr
    out = res.stdout.splitlines()
    gold = golden_lines(golden)
    # identical text lines, as test.sh requires
    assert gold[1] in out                                   # "There are 35721 nodes 6081 restrained and 98360 equations"
    assert gold[4] in out                                   # "The total load is: -0.1000E+03"
    assert gold[7] in out                                   # "The central nodal displacement is : -0.8571E+00"
    it = int(re.search(r"iterations to convergence was\s+(\d+)", res.stdout).group(1))
    assert abs(it - 295) <= 2                                # golden 295; see DESIGN.md section 2
    sig = [float(v) for v in out[out.index("Point     1") + 1].split()]
    gold_sig = [float(v) for v in gold[10].split()]
    assert all(abs(a - b) < 6e-3 for a, b in zip(sig[:3], gold_sig[:3]))


def test_cpp_driver_reads_the_tiny_deck(golden, tmp_path):
    exe = os.path.join(ROOT, "parafem_b200", "p121_b200")
    res = subprocess.run([exe, os.path.join(golden, "xx3-tiny")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "restrained and         1640 equations" in res.stdout
    it = int(re.search(r"iterations to convergence was\s+(\d+)", res.stdout).group(1))
    assert abs(it - 79) <= 1
    assert "The total load is: -0.1000E+03" in res.stdout
    assert os.path.exists(os.path.join(golden, "xx3-tiny.b200.ensi.DISPL-000001"))


def test_python_driver_res_file(golden, tmp_path, demo):
    with solver.Solver(0, 1, 0) as s:
        res = driver.run(demo, s)
    path = tmp_path / "p121_demo.res"
    driver.write_res(path, demo, res)
    out = open(path).read().splitlines()
    gold = golden_lines(golden)
    for k in (1, 4, 7):
        assert gold[k] in out, (gold[k], out)
    assert abs(res["iters"] - 295) <= 2


def test_python_driver_p123(tmp_path):
    p = host.cube_p123(10, 10, 10)
    with solver.Solver(0, 1, 0) as s:
        res = driver.run(p, s)
    path = tmp_path / "p123.res"
    driver.write_res(path, p, res)
    txt = open(path).read()
    assert "The total load is   0.1000E+02" in txt             # p123.res line format
    assert res["converged"]


def test_cpp_scalar_driver_reproduces_p124_and_p125_golden_logs(golden):
    """p12x_b200 (C++ host code above the C-ABI) on the 25^3 demo box, deck-rounded coordinates: every
    '  Time  Temperature  Iterations' line of p124_demo.res and every '  Time  Pressure' line of p125_demo.res,
    as text."""
    exe = os.path.join(ROOT, "parafem_b200", "p12x_b200")
    for prog, pat in (("p124", r"^\s+0\.\d+E[+-]\d+\s+0\.\d+E[+-]\d+\s+\d+\s*$"), ("p125", r"^\s+0\.\d+E[+-]\d+\s+0\.\d+E[+-]\d+\s*$")):
        res = subprocess.run([exe, prog, "25", "1"], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr
        out = res.stdout.splitlines()
        rows = [l for l in open(os.path.join(golden, f"{prog}_demo.res")).read().splitlines() if re.match(pat, l)]
        assert len(rows) == (15 if prog == "p124" else 11)
        for line in rows:
            assert line in out, (prog, line, out)
        assert [int(v) for v in re.findall(r"\d+", [l for l in out if l.startswith("There are")][0])] == [17576, 1951, 15625]


def test_cpp_scalar_driver_p123_matches_python_driver():
    exe = os.path.join(ROOT, "parafem_b200", "p12x_b200")
    res = subprocess.run([exe, "p123", "20"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    p = host.cube_p123(20, 20, 20, limit=10000)
    with solver.Solver(0, 1, 0) as s:
        ref = driver.run(p, s)
    it = int(re.search(r"iterations to convergence was\s+(\d+)", res.stdout).group(1))
    assert it == ref["iters"]
    assert f"{p.nres:8d}     {driver._fe(ref['x'][p.nres - 1])}" in res.stdout.splitlines()
    assert "The total load is   0.1000E+02" in res.stdout


def test_driver_cli_writes_res_and_ensight(tmp_path):
    rc = driver.main(["--cube", "10", "--out", str(tmp_path)])
    assert rc == 0
    res = open(tmp_path / "p121_cube10_hex20.res").read()
    assert "The total load is: -0.1000E+03" in res and "iterations to convergence" in res
    ens = open(tmp_path / "p121_cube10_hex20.ensi.DISPL-000001").read().splitlines()
    p = host.cube_p121(10, 10, 10, 20)
    assert ens[0].startswith("Alya Ensight Gold --- Vector") and len(ens) == 4 + 3 * p.nn
    rc = driver.main(["--p123", "8", "--out", str(tmp_path)])
    assert rc == 0
    ens = open(tmp_path / "p123_box8.ensi.NDPTL-000001").read().splitlines()
    assert ens[0].startswith("Alya Ensight Gold --- Scalar") and len(ens) == 4 + 9 ** 3
    assert driver.main(["--p124", "6", "--out", str(tmp_path)]) == 0
    assert len(open(tmp_path / "p124_box6.res").read().splitlines()) == 4 + 1 + 15 + 3
    assert os.path.exists(tmp_path / "p124_box6.ensi.NDTTR-000150")
    assert driver.main(["--p125", "6", "--out", str(tmp_path)]) == 0
    assert len(open(tmp_path / "p125_box6.res").read().splitlines()) == 6 + 11 + 2
    assert os.path.exists(tmp_path / "p125_box6.ensi.NDPRE-005000")


def test_bad_arguments_return_status_codes():
    """Errors are status codes + pf_last_error text (xx3 convention), never exit()/abort()."""
    import ctypes as C
    import numpy as np
    from parafem_b200 import PfError, solver as S
    from parafem_b200._lib import lib, ptr
    p = host.cube_p121(3, 3, 3, 20)
    with S.Solver(0, 1, 0) as s:
        with pytest.raises(PfError, match="neq_pp/ieq_start"):
            s._ck(lib().pf_setup_mesh(s._h, 20, 3, 8, p.nels_pp, ptr(p.g_coord_pp), ptr(p.g_g_pp), p.neq, 1, p.neq - 1), "x")
        with pytest.raises(PfError, match="nod must be"):
            s._ck(lib().pf_setup_mesh(s._h, 10, 3, 8, p.nels_pp, ptr(p.g_coord_pp), ptr(p.g_g_pp), p.neq, 1, p.neq), "x")
        bad = p.g_g_pp.copy(); bad[0, 0] = p.neq + 5
        with pytest.raises(PfError, match="outside"):
            s._ck(lib().pf_setup_mesh(s._h, 20, 3, 8, p.nels_pp, ptr(p.g_coord_pp), ptr(bad), p.neq, 1, p.neq), "x")
        s.setup_mesh(p)
        with pytest.raises(PfError, match="element matrices"):
            s.build_precon()
        s.form_km_elastic(p.e, p.v)
        with pytest.raises(PfError, match="pf_build_precon"):
            s.pcg_run(1e-5, 10)
        s.build_precon()
        with pytest.raises(PfError, match="limit"):
            s.pcg_run(1e-5, 0)
        with pytest.raises(PfError):
            s.get_storkm(p.nels_pp - 1, 5)
    h = C.c_void_p()
    assert lib().pf_init(0, 1, 99, None, C.byref(h)) > 0          # no such device
    assert lib().pf_init(3, 2, 0, None, C.byref(h)) > 0           # rank outside [0, nranks)
