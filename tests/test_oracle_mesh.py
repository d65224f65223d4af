"""The oracle's own p12meshgen restatement (oracle/pf_oracle.c: orc_cube_elements / orc_cube_rest / orc_load_p121 /
orc_find_g_all) against the product's host library: two independent restatements of geometry_*bxz, cube_bc*, box_bc8,
load_p121, rearrange and find_g3 / find_g4 must agree bit for bit.  It is the mesh of bench.py's reference arm, which
must run without ever mapping libparafem_b200.so, on every host core regardless of the launcher's OMP_NUM_THREADS."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from parafem_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dims,nod", [((5, 5, 5), 20), ((6, 7, 5), 20), ((5, 9, 2), 20), ((10, 9, 6), 8), ((10, 10, 10), 8),
                                      ((20, 20, 20), 20)])
def test_p121_cube_equals_host_library(dims, nod):
    a = oracle.cube_p121(*dims, nod=nod, aa=.5, bb=.4, cc=.3)
    b = host.cube_p121(*dims, nod, aa=.5, bb=.4, cc=.3)
    assert (a.nn, a.nr, a.neq, a.nels) == (b.nn, b.nr, b.neq, b.nels)
    assert np.array_equal(a.g_num_pp, b.g_num_pp) and np.array_equal(a.g_coord_pp, b.g_coord_pp)
    assert np.array_equal(a.g_g_pp, b.g_g_pp) and np.array_equal(a.r_pp, b.r_pp)
    assert a.total_load == b.total_load


def test_p121_demo_deck_rounding_equals_host_library(demo):
    """The demo deck (coordinates through E14.6, loads through E16.8) from the oracle's generator."""
    a = oracle.cube_p121(20, 20, 20, 20, aa=.5, bb=.5, cc=.5, deck_rounding=True)
    assert np.array_equal(a.g_coord_pp, demo.g_coord_pp) and np.array_equal(a.r_pp, demo.r_pp)
    assert (a.nn, a.nr, a.neq) == (35721, 6081, 98360)            # p121_demo.res


@pytest.mark.parametrize("dims", [(9, 11, 8), (10, 10, 10), (4, 3, 7)])
def test_p123_box_equals_host_library(dims):
    a, b = oracle.cube_p123(*dims), host.cube_p123(*dims)
    assert (a.nn, a.nr, a.neq, a.nres) == (b.nn, b.nr, b.neq, b.nres)
    assert np.array_equal(a.g_num_pp, b.g_num_pp) and np.array_equal(a.g_coord_pp, b.g_coord_pp)
    assert np.array_equal(a.g_g_pp, b.g_g_pp) and np.array_equal(a.r_pp, b.r_pp)


def test_book_case_sizes_from_the_oracle_generator():
    """p121.res (40^3 hex20): 270 641 nodes, 24 161 restrained, 777 520 equations; p123.res (200^3 hex8) sizes by formula."""
    a = oracle.cube_p121(40, 40, 40, 20, aa=.25, bb=.25, cc=.25)
    assert (a.nn, a.nr, a.neq) == (270641, 24161, 777520)


def test_reference_arm_maps_only_the_oracle_and_uses_every_core():
    """bench.py --impl reference under OMP_NUM_THREADS=1 (what torchrun exports): the process maps oracle/libpf_oracle.so
    and not the product's library, and reports the cores of its affinity mask."""
    code = (
        "import sys, runpy, json\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '1', '--cpu-n', '6']\n"
        "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
        "print('MAPS', json.dumps(sorted({l.split()[-1] for l in open('/proc/self/maps') if '.so' in l and '/repo/' in l})))\n")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    maps = json.loads([l for l in res.stdout.splitlines() if l.startswith("MAPS ")][0][5:])
    assert any(m.endswith("oracle/libpf_oracle.so") for m in maps)
    assert not any("libparafem_b200" in m for m in maps), maps
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][0])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == line["host"]["cores"]
    assert line["config"]["nels"] == 216 and line["same_workload_as_gpu_arm"] is False


def test_p129_cantilever_reproduces_the_shipped_deck():
    """examples/5th_ed/p129/p129_tiny.{d,bnd,lds,dat} (the reference ships this deck without outputs): both in-memory
    generators -- the product's host.cube_p129 and the oracle's cube_p129 -- reproduce its connectivity, coordinates,
    restraints, steering array and (after the deck's E16.8) its loads: digests of the parsed reference files,
    tests/golden/make_golden.py."""
    import hashlib
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "p129_tiny_digests.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    h = host.cube_p129(8, 40, 8, .125, .125, .125)
    o = oracle.cube_p129(8, 40, 8, .125, .125, .125, deck_rounding=True)
    assert (h.nn, h.nr, h.neq, h.nels, h.nres) == (o.nn, o.nr, o.neq, o.nels, o.nres) == (d["nn"], d["nr"], d["neq"], d["nels"], d["nres"])
    assert (d["nn"], d["nr"], d["neq"], d["nels"], d["nres"], d["nip"]) == (12465, 225, 36720, 2560, 36072, 27)
    for m in (h, o):
        assert sha(m.g_num_pp) == d["g_num_sg"] and sha(m.g_coord_pp + 0.0) == d["g_coord_pp"] and sha(m.g_g_pp) == d["g_g"]   # (+ 0.0: the deck prints -0.0 as 0.0000)
        assert sha(m.rest) == d["rest"]
    assert sha(o.r_pp) == d["r"]                                        # loads through E16.8
    assert np.abs(h.r_pp - o.r_pp).max() <= 5e-8 and abs(h.total_load - 100.0) < 1e-12
