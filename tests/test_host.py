"""Host-side logic (section B of the C-ABI): partition arithmetic, mesh generators, steering,
deck readers, gather tables.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from parafem_b200 import host


@pytest.mark.parametrize("n,npes", [(8000, 1), (8000, 3), (98360, 7), (5, 8), (1953125, 8)])
def test_partition_matches_reference_formula(n, npes):
    """calc_nels_pp / calc_neq_pp (gather_scatter.f90:217-238, 319-339): host.cpp's single
    even_split against the oracle's literal restatement; ranges tile [1, n]."""
    import ctypes as C
    nxt = 1
    for numpe in range(1, npes + 1):
        cnt, start = host.calc_nels_pp(n, npes, numpe)
        a, b = C.c_int64(), C.c_int64()
        oracle.lib().orc_partition(n, npes, numpe, C.byref(a), C.byref(b))
        assert (cnt, start) == (a.value, b.value) == host.calc_neq_pp(n, npes, numpe)
        assert start == nxt
        nxt += cnt
    assert nxt == n + 1


def test_hex20_steering_equals_find_g3_on_generated_cube():
    p = host.cube_p121(4, 3, 5, 20, aa=1., bb=1., cc=1.)
    import ctypes as C
    from parafem_b200._lib import lib, ptr
    rest = np.zeros((4, p.nr), np.int32)
    assert lib().pf_cube_rest(0, 4, 3, 5, 20, p.nr, ptr(rest)) == 0
    assert np.all(np.diff(rest[0]) > 0)                    # ascending nodes: find_g3's binary search needs it
    assert np.array_equal(oracle.find_g3(p.g_num_pp, rest), p.g_g_pp)


def test_hex8_steering_equals_find_g3():
    p = host.cube_p121(6, 4, 5, 8, aa=1., bb=1., cc=1.)
    from parafem_b200._lib import lib, ptr
    rest = np.zeros((4, p.nr), np.int32)
    assert lib().pf_cube_rest(0, 6, 4, 5, 8, p.nr, ptr(rest)) == 0
    assert np.array_equal(oracle.find_g3(p.g_num_pp, rest), p.g_g_pp)


def test_p121_total_load_is_minus_100():
    """load_p121 scaling (p12meshgen.f90:176-177, 218-219): 'The total load is: -0.1000E+03'."""
    for nod in (20, 8):
        for n in (5, 10, 20):
            p = host.cube_p121(n, n, n, nod)
            assert abs(p.total_load + 100.0) < 1e-9
            assert abs(p.r_pp.sum() + 100.0) < 1e-9


def test_tiny_mg_equals_xx3_tiny_deck(tiny, golden):
    """p121_tiny.mg (5^3, aa=2) generated in memory == the shipped xx3-tiny deck's connectivity
    and steering (the .d file lists nodes in its own coordinate frame)."""
    p = host.cube_p121(5, 5, 5, 20, aa=2., bb=2., cc=2.)
    assert (p.nn, p.nr, p.neq) == (tiny.nn, tiny.nr, tiny.neq)
    assert np.array_equal(p.g_num_pp, tiny.g_num_pp)
    assert np.array_equal(p.g_g_pp, tiny.g_g_pp)


def test_slices_of_a_partitioned_cube_tile_the_serial_cube():
    full = host.cube_p121(6, 5, 4, 20)
    parts = [host.cube_p121(6, 5, 4, 20, npes=3, numpe=k) for k in (1, 2, 3)]
    assert np.array_equal(np.concatenate([q.g_g_pp for q in parts]), full.g_g_pp)
    assert np.array_equal(np.concatenate([q.g_coord_pp for q in parts]), full.g_coord_pp)
    assert np.array_equal(np.concatenate([q.r_pp for q in parts]), full.r_pp)


def test_make_ggl_single_rank_is_identity_plus_dump_slot():
    p = host.cube_p121(4, 4, 4, 20)
    ggl, halo, cnt = host.make_ggl(p)
    assert halo.size == 0 and cnt.sum() == 0
    assert np.array_equal(ggl, p.g_g_pp)                   # slot = equation number, 0 = restrained


def test_make_ggl_multi_rank_tables():
    npes = 4
    full = host.cube_p121(5, 8, 3, 20)
    for numpe in range(1, npes + 1):
        p = host.cube_p121(5, 8, 3, 20, npes=npes, numpe=numpe)
        ggl, halo, cnt = host.make_ggl(p)
        lo, hi = p.ieq_start, p.ieq_start + p.neq_pp
        g = p.g_g_pp
        own = (g >= lo) & (g < hi)
        assert np.array_equal(ggl[own], g[own] - lo + 1)
        assert np.all(ggl[g == 0] == 0)
        rem = (g != 0) & ~own
        assert np.array_equal(halo[ggl[rem] - p.neq_pp - 1], g[rem])
        assert np.all(np.diff(halo) > 0) and cnt[numpe - 1] == 0 and cnt.sum() == halo.size
        # owner of each halo equation from the closed form
        owners = np.array([next(r for r in range(1, npes + 1)
                                if host.calc_neq_pp(full.neq, npes, r)[1] <= q < sum(host.calc_neq_pp(full.neq, npes, r)))
                           for q in halo])
        assert np.array_equal(np.bincount(owners - 1, minlength=npes), cnt)


def test_deck_reader_roundtrip(tmp_path, tiny, golden):
    """read_p121-family readers on the shipped tiny deck: sizes from the .dat, Abaqus->S&G
    permutation applied, loads placed on z freedoms."""
    assert (tiny.nels, tiny.nn, tiny.nr, tiny.nod, tiny.nip) == (125, 756, 396, 20, 8)
    assert (tiny.e, tiny.v, tiny.tol, tiny.limit) == (100.0, 0.3, 1e-5, 200)
    assert np.count_nonzero(tiny.r_pp) == 8
    # S&G ordering puts the 8 corner nodes at local positions 1,3,5,7,13,15,17,19: their
    # coordinates span the element's bounding box
    c = tiny.g_coord_pp[0]
    corners = c[:, [0, 2, 4, 6, 12, 14, 16, 18]]
    assert np.allclose(corners.min(axis=1), c.min(axis=1)) and np.allclose(corners.max(axis=1), c.max(axis=1))


def test_ensight_writer_matches_the_shipped_file_format(tmp_path, golden, demo):
    """dismsh_ensi_p (output.f90:2983-3111) restated in host.cpp: header lines byte-identical to
    p121_demo.ensi.DISPL-000001, one value per line in the golden file's 12-column 0.ddddE+xx
    format, component-major; values (here: the oracle's solution) agree with the golden's to
    the printed digits."""
    import oracle
    km = oracle.form_km_elastic(demo.g_coord_pp, 20, 8, demo.e, demo.v)
    r = oracle.pcg(km, demo.g_g_pp, demo.neq, demo.r_pp, demo.tol, demo.limit, npes=4, red_mode=1)
    disp = host.nodal_values(demo, r["x"])
    assert disp.shape == (demo.nn, 3)
    path = tmp_path / "demo.ensi.DISPL-000001"
    host.write_ensi(path, disp, decimals=4)
    mine = open(path).read().splitlines()
    gold = open(os.path.join(golden, "p121_demo_ensi_head.txt")).read().splitlines()
    assert mine[:4] == gold[:4]
    assert len(mine) == 4 + 3 * demo.nn
    assert all(len(l) == 12 for l in mine[4:])
    import re
    assert all(re.fullmatch(r" {1,2}-?0\.\d{4}E[+-]\d\d", l) for l in mine[4:2000])
    a = np.array([float(l) for l in mine[4:204]])
    b = np.array([float(l) for l in gold[4:204]])
    assert np.abs(a - b).max() <= 1.0001e-4                                # one unit of the 4th printed digit
    assert sum(x == y for x, y in zip(mine[4:204], gold[4:204])) > 100      # most lines are identical text (119 of 200;
    # the rest differ by one unit of the last digit: golden stopped at iteration 295, this solve at 297)
    # y-displacements start after nn x-values (component-major)
    assert float(mine[4 + demo.nn]) == 0.0


def test_deck_writer_reader_round_trip(tmp_path):
    """p12meshgen restated end to end: generate a cube, write <job>.d/.bnd/.lds/.dat in the
    reference's formats, read it back with the read_p121-family readers: same steering and loads;
    coordinates and loads as they survive the E14.6 / E16.8 text fields (round_mode 1)."""
    import ctypes as C
    from parafem_b200._lib import lib, ptr
    for nod, dims in ((20, (5, 4, 3)), (8, (5, 5, 5))):
        nxe, nye, nze = dims
        exact = host.cube_p121(nxe, nye, nze, nod, aa=1. / 3, bb=.7, cc=1.1)
        text = host.cube_p121(nxe, nye, nze, nod, aa=1. / 3, bb=.7, cc=1.1, round_mode=1)
        rest = np.zeros((4, exact.nr), np.int32)
        assert lib().pf_cube_rest(0, nxe, nye, nze, nod, exact.nr, ptr(rest)) == 0
        nn, nr, loaded = C.c_int64(), C.c_int64(), C.c_int64()
        lib().pf_p121_sizes(nxe, nye, nze, nod, C.byref(nn), C.byref(nr), C.byref(loaded))
        node = np.empty(loaded.value, np.int32)
        val = np.empty((loaded.value, 3))
        assert lib().pf_p121_loads(nxe, nze, nod, 1. / 3, .7, 0, ptr(node), ptr(val)) == 0
        g_coord = np.zeros((exact.nn, 3))
        g_coord[exact.g_num_pp - 1] = np.transpose(exact.g_coord_pp, (0, 2, 1))
        job = str(tmp_path / f"cube{nod}")
        host.write_deck_p121(job, nod, 8, 100.0, 0.3, 1e-5, 321, g_coord, exact.g_num_pp, rest, node, val)
        back = host.read_deck_p121(job)
        assert (back.nod, back.nip, back.limit, back.nn, back.nr, back.neq) == (nod, 8, 321, exact.nn, exact.nr, exact.neq)
        assert np.array_equal(back.g_num_pp, exact.g_num_pp)          # Abaqus order on disk, S&G after abaqus2sg
        assert np.array_equal(back.g_g_pp, exact.g_g_pp)
        assert np.array_equal(back.g_coord_pp, text.g_coord_pp)
        assert np.array_equal(back.r_pp, text.r_pp)
        assert np.abs(back.g_coord_pp - exact.g_coord_pp).max() < 1e-5 and np.abs(back.r_pp - exact.r_pp).max() < 1e-6


def test_nodal_values_and_node_partition():
    p = host.cube_p121(4, 4, 4, 20)
    x = np.arange(1, p.neq + 1, dtype=np.float64)
    d = host.nodal_values(p, x)
    m = p.nf > 0
    assert np.array_equal(d[m], p.nf[m].astype(np.float64)) and np.all(d[~m] == 0.0)
    import ctypes as C
    from parafem_b200._lib import lib
    nxt = 1
    for numpe in range(1, 6):
        a, b = C.c_int64(), C.c_int64()
        lib().pf_calc_nodes_pp(p.nn, 5, numpe, C.byref(a), C.byref(b))
        assert b.value == nxt
        nxt += a.value
    assert nxt == p.nn + 1


def test_psize_partitioner_2(tmp_path):
    """calc_nels_pp partitioner 2 = read_nels_pp (input.f90:3108-3196): '<npes> n_1 ... n_npes'."""
    from parafem_b200 import PfError
    job = str(tmp_path / "ext")
    open(job + ".psize", "w").write("3   70 20\n 30\n")             # list-directed: tokens may span lines
    assert [host.read_psize(job, 3, r) for r in (1, 2, 3)] == [(70, 1), (20, 71), (30, 91)]
    with pytest.raises(PfError):
        host.read_psize(job, 2, 1)                                     # "Number of partitions is different ..."
    with pytest.raises(PfError):
        host.read_psize(str(tmp_path / "missing"), 3, 1)
    # the slices of an unevenly partitioned cube tile the serial cube, and the equation partition
    # stays calc_neq_pp whatever the element partition is
    full = host.cube_p121(5, 6, 4, 20, aa=1., bb=1., cc=1.)
    parts = [host.cube_p121(5, 6, 4, 20, aa=1., bb=1., cc=1., npes=3, numpe=r, psize=[70, 20, 30]) for r in (1, 2, 3)]
    assert np.array_equal(np.concatenate([q.g_g_pp for q in parts]), full.g_g_pp)
    assert np.array_equal(np.concatenate([q.g_coord_pp for q in parts]), full.g_coord_pp)
    assert [(q.neq_pp, q.ieq_start) for q in parts] == [host.calc_neq_pp(full.neq, 3, r) for r in (1, 2, 3)]
    with pytest.raises(PfError):
        host.calc_nels_pp(120, 3, 1, psize=[70, 20, 31])


def test_d_reader_fast_path_and_fallback_agree(tmp_path, tiny, golden):
    """<job>.d is parsed from memory in parallel chunks; unusual but legal list-directed input (blank lines, CRLF,
    comma separators, a leading '+') is still read, in line order as the reference does (input.f90:388-390,
    1053); a file with a missing record is an error in both readers."""
    import ctypes as C
    from parafem_b200._lib import lib, ptr
    src = open(os.path.join(golden, "xx3-tiny.d")).read().splitlines()
    k_el = src.index("*ELEMENTS")
    ref_c, ref_n = np.empty((tiny.nn, 3)), np.empty((tiny.nels, 20), np.int32)
    assert lib().pf_read_d(os.path.join(golden, "xx3-tiny").encode(), tiny.nn, tiny.nels, 20, ptr(ref_c), ptr(ref_n)) == 0

    def read(lines, nn=tiny.nn, nels=tiny.nels, eol="\n"):
        job = str(tmp_path / "variant")
        with open(job + ".d", "w", newline="") as f:
            f.write(eol.join(lines) + eol)
        c, n, et = np.full((nn, 3), np.nan), np.zeros((nels, 20), np.int32), np.zeros(nels, np.int32)
        return lib().pf_read_d_mat(job.encode(), nn, nels, 20, ptr(c), ptr(n), ptr(et)), c, n, et

    odd = list(src)
    odd.insert(5, "")                                             # blank record inside the node section
    odd.insert(k_el + 3, "   ")
    odd[3] = "  " + ", ".join(odd[3].split())                     # comma separators
    odd[4] = odd[4].replace(" 0.", " +0.", 1)                     # explicit sign
    for eol in ("\n", "\r\n"):
        rc, c, n, et = read(odd, eol=eol)
        assert rc == 0 and np.array_equal(c, ref_c) and np.array_equal(n, ref_n) and np.all(et == 1)
    # line order, not the leading number, places a record (READ(10,*) bitBucket, g_coord(:,j))
    swapped = list(src)
    swapped[2], swapped[3] = swapped[3], swapped[2]
    rc, c, n, et = read(swapped)
    assert rc == 0 and np.array_equal(c[0], ref_c[1]) and np.array_equal(c[1], ref_c[0]) and np.array_equal(c[2:], ref_c[2:])
    # a record short: both the fast path and the record reader refuse
    rc, *_ = read(src[:-1])
    assert rc != 0
    # Fortran's double-precision exponent letter is read like E
    dexp = [l.replace("E+", "D+").replace("E-", "D-") if 2 <= i < k_el else l for i, l in enumerate(src)]
    rc, c, n, et = read(dexp)
    assert rc == 0 and np.array_equal(c, ref_c) and np.array_equal(n, ref_n)


def test_scalar_deck_writer_reproduces_the_shipped_p124_and_p125_decks(tmp_path):
    """p12meshgen's output side for p124 / p125 (p12meshgen.f90:879-987, 1075-1160): <job>.d (2.8 MB), .bnd, .dat
    (and p124's .mat) of examples/5th_ed/p124/demo and p125/demo, byte for byte (SHA-256 of the shipped files), and
    the readers bring the problem back."""
    import hashlib
    import json
    from parafem_b200._lib import lib, ptr
    dg = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "p124_demo_digests.json")))["files"]
    for prog, make, read in ((124, host.cube_p124, host.read_deck_p124), (125, host.cube_p125, host.read_deck_p125)):
        p = make(25, 25, 25, aa=.04, bb=.04, cc=.04)
        g_coord = np.zeros((p.nn, 3))
        g_coord[p.g_num_pp - 1] = np.transpose(p.g_coord_pp, (0, 2, 1))
        rest = np.zeros((2, p.nr), np.int32)
        assert lib().pf_cube_rest(1, 25, 25, 25, 8, p.nr, ptr(rest)) == 0
        job = str(tmp_path / f"p{prog}_demo")
        host.write_deck_scalar(job, p, g_coord, p.g_num_pp, rest)
        for ext in (".d", ".bnd", ".dat") + ((".mat",) if prog == 124 else ()):
            assert hashlib.sha256(open(job + ext, "rb").read()).hexdigest() == dg[f"p{prog}_demo{ext}"], (prog, ext)
        q = read(job)
        ref = make(25, 25, 25, aa=.04, bb=.04, cc=.04, round_mode=1)
        assert np.array_equal(q.g_num_pp, ref.g_num_pp) and np.array_equal(q.g_g_pp, ref.g_g_pp)
        assert np.array_equal(q.g_coord_pp, ref.g_coord_pp + 0.0)        # the deck prints -0.0 unsigned
        assert (q.neq, q.nres, q.dtim, q.nstep, q.npri, q.val0) == (ref.neq, ref.nres, ref.dtim, ref.nstep, ref.npri, ref.val0)
        if prog == 124:
            assert (q.kx, q.ky, q.kz, q.rho, q.cp, q.theta, q.tol, q.limit) == (1., 1., 1., 1., 1., .5, 1e-4, 100)


def test_meshgen_tool_reproduces_the_shipped_decks_from_their_mg_files(tmp_path, golden):
    """p12meshgen restated as a tool (parafem_b200/meshgen.py): the reference's own <job>.mg inputs give the
    reference's own decks byte for byte -- p121_demo.mg -> p121_demo.{d,bnd,lds,dat} (4 MB .d, signed zeros
    included) and p124_tiny.mg -> p124_demo.{d,bnd,dat,mat} (SHA-256 of the shipped files)."""
    import hashlib
    import json
    import shutil
    from parafem_b200 import meshgen
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    want = dict(json.load(open(os.path.join(gdir, "p121_demo_digests.json")))["files"])
    want.update(json.load(open(os.path.join(gdir, "p124_demo_digests.json")))["files"])
    for mg, job, exts in (("p121_demo.mg", "p121_demo", (".d", ".bnd", ".lds", ".dat")),
                          ("p124_tiny.mg", "p124_demo", (".d", ".bnd", ".dat", ".mat"))):
        base = str(tmp_path / job)
        shutil.copy(os.path.join(golden, mg), base + ".mg")
        assert meshgen.main([base]) == 0
        for ext in exts:
            assert hashlib.sha256(open(base + ext, "rb").read()).hexdigest() == want[job + ext], (job, ext)
    with open(tmp_path / "bad.mg", "w") as f:
        f.write("'p126' 'parafem' 8 2 2 8\n")
    with pytest.raises(Exception, match="not one of"):
        meshgen.generate(str(tmp_path / "bad"))


def test_partitioner_2_on_a_deck_produced_by_the_reference_partition_tool(tmp_path):
    """The reference's own metout2pf (tools/preprocessing/partitioner/metout2pf.c, compiled into oracle/_ref) turns a
    deck + a METIS element partition into a deck sorted by partition with renumbered nodes + <job>.psize -- exactly what
    partitioner 2 (read_nels_pp, input.f90:3108-3196) consumes.  Its output goes through this repo's readers and
    partition arithmetic at 3 ranks, and the solve on it equals the solve on the original deck node by node."""
    import subprocess
    from parafem_b200 import meshgen
    if oracle.ref_tool("metout2pf") is None:              # built by build() where /root/reference exists
        pytest.skip("oracle/_ref/metout2pf was not built (needs /root/reference at build time)")
    d = str(tmp_path)
    with open(f"{d}/cube.mg", "w") as f:
        f.write("'p121'\n'parafem' 60 5 3 20 8\n1.0 1.0 1.0 100.0 0.3\n1.0e-11 3000\n")
    meshgen.generate(f"{d}/cube")
    for ext in (".d", ".bnd", ".lds"):                     # the tool reads records into a 256-byte buffer: compact them
        rows = [" ".join(l.split()) for l in open(f"{d}/cube{ext}")]
        open(f"{d}/cube{ext}", "w").write("\n".join(rows) + "\n")
    # the other half of the reference's METIS glue, pf2metin (deck -> METIS mesh file; 8-node bricks and 4-node
    # tetrahedra only), reads a deck this repo wrote: element count, METIS element type 3, the eight nodes per element
    if oracle.ref_tool("pf2metin"):
        with open(f"{d}/cube8.mg", "w") as f:
            f.write("'p121'\n'parafem' 60 5 3 8 8\n1.0 1.0 1.0 100.0 0.3\n1.0e-5 200\n")
        meshgen.generate(f"{d}/cube8")
        rows = [" ".join(l.split()) for l in open(f"{d}/cube8.d")]
        open(f"{d}/cube8.d", "w").write("\n".join(rows) + "\n")
        res = subprocess.run([oracle.ref_tool("pf2metin"), f"{d}/cube8.d", f"{d}/cube8.met"], capture_output=True, text=True, timeout=120)
        met = open(f"{d}/cube8.met").read().split()
        deck = [r.split() for r in rows[rows.index("*ELEMENTS") + 1:]]
        assert res.returncode == 0 and met[:2] == ["60", "3"] and len(met) == 2 + 60 * 8
        order = (4, 0, 3, 7, 5, 1, 2, 6)                    # the tool's METIS hexahedron order (pf2metin.c:184-188)
        assert [met[2 + 8 * e:10 + 8 * e] for e in range(60)] == [[row[4 + k] for k in order] for row in deck]
    part = np.random.RandomState(4).randint(0, 3, 60)
    open(f"{d}/cube.epart.3", "w").write("\n".join(str(v) for v in part) + "\n")
    res = subprocess.run([oracle.ref_tool("metout2pf"), f"{d}/cube.epart.3", f"{d}/cube", f"{d}/cube_part"],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and os.path.exists(f"{d}/cube_part.psize")
    dat = open(f"{d}/cube.dat").read().splitlines()
    dat[2] = "2"                                            # partitioner 2: element counts from <job>.psize
    open(f"{d}/cube_part.dat", "w").write("\n".join(dat) + "\n")
    sizes = [int(v) for v in open(f"{d}/cube_part.psize").read().split()][1:]
    assert sizes == list(np.bincount(part, minlength=3))
    ranks = [host.read_deck_p121(f"{d}/cube_part", npes=3, numpe=k) for k in (1, 2, 3)]
    assert [r.nels_pp for r in ranks] == sizes and [r.iel_start for r in ranks] == [1, 1 + sizes[0], 1 + sizes[0] + sizes[1]]
    assert sum(r.neq_pp for r in ranks) == ranks[0].neq
    # the solve on the partitioned deck (3 emulated ranks on the tool's partition) ...
    g_g = np.concatenate([r.g_g_pp for r in ranks])
    coord = np.concatenate([r.g_coord_pp for r in ranks])
    rhs = np.concatenate([r.r_pp for r in ranks])
    oracle.set_element_partition(sizes)
    try:
        a = oracle.pcg(oracle.form_km_elastic(coord, 20, 8, 100.0, 0.3), g_g, ranks[0].neq, rhs, 1e-11, 3000, npes=3, red_mode=0)
    finally:
        oracle.set_element_partition(None)
    # ... against the original deck with the loads as the tool rewrote them (single precision, 5 digits)
    o = host.read_deck_p121(f"{d}/cube")
    node, val = np.loadtxt(f"{d}/cube.lds")[:, 0].astype(int), np.loadtxt(f"{d}/cube.lds")[:, 1:]
    val5 = np.array([[float(f"{np.float32(v): 1.4E}") for v in row] for row in val])
    r_o = np.zeros(o.neq)
    for n, v in zip(node, val5):
        for k in range(3):
            if o.nf[n - 1, k]:
                r_o[o.nf[n - 1, k] - 1] = v[k]
    b = oracle.pcg(oracle.form_km_elastic(o.g_coord_pp, 20, 8, 100.0, 0.3), o.g_g_pp, o.neq, r_o, 1e-11, 3000, npes=1, red_mode=0)
    assert a["converged"] and b["converged"] and ranks[0].neq == o.neq
    field = lambda p, x: np.where(p.nf > 0, x[np.maximum(p.nf, 1) - 1], 0.0)
    ua, ub = field(ranks[0], a["x"]), field(o, b["x"])
    key = lambda c: tuple(np.round(c * 4).astype(int))      # coordinates are multiples of 0.5: exact in the tool's % 1.4E
    where = {key(c): i for i, c in enumerate(o.g_coord)}
    perm = np.array([where[key(c)] for c in ranks[0].g_coord])
    assert sorted(perm) == list(range(o.nn))
    assert np.abs(ua - ub[perm]).max() <= 1e-8 * np.abs(ub).max()


def test_xx3_style_res_table(tmp_path, tiny, golden):
    """driver.write_res_xx3 writes the reference GPU driver's log layout: fed the numbers of the shipped xx3-tiny.res
    it reproduces the WHOLE file as text (BASIC JOB DATA block, section table, %total column); device kernel rows are
    appended under 'Solve equations' when given."""
    from parafem_b200 import driver
    gold = open(os.path.join(golden, "xx3-tiny.res")).read().splitlines()
    tiny_4 = host.read_deck_p121(os.path.join(golden, "xx3-tiny"), npes=4, numpe=1)
    assert gold[1].startswith("BASIC JOB DATA") and gold[11].startswith("Setup")
    secs = {name: float(line[44:56]) for name, line in zip(driver.XX3_SECTIONS, gold[11:24])}
    path = tmp_path / "xx3-tiny.res"
    driver.write_res_xx3(path, tiny_4, dict(iters=79, total_load=tiny.total_load), 8, secs)
    assert open(path).read().splitlines() == gold[:len(open(path).read().splitlines())]
    assert len(open(path).read().splitlines()) == 25
    driver.write_res_xx3(path, tiny_4, dict(iters=79, total_load=tiny.total_load), 8, secs,
                         kernels={"mat-vec": (12.5, 79), "scatter": (1.25, 79)})
    out = open(path).read().splitlines()
    kern = [l for l in out if l.startswith("  ")]
    assert len(out) == 27 and out.index(kern[0]) == out.index(gold[22]) + 1          # right under 'Solve equations'
    assert kern[0].startswith("  mat-vec (79 launches)") and kern[0][44:56] == "    0.012500"


def test_deck_node_numbers_are_range_checked(tmp_path, tiny, golden):
    """What a deck file holds is never used as an index unchecked: a connectivity, load or restraint
    record naming a node outside 1..nn, or a .dat with sizes no deck can have, is a status code -- not an
    out-of-bounds read."""
    import shutil
    from parafem_b200 import PfError
    from parafem_b200._lib import lib, ptr
    src = os.path.join(golden, "xx3-tiny")
    good = host.read_deck_p121(src)
    nn = good.nn

    def variant(name, suffix, edit):
        base = str(tmp_path / name)
        for sfx in (".dat", ".d", ".bnd", ".lds"):
            shutil.copy(src + sfx, base + sfx)
        lines = open(base + suffix).read().splitlines()
        edit(lines)
        open(base + suffix, "w").write("\n".join(lines) + "\n")
        return base

    def bad_element(lines):
        k = lines.index("*ELEMENTS") + 3
        f = lines[k].split()
        f[6] = str(nn + 7)                      # a node number past the end of the coordinate list
        lines[k] = " ".join(f)
    with pytest.raises(PfError):
        host.read_deck_p121(variant("bad_d", ".d", bad_element))

    def bad_load(lines):
        f = lines[0].split()
        f[0] = "0"
        lines[0] = " ".join(f)
    with pytest.raises(PfError):
        host.read_deck_p121(variant("bad_lds", ".lds", bad_load))

    def bad_rest(lines):
        f = lines[2].split()
        f[0] = str(nn + 1)
        lines[2] = " ".join(f)
    with pytest.raises(PfError):
        host.read_deck_p121(variant("bad_bnd", ".bnd", bad_rest))

    def bad_dat(lines):
        for i, l in enumerate(lines):
            f = l.split()
            if len(f) >= 6 and f[0].isdigit() and int(f[0]) == good.nels:
                f[1] = "-5"                     # nn
                lines[i] = " ".join(f)
                return
        raise AssertionError("size line not found")
    with pytest.raises(PfError):
        host.read_deck_p121(variant("bad_dat", ".dat", bad_dat))
    # the steering routines themselves
    g_num = good.g_num_pp.copy()
    g_num[3, 5] = nn + 1
    g_g = np.empty_like(good.g_g_pp)
    assert lib().pf_find_g(20, 3, good.nels_pp, nn, ptr(g_num), ptr(good.nf), ptr(g_g)) == 5
    out = np.empty((good.nels_pp, 3, 20))
    assert lib().pf_coords_pp(20, good.nels_pp, nn, ptr(g_num), ptr(np.zeros((nn, 3))), ptr(out)) == 5
    r = np.empty(good.neq_pp)
    assert lib().pf_load(3, 1, nn, ptr(np.array([nn + 1], np.int32)), ptr(np.ones(3)), ptr(good.nf), 1, good.neq_pp, ptr(r)) == 5


@pytest.mark.parametrize("prog,nod", [("p121", 20), ("p121", 8), ("p123", 8)])
def test_binary_geometry_deck_round_trip(tmp_path, prog, nod):
    """<job>.bin.ensi.geo as p12meshgenbin writes it (mesh_ensi_geo_bin, input.f90:7986-8164) and
    read_g_coord_pp_be / read_g_num_pp_be read it (input.f90:632-790, 1254-1420), against the ASCII deck of the same
    mesh: identical steering (connectivity, g_g_pp, loads), coordinates equal after the trip through the file's
    C floats, and the byte layout of the header (nine 80-character records, part number, nn)."""
    import struct
    from parafem_b200 import meshgen
    job = str(tmp_path / "box")
    if prog == "p121":
        open(job + ".mg", "w").write(f"'p121'\n'parafem'\n{6 * 5 * 4} 6 4 {nod} 8\n0.5 0.4 0.25 100.0 0.3\n1.0e-5 300\n")
    else:
        open(job + ".mg", "w").write("'p123'\n'parafem'\n120 6 4 8\n0.5 0.4 0.25 2.0 2.0 2.0\n1.0e-5 300\n1 0\n")
    written = meshgen.generate(job, binary=True)
    raw = open(job + ".bin.ensi.geo", "rb").read()
    recs = [raw[80 * k:80 * k + 80].decode().rstrip() for k in range(6)]
    assert recs == ["C Binary", "Problem name: box", "Geometry files", "node id off", "element id off", "part"]
    assert struct.unpack("<i", raw[480:484])[0] == 1 and raw[484:564].decode().rstrip() == "Volume"
    assert raw[564:644].decode().rstrip() == "coordinates" and struct.unpack("<i", raw[644:648])[0] == written.nn
    off = 648 + 12 * written.nn
    assert raw[off:off + 80].decode().rstrip() == ("hexa20" if nod == 20 else "hexa8")
    assert struct.unpack("<i", raw[off + 80:off + 84])[0] == written.nels and len(raw) == off + 84 + 4 * nod * written.nels
    read = host.read_deck_p121 if prog == "p121" else host.read_deck_p123
    a, b = read(job), read(job, binary=True)
    assert np.array_equal(a.g_num_pp, b.g_num_pp) and np.array_equal(a.g_g_pp, b.g_g_pp) and np.array_equal(a.r_pp, b.r_pp)
    assert np.array_equal(b.g_coord_pp, a.g_coord_pp.astype(np.float32).astype(np.float64))
    assert np.abs(a.g_coord_pp - b.g_coord_pp).max() <= 2e-7 * np.abs(a.g_coord_pp).max()
    # the first element's connectivity in the file is EnSight's order of the S&G numbering (input.f90:8099-8119)
    first = struct.unpack(f"<{nod}i", raw[off + 84:off + 84 + 4 * nod])
    order = [1, 4, 8, 5, 2, 3, 7, 6] if nod == 8 else [1, 7, 19, 13, 3, 5, 17, 15, 8, 12, 20, 9, 4, 11, 16, 10, 2, 6, 18, 14]
    assert list(first) == [int(written.g_num_pp[0, k - 1]) for k in order]
    # a truncated file and a node number out of range are status codes
    from parafem_b200 import PfError
    open(job + ".bin.ensi.geo", "wb").write(raw[:-8])
    with pytest.raises(PfError):
        read(job, binary=True)
    bad = bytearray(raw)
    bad[off + 84:off + 88] = struct.pack("<i", written.nn + 3)
    open(job + ".bin.ensi.geo", "wb").write(bytes(bad))
    with pytest.raises(PfError):
        read(job, binary=True)


def test_calc_npes_pp_is_the_reference_table():
    """gather_scatter.f90:376-389: npes for up to 15 ranks, then /2, /4, /7, /12."""
    from parafem_b200._lib import lib
    L = lib()
    assert [L.pf_calc_npes_pp(n) for n in (1, 8, 15, 16, 32, 33, 256, 257, 1024, 1025, 12000)] == [1, 8, 15, 8, 16, 8, 64, 36, 146, 85, 1000]
    assert L.pf_calc_npes_pp(0) == 0
