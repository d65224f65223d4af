"""Regenerates tests/golden/ from the read-only reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

The reference's example decks and logs for the p121/p123 path are not copied as files; they
are parsed and packed:
  arrays.npz      xx3-tiny deck (coordinates, connectivity in S&G order, restraints, loads) and
                  its golden displacements (.dis); p121_demo loads (.lds) and the golden EnSight
                  displacement field (float32 of the 4-digit values)
  fixtures.json   the small text files (.res logs, .mg / .dat control files) as lists of lines
  p121_demo_digests.json   SHA-256 of the parsed 4 MB p121_demo.d / .bnd arrays (the in-memory
                  generator must reproduce them)
  p121_demo_ensi_head.txt  the first 204 lines of the 107 167-line EnSight golden (format check)
  (arrays.npz also holds xx2-tiny's material numbers, material table and golden displacements; its mesh,
  restraints and loads are xx3-tiny's, asserted here)
  p124_demo_digests.json   SHA-256 of the parsed p124_demo.d / .bnd arrays (25^3 8-node bricks); arrays.npz
                  also holds three of the sixteen golden nodal temperature files (steps 10, 80, 150;
                  float32 of the 5-digit values) and fixtures.json the p124 logs / control files
  p1210_tiny.json / p1210_tiny_dis.npz   the p1210 deck files and log as lines; eight of its hundred golden
                  displacement fields (p1210_tiny.dis)
tests/conftest.py materialises decks and logs from these into a temporary directory with the
repo's own deck writer (pf_write_deck_p121), which reproduces xx3-tiny.d byte for byte.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/parafem/examples"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lines(path):
    return open(path).read().splitlines()


def main():
    from parafem_b200 import host
    from parafem_b200._lib import lib, ptr
    tiny = f"{REF}/dev/xx3/demo/xx3-tiny"
    t = host.read_deck_p121(tiny)
    node = np.empty(8, np.int32)
    val = np.empty((8, 3))
    assert lib().pf_read_lds(tiny.encode(), 8, 3, ptr(node), ptr(val)) == 0
    dis = np.loadtxt(tiny + ".dis", skiprows=2)[:, 1:]
    demo = f"{REF}/5th_ed/p121/demo/p121_demo"
    p = host.read_deck_p121(demo)
    dnode = np.empty(65, np.int32)
    dval = np.empty((65, 3))
    assert lib().pf_read_lds(demo.encode(), 65, 3, ptr(dnode), ptr(dval)) == 0
    disp = np.loadtxt(demo + ".ensi.DISPL-000001", skiprows=4)
    # xx2-tiny: the xx3-tiny mesh with five materials (per-element e, v) -- 59 iterations
    x2 = f"{REF}/dev/xx2/xx2-tiny"
    t2 = host.read_deck_xx2(x2)
    assert np.array_equal(t2.g_num_pp, t.g_num_pp) and np.array_equal(t2.g_coord, t.g_coord) and np.array_equal(t2.rest, t.rest)
    assert np.array_equal(t2.r_pp, t.r_pp)
    xx2 = dict(xx2_etype=t2.etype_pp, xx2_prop=t2.prop, xx2_dis=np.loadtxt(x2 + ".dis", skiprows=2)[:, 1:])
    # xx11 (dev program sharing p123's deck format): 64 8-node bricks, nr = 0, 25 loaded + 25 fixed freedoms.
    # Its golden xx11.ttr belongs to the loads of hexahedron_cube/xx11_hexcube.lds (= xx11.old.lds: 100 per
    # freedom, three value columns of which read_loads takes the first); xx11.lds itself holds 10 per freedom.
    x11 = f"{REF}/dev/xx11/xx11"
    t11 = host.read_deck_p123(f"{REF}/dev/xx11/hexahedron_cube/xx11_hexcube")
    t11b = host.read_deck_p123(x11)
    assert np.array_equal(t11.g_num_pp, t11b.g_num_pp) and np.array_equal(t11.g_coord, t11b.g_coord)
    assert np.array_equal(t11.no_f, t11b.no_f) and np.array_equal(t11.r_pp, 10.0 * t11b.r_pp)
    xx11 = dict(xx11_coord=t11.g_coord, xx11_gnum_sg=t11.g_num_pp, xx11_lds_eq=(np.flatnonzero(t11.r_pp) + 1).astype(np.int32),
                xx11_lds_val=t11.r_pp[np.flatnonzero(t11.r_pp)], xx11_fix_node=t11.no_f, xx11_fix_val=t11.val_f,
                xx11_ttr=np.loadtxt(x11 + ".ttr", skiprows=2)[:, 1])
    # the same cube meshed with 569 4-node tetrahedra (no output shipped): mesh in S&G order, loads, fixed nodes
    tt = host.read_deck_p123(f"{REF}/dev/xx11/tetrahedron_cube/xx11_tetcube")
    xx11.update(xx11tet_coord=tt.g_coord, xx11tet_gnum_sg=tt.g_num_pp, xx11tet_lds_eq=(np.flatnonzero(tt.r_pp) + 1).astype(np.int32),
                xx11tet_lds_val=tt.r_pp[np.flatnonzero(tt.r_pp)], xx11tet_fix_node=tt.no_f, xx11tet_fix_val=tt.val_f)
    # p122 demo deck (elasto-plasticity; oracle only this round): unstructured 8-node mesh, 19 fixed freedoms
    d122 = f"{REF}/5th_ed/p122/demo/p122_demo"
    tk = open(d122 + ".dat").read().split()
    nels2, nn2, nr2, nod2, fix2 = int(tk[3]), int(tk[4]), int(tk[5]), int(tk[7]), int(tk[8])
    gc2, gn2 = np.empty((nn2, 3)), np.empty((nels2, nod2), np.int32)
    assert lib().pf_read_d(d122.encode(), nn2, nels2, nod2, ptr(gc2), ptr(gn2)) == 0 and int(tk[1]) == 1
    rest2 = np.zeros((4, nr2), np.int32)
    assert lib().pf_read_bnd(d122.encode(), nr2, 3, ptr(rest2)) == 0
    fn2, fs2, fv2 = np.empty(fix2, np.int32), np.empty(fix2, np.int32), np.empty(fix2)
    assert lib().pf_read_fix(d122.encode(), fix2, ptr(fn2), ptr(fs2), ptr(fv2)) == 0
    p122 = dict(p122_coord=gc2, p122_gnum_sg=gn2, p122_rest=rest2, p122_fix_node=fn2, p122_fix_sense=fs2, p122_fix_val=fv2,
                p122_displ_010=np.loadtxt(d122 + ".ensi.DISPL-000010", skiprows=4).astype(np.float32))
    # p124 demo deck (transient conduction, 25^3 8-node bricks, Abaqus node order on disk)
    d124 = f"{REF}/5th_ed/p124/demo/p124_demo"
    dat = open(d124 + ".dat").read().split()
    nels4, nn4, nr4, nip4, nod4 = (int(v) for v in dat[4:9])
    gc4 = np.empty((nn4, 3), np.float64)
    gn4 = np.empty((nels4, nod4), np.int32)
    assert lib().pf_read_d(d124.encode(), nn4, nels4, nod4, ptr(gc4), ptr(gn4)) == 0
    assert lib().pf_abaqus2sg(nod4, nels4, ptr(gn4)) == 0
    rest4 = np.zeros((2, nr4), np.int32)
    assert lib().pf_read_bnd(d124.encode(), nr4, 1, ptr(rest4)) == 0
    gcpp4 = np.empty((nels4, 3, nod4), np.float64)
    assert lib().pf_coords_pp(nod4, nels4, gc4.shape[0], ptr(gn4), ptr(gc4), ptr(gcpp4)) == 0
    # (+ 0.0: p124_demo.d prints the top face's -(is-1)*cc = -0.0 unsigned, p121_demo.d prints it signed)
    fsha = lambda path: hashlib.sha256(open(path, "rb").read()).hexdigest()
    files = {f"p124_demo{ext}": fsha(d124 + ext) for ext in (".d", ".bnd", ".dat", ".mat")}
    files.update({f"p125_demo{ext}": fsha(f"{REF}/5th_ed/p125/demo/p125_demo{ext}") for ext in (".d", ".bnd", ".dat")})
    json.dump(dict(files=files, g_num_sg=sha(gn4), g_coord_pp=sha(gcpp4 + 0.0), rest=sha(rest4), nn=nn4, nr=nr4, nels=nels4, nip=nip4,
                   nod=nod4), open(f"{HERE}/p124_demo_digests.json", "w"), indent=1)
    ndttr = {f"p124_ndttr_{j:03d}": np.loadtxt(f"{d124}.ensi.NDTTR-{j:06d}", skiprows=4).astype(np.float32)
             for j in (10, 80, 150)}
    # p125 demo (explicit transient conduction): same 25^3 deck; two of the ten golden nodal files
    d125 = f"{REF}/5th_ed/p125/demo/p125_demo"
    gn5 = np.empty((nels4, nod4), np.int32)
    gc5 = np.empty((nn4, 3), np.float64)
    assert lib().pf_read_d(d125.encode(), nn4, nels4, nod4, ptr(gc5), ptr(gn5)) == 0
    assert lib().pf_abaqus2sg(nod4, nels4, ptr(gn5)) == 0
    assert np.array_equal(gn5, gn4) and np.array_equal(gc5, gc4)            # p125_demo.d == p124_demo.d
    assert open(d125 + ".bnd").read() == open(d124 + ".bnd").read()
    ndpre = {f"p125_ndpre_{j:04d}": np.loadtxt(f"{d125}.ensi.NDPRE-{j:06d}", skiprows=4).astype(np.float32)
             for j in (500, 5000)}
    np.savez_compressed(f"{HERE}/arrays.npz", **ndttr, **ndpre, **xx2, **xx11, **p122, tiny_coord=t.g_coord, tiny_gnum_sg=t.g_num_pp, tiny_rest=t.rest,
                        tiny_lds_node=node, tiny_lds_val=val, tiny_dis=dis, demo_lds_node=dnode, demo_lds_val=dval,
                        demo_displ=disp.reshape(3, p.nn).T.astype(np.float32))
    texts = {
        "xx3-tiny.res": lines(tiny + ".res"), "xx3-tiny.dat": lines(tiny + ".dat"),
        "p121_demo.res": lines(demo + ".res"), "p121_demo.dat": lines(demo + ".dat"), "p121_demo.mg": lines(demo + ".mg"),
        "p121_book.res": lines(f"{REF}/5th_ed/p121/book/p121.res"), "p121_book.mg": lines(f"{REF}/5th_ed/p121/book/p121.mg"),
        "p123_book.res": lines(f"{REF}/5th_ed/p123/book/p123.res"), "p123_book.mg": lines(f"{REF}/5th_ed/p123/book/p123.mg"),
        "xx2-tiny.res": lines(x2 + ".res"), "xx2-tiny.dat": lines(x2 + ".dat"), "xx2-tiny.mat": lines(x2 + ".mat"),
        "p122_demo.dat": lines(d122 + ".dat"), "p122_demo.res": lines(d122 + ".res"),
        "xx11.dat": lines(x11 + ".dat"), "xx11_tetcube.dat": lines(f"{REF}/dev/xx11/tetrahedron_cube/xx11_tetcube.dat"),
        "p125_demo.res": lines(d125 + ".res"), "p125_demo.dat": lines(d125 + ".dat"),
        "p124_demo.res": lines(d124 + ".res"), "p124_demo.dat": lines(d124 + ".dat"), "p124_demo.mat": lines(d124 + ".mat"),
        "p124_book.res": lines(f"{REF}/5th_ed/p124/book/p124.res"), "p124_book.mg": lines(f"{REF}/5th_ed/p124/book/p124.mg"),
        "p124_tiny.mg": lines(f"{REF}/5th_ed/p124/mg/p124_tiny.mg"),
        "p123_small.mg": lines(f"{REF}/5th_ed/p123/mg/p123_small.mg"), "p121_tiny.mg": lines(f"{REF}/5th_ed/p121/mg/p121_tiny.mg"),
    }
    json.dump(texts, open(f"{HERE}/fixtures.json", "w"), indent=1)
    fsha_ = lambda path: hashlib.sha256(open(path, "rb").read()).hexdigest()
    digests = dict(files={f"p121_demo{ext}": fsha_(demo + ext) for ext in (".d", ".bnd", ".lds", ".dat")},
                   g_num_sg=sha(p.g_num_pp), g_coord_pp=sha(p.g_coord_pp), rest=sha(p.rest), g_g=sha(p.g_g_pp),
                   r=sha(p.r_pp), nn=int(p.nn), nr=int(p.nr), neq=int(p.neq), nels=int(p.nels))
    json.dump(digests, open(f"{HERE}/p121_demo_digests.json", "w"), indent=1)
    with open(demo + ".ensi.DISPL-000001") as f, open(f"{HERE}/p121_demo_ensi_head.txt", "w") as g:
        g.writelines([next(f) for _ in range(204)])   # header + the first 200 x-displacements, verbatim
    print("golden fixtures written to", HERE)


def p129_digests():
    """p129_tiny (examples/5th_ed/p129: a deck without outputs -- 2560 20-node bricks, 12 465 nodes, nip = 27): SHA-256 of
    the parsed arrays and the control values of its .dat; the in-memory generators (host.cube_p129, oracle.cube_p129)
    must reproduce them.  Written alone: `python tests/golden/make_golden.py p129`."""
    from parafem_b200 import host
    d = host.read_deck_p129(f"{REF}/5th_ed/p129/p129_tiny")
    out = dict(g_num_sg=sha(d.g_num_pp), g_coord_pp=sha(d.g_coord_pp + 0.0), rest=sha(d.rest), g_g=sha(d.g_g_pp), r=sha(d.r_pp),
               nn=int(d.nn), nr=int(d.nr), neq=int(d.neq), nels=int(d.nels), nres=int(d.nres), nip=int(d.nip),
               dat=dict(rho=d.rho, e=d.e, v=d.v, alpha1=d.alpha1, beta1=d.beta1, nstep=d.nstep, npri=d.npri, theta=d.theta,
                        omega=d.omega, tol=d.tol, limit=d.limit), total_load=d.total_load)
    json.dump(out, open(f"{HERE}/p129_tiny_digests.json", "w"), indent=1)
    print("p129_tiny digests written")


def p1210_golden():
    """p1210_tiny (examples/5th_ed/p1210: five 20-node bricks, a cantilever; the only p1210 deck the reference ships,
    with the displacement fields of 100 output steps of a 300 000-step elasto-plastic run in p1210_tiny.dis): the four
    small deck files and the log as lists of lines, and the golden fields of eight of the hundred steps (float64 of the
    5-digit values).  Written alone: `python tests/golden/make_golden.py p1210`."""
    job = f"{REF}/5th_ed/p1210/p1210_tiny"
    txt = {ext: lines(f"{job}.{ext}") for ext in ("dat", "d", "bnd", "lds", "res")}
    json.dump(txt, open(f"{HERE}/p1210_tiny.json", "w"), indent=0)
    dis = lines(job + ".dis")
    nn, keep, fields = 68, (3000, 6000, 9000, 30000, 60000, 150000, 240000, 300000), {}
    i = 0
    while i < len(dis):
        if dis[i].startswith("*DISPLACEMENT"):
            step = int(dis[i + 1])
            if step in keep:
                fields[str(step)] = np.array([[float(x) for x in l.split()[1:]] for l in dis[i + 2:i + 2 + nn]])
            i += 2 + nn
        else:
            i += 1
    assert len(fields) == len(keep)
    np.savez_compressed(f"{HERE}/p1210_tiny_dis.npz", **fields)
    print("p1210_tiny golden written")


if __name__ == "__main__":
    if sys.argv[1:] == ["p129"]:
        p129_digests()
    elif sys.argv[1:] == ["p1210"]:
        p1210_golden()
    else:
        main()
        p129_digests()
        p1210_golden()
