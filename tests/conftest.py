import json
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PACKED = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared objects are built in-tree once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden(tmp_path_factory, _built):
    """The reference's decks and logs, materialised from the packed fixtures (tests/golden/
    make_golden.py) into a scratch directory: decks through the repo's own deck writer, logs
    and control files from their stored lines."""
    from parafem_b200 import host
    d = str(tmp_path_factory.mktemp("golden"))
    texts = json.load(open(os.path.join(PACKED, "fixtures.json")))
    for name, ls in texts.items():
        with open(os.path.join(d, name), "w") as f:
            f.write("\n".join(ls) + "\n")
    a = np.load(os.path.join(PACKED, "arrays.npz"))
    xx3_dat = os.path.join(d, "xx3-tiny.dat")
    keep = open(xx3_dat).read()               # the xx3-style .dat (leading 'gpu' tag) stays as shipped
    host.write_deck_p121(os.path.join(d, "xx3-tiny"), 20, 8, 100.0, 0.3, 1e-5, 200, a["tiny_coord"], a["tiny_gnum_sg"],
                         a["tiny_rest"], a["tiny_lds_node"], a["tiny_lds_val"])
    open(xx3_dat, "w").write(keep)
    # xx2-tiny = the same mesh, restraints and loads with five materials: the element lines of the .d carry
    # the material number in their last column; .dat / .mat as shipped
    x2 = os.path.join(d, "xx2-tiny")
    keep2 = open(x2 + ".dat").read()
    host.write_deck_p121(x2, 20, 8, 100.0, 0.3, 1e-5, 200, a["tiny_coord"], a["tiny_gnum_sg"], a["tiny_rest"],
                         a["tiny_lds_node"], a["tiny_lds_val"])
    open(x2 + ".dat", "w").write(keep2)
    body = open(x2 + ".d").read().splitlines()
    k0 = body.index("*ELEMENTS") + 1
    for e, m in enumerate(a["xx2_etype"]):
        assert body[k0 + e].endswith(" 1")
        body[k0 + e] = body[k0 + e][:-1] + str(int(m))
    open(x2 + ".d", "w").write("\n".join(body) + "\n")
    with open(x2 + ".dis", "w") as f:
        f.write("*DISPLACEMENT\n           1\n")
        for i, u in enumerate(a["xx2_dis"]):
            f.write(f"{i + 1:8d} {u[0]: .4E} {u[1]: .4E} {u[2]: .4E}\n")
    # xx11 (p123 deck format, nr = 0, loads + fixed freedoms) on bricks and on tetrahedra: rewritten in S&G node
    # order, so meshgen = 1; loads in the three-value-column form of xx11_hexcube.lds (read_loads takes the first)
    for job, key, nod in (("xx11", "xx11", 8), ("xx11_tetcube", "xx11tet", 4)):
        base = os.path.join(d, job)
        dat = open(base + ".dat").read().split()
        dat[1] = "1"
        open(base + ".dat", "w").write("\n".join(dat[:3]) + "\n" + " ".join(dat[3:10]) + "\n" + " ".join(dat[10:]) + "\n")
        with open(base + ".d", "w") as f:
            f.write("*THREE_DIMENSIONAL\n*NODES\n")
            for i, c in enumerate(a[key + "_coord"]):
                f.write(f"{i + 1}  {float(c[0])!r}  {float(c[1])!r}  {float(c[2])!r}\n")
            f.write("*ELEMENTS\n")
            for e, g in enumerate(a[key + "_gnum_sg"]):
                f.write(f"{e + 1}  3  {nod}  1  " + "  ".join(str(int(v)) for v in g) + "  1\n")
        with open(base + ".lds", "w") as f:
            for q, v in zip(a[key + "_lds_eq"], a[key + "_lds_val"]):
                f.write(f"{int(q)}   {float(v)!r} 0.0 0.0\n")
        with open(base + ".fix", "w") as f:
            for q, v in zip(a[key + "_fix_node"], a[key + "_fix_val"]):
                f.write(f"{int(q)} 1 {float(v)!r}\n")
    with open(os.path.join(d, "xx11.ttr"), "w") as f:
        f.write("*TEMPERATURE\n 1\n")
        for i, v in enumerate(a["xx11_ttr"]):
            f.write(f"{i + 1:8d} {v: .4E}\n")
    with open(os.path.join(d, "xx3-tiny.dis"), "w") as f:
        f.write("*DISPLACEMENT\n           1\n")
        for i, u in enumerate(a["tiny_dis"]):
            f.write(f"{i + 1:8d} {u[0]: .4E} {u[1]: .4E} {u[2]: .4E}\n")
    with open(os.path.join(d, "p121_demo.lds"), "w") as f:
        for n, v in zip(a["demo_lds_node"], a["demo_lds_val"]):
            f.write(f"{n:12d} {v[0]: .8E} {v[1]: .8E} {v[2]: .8E}\n")
    np.savez(os.path.join(d, "p121_demo_displ.npz"), displ=a["demo_displ"])
    for name in ("p121_demo_digests.json", "p121_demo_ensi_head.txt"):
        shutil.copy(os.path.join(PACKED, name), os.path.join(d, name))
    return d


@pytest.fixture(scope="session")
def tiny(golden):
    from parafem_b200 import host
    return host.read_deck_p121(os.path.join(golden, "xx3-tiny"))


@pytest.fixture(scope="session")
def tiny_xx2(golden):
    from parafem_b200 import host
    return host.read_deck_xx2(os.path.join(golden, "xx2-tiny"))


@pytest.fixture(scope="session")
def demo():
    from parafem_b200 import host
    return host.cube_p121(20, 20, 20, 20, aa=.5, bb=.5, cc=.5, round_mode=1)
