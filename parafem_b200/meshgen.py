"""p12meshgen (tools/preprocessing/p12meshgen/p12meshgen.f90) for the programs this library serves: reads <job>.mg
and writes the ParaFEM input deck <job>.d / .bnd / .lds / .dat (/ .mat) in the reference's formats.

  python -m parafem_b200.meshgen <job>          # <job>.mg -> deck, iotype 'parafem'
  python -m parafem_b200.meshgen --bin <job>    # p12meshgenbin: the deck plus <job>.bin.ensi.geo (binary geometry)

.mg layouts (tokens may be spread over lines arbitrarily):
  'p121' / iotype nels nxe nze nod nip / aa bb cc e v / tol limit                       (p12meshgen.f90:118-127)
  'p123' / iotype nels nxe nze nip / aa bb cc kx ky kz / tol limit / loaded fixed        (:660-663)
  'p124' / iotype nels nxe nze nip / aa bb cc kx ky kz rho cp / dtim nstep theta /
           npri tol limit val0 / np_types loaded fixed                                  (:829-833)
  'p125' / iotype nels nxe nze nip / aa bb cc kx ky kz / dtim nstep / npri val0          (:1019-1022)
"""
import ctypes as C
import sys

import numpy as np

from . import host
from ._lib import PfError, c_i64, check, lib, ptr


def read_mg(path):
    tok = open(path).read().replace(",", " ").split()
    prog = tok[0].strip("'" + '"').lower()
    iotype = tok[1].strip("'" + '"').lower()
    return prog, iotype, [float(t.lower().replace("d", "e")) for t in tok[2:]]


def _global_coords(p):
    g = np.zeros((p.nn, 3))
    g[p.g_num_pp - 1] = np.transpose(p.g_coord_pp, (0, 2, 1))
    return g


def generate(job, binary=False):
    """<job>.mg -> deck files; returns the host.Problem that was written.  binary: also <job>.bin.ensi.geo, as
    tools/preprocessing/p12meshgenbin does (mesh_ensi_geo_bin)."""
    p = _generate(job)
    if binary:
        host.write_geo_bin(job, _global_coords(p), p.g_num_pp)
    return p


def _generate(job):
    prog, iotype, v = read_mg(job + ".mg")
    if iotype != "parafem":
        raise PfError(f"iotype '{iotype}': only 'parafem' decks are written (the 'paraview' branch is a viewer export)")
    if prog not in ("p121", "p123", "p124", "p125"):
        raise PfError(f"program '{prog}' is not one of p121, p123, p124, p125")
    L = lib()
    if prog == "p121":
        nels, nxe, nze, nod, nip = (int(x) for x in v[:5])
        aa, bb, cc, e, nu, tol, limit = v[5:12]
        nye = nels // nxe // nze
        p = host.cube_p121(nxe, nye, nze, nod, aa=aa, bb=bb, cc=cc, e=e, v=nu, tol=tol, limit=int(limit), nip=nip)
        rest = np.zeros((4, p.nr), np.int32)
        check(L.pf_cube_rest(0, nxe, nye, nze, nod, p.nr, ptr(rest)), what="pf_cube_rest")
        nn, nr, loaded = c_i64(), c_i64(), c_i64()
        check(L.pf_p121_sizes(nxe, nye, nze, nod, C.byref(nn), C.byref(nr), C.byref(loaded)), what="pf_p121_sizes")
        node = np.empty(loaded.value, np.int32)
        val = np.empty((loaded.value, 3))
        check(L.pf_p121_loads(nxe, nze, nod, aa, bb, 0, ptr(node), ptr(val)), what="pf_p121_loads")
        host.write_deck_p121(job, nod, nip, e, nu, tol, int(limit), _global_coords(p), p.g_num_pp, rest, node, val)
        return p
    nels, nxe, nze, nip = (int(x) for x in v[:4])
    nye = nels // nxe // nze
    aa, bb, cc, kx, ky, kz = v[4:10]
    if prog == "p123":
        tol, limit, loaded, fixed = v[10], int(v[11]), int(v[12]), int(v[13])
        p = host.cube_p123(nxe, nye, nze, aa, bb, cc, kx, ky, kz, tol, limit, nip, fixed=fixed > 0)
    elif prog == "p124":
        rho, cp, dtim, nstep, theta, npri, tol, limit, val0 = v[10], v[11], v[12], int(v[13]), v[14], int(v[15]), v[16], int(v[17]), v[18]
        loaded, fixed = int(v[20]), int(v[21])
        theta = 0.5                                        # p12meshgen.f90:852 overrides the .mg value
        p = host.cube_p124(nxe, nye, nze, aa, bb, cc, kx, ky, kz, rho, cp, dtim, nstep, theta, npri, tol, limit, val0, nip)
    else:
        dtim, nstep, npri, val0 = v[10], int(v[11]), int(v[12]), v[13]
        loaded = fixed = 0
        p = host.cube_p125(nxe, nye, nze, aa, bb, cc, kx, ky, kz, dtim, nstep, npri, val0, nip)
    rest = np.zeros((2, p.nr), np.int32)
    check(L.pf_cube_rest(1, nxe, nye, nze, 8, p.nr, ptr(rest)), what="pf_cube_rest")
    host.write_deck_scalar(job, p, _global_coords(p), p.g_num_pp, rest, loaded=loaded, fixed=fixed)
    return p


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    binary = "--bin" in argv
    argv = [a for a in argv if a != "--bin"]
    if len(argv) != 1:
        print(__doc__)
        return 2
    p = generate(argv[0], binary=binary)
    print(f"{argv[0]}: program p{p.program}, {p.nels} elements, {p.nn} nodes, {p.nr} restrained, {p.neq} equations")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
