"""Host-side mirror of the ParaFEM library calls a p121 / p123 driver makes before
the device path (section B of include/parafem_b200.h, implemented in
csrc/host.cpp).  numpy arrays hold the Fortran arrays with axes reversed, i.e.
``g_num_pp(nod, nels_pp)`` is a C-contiguous ``(nels_pp, nod)`` array, so the raw
memory is exactly what the Fortran driver would pass.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._lib import DeckInfo, PfError, c_i64, check, f64, i32, lib, ptr


def calc_nels_pp(nels, npes, numpe, psize=None):
    """calc_nels_pp (gather_scatter.f90:146-257) -> (nels_pp, iel_start).  Partitioner 1 (internal,
    :217-238) unless ``psize`` -- the element counts per rank of a .psize file (partitioner 2,
    read_nels_pp input.f90:3108-3196) -- is given."""
    if psize is not None:
        psize = [int(c) for c in psize]
        if len(psize) != npes or sum(psize) != nels or min(psize) < 1:
            raise PfError(f"psize {psize} does not partition {nels} elements over {npes} ranks")
        return psize[numpe - 1], 1 + sum(psize[:numpe - 1])
    a, b = c_i64(), c_i64()
    lib().pf_calc_nels_pp(nels, npes, numpe, C.byref(a), C.byref(b))
    return a.value, b.value


def read_psize(job, npes, numpe):
    """read_nels_pp (input.f90:3108-3196): this rank's (nels_pp, iel_start) from <job>.psize."""
    a, b = c_i64(), c_i64()
    check(lib().pf_read_psize(job.encode(), npes, numpe, C.byref(a), C.byref(b)), what="pf_read_psize")
    return a.value, b.value


def calc_neq_pp(neq, npes, numpe):
    """calc_neq_pp (gather_scatter.f90:319-339) -> (neq_pp, ieq_start)."""
    a, b = c_i64(), c_i64()
    lib().pf_calc_neq_pp(neq, npes, numpe, C.byref(a), C.byref(b))
    return a.value, b.value


@dataclass
class Problem:
    """Everything one rank of a p121 / p123 run holds after make_ggl (p121.f90:49)."""
    program: int
    nod: int
    nodof: int
    nip: int
    nels: int
    nn: int
    nr: int
    neq: int
    npes: int
    numpe: int
    nels_pp: int
    iel_start: int
    neq_pp: int
    ieq_start: int
    g_num_pp: np.ndarray      # (nels_pp, nod) int32, S&G node order
    g_coord_pp: np.ndarray    # (nels_pp, 3, nod) float64
    g_g_pp: np.ndarray        # (nels_pp, ntot) int32, 0 = restrained
    nf: np.ndarray            # (nn, nodof) int32
    r_pp: np.ndarray          # (neq_pp,) starting residual (loads)
    e: float = 0.0
    v: float = 0.0
    kx: float = 0.0
    ky: float = 0.0
    kz: float = 0.0
    tol: float = 1e-5
    limit: int = 2000
    nres: int = 0
    no_f: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # fixed eq (global)
    val_f: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    total_load: float = 0.0
    # xx2: per-element materials -- prop (np_types, 2) = (e, v), etype_pp (nels_pp) 1-based; None = one material
    prop: np.ndarray = None
    etype_pp: np.ndarray = None
    # p124 (transient conduction): rho, cp, theta method, time stepping, initial value, print interval
    rho: float = 1.0
    cp: float = 1.0
    theta: float = 0.5
    dtim: float = 0.01
    nstep: int = 0
    npri: int = 1
    val0: float = 0.0

    @property
    def ntot(self):
        return self.nod * self.nodof


def _steer(nn, nodof, rest, g_num_pp, nod):
    nr = rest.shape[1]
    nf = np.empty((nn, nodof), np.int32)
    neq = c_i64()
    check(lib().pf_form_nf(nn, nodof, nr, ptr(rest), ptr(nf), C.byref(neq)), what="pf_form_nf")
    nels_pp = g_num_pp.shape[0]
    g_g = np.empty((nels_pp, nod * nodof), np.int32)
    check(lib().pf_find_g(nod, nodof, nels_pp, nn, ptr(g_num_pp), ptr(nf), ptr(g_g)), what="pf_find_g")
    return nf, g_g, neq.value


def cube_p121(nxe, nye, nze, nod=20, aa=None, bb=None, cc=None, e=100.0, v=0.3, tol=1e-5,
              limit=2000, nip=8, npes=1, numpe=1, round_mode=0, distort=0.0, seed=12345, psize=None):
    """In-memory p12meshgen cube for p121 (p12meshgen.f90:118-236): this rank's share.

    ``distort`` > 0 perturbs interior node coordinates by up to distort*h (uniform,
    deterministic per node number) -- not in the reference, kernel parity only.
    """
    aa = 10.0 / nxe if aa is None else aa
    bb = 10.0 / nye if bb is None else bb
    cc = 10.0 / nze if cc is None else cc
    L = lib()
    nn, nr, loaded = c_i64(), c_i64(), c_i64()
    check(L.pf_p121_sizes(nxe, nye, nze, nod, C.byref(nn), C.byref(nr), C.byref(loaded)), what="pf_p121_sizes")
    nn, nr, loaded = nn.value, nr.value, loaded.value
    nels = nxe * nye * nze
    nels_pp, iel_start = calc_nels_pp(nels, npes, numpe, psize)
    g_num = np.empty((nels_pp, nod), np.int32)
    g_coord = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_cube_elements(nxe, nze, nod, aa, bb, cc, iel_start, nels_pp, round_mode, ptr(g_num), ptr(g_coord)),
          what="pf_cube_elements")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_cube_rest(0, nxe, nye, nze, nod, nr, ptr(rest)), what="pf_cube_rest")
    if distort > 0.0:
        g_coord = _distort(g_coord, g_num, rest, nn, distort * min(aa, bb, cc), seed)
    nf, g_g, neq = _steer(nn, 3, rest, g_num, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    node = np.empty(loaded, np.int32)
    val = np.empty((loaded, 3), np.float64)
    check(L.pf_p121_loads(nxe, nze, nod, aa, bb, round_mode, ptr(node), ptr(val)), what="pf_p121_loads")
    r = np.empty(neq_pp, np.float64)
    check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
    return Problem(121, nod, 3, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                   g_num, g_coord, g_g, nf, r, e=e, v=v, tol=tol, limit=limit,
                   total_load=float(val[:, 2].sum()))


def _distort(g_coord, g_num, rest, nn, amp, seed):
    """Deterministic per-node jitter; nodes listed in rest (boundary) stay put so BCs hold."""
    rng = np.random.RandomState(seed)
    jit = (rng.rand(nn, 3) * 2.0 - 1.0) * amp
    jit[rest[0] - 1] = 0.0
    out = g_coord.copy()
    out += np.transpose(jit[g_num - 1], (0, 2, 1))
    return out


def cube_p123(nxe, nye, nze, aa=None, bb=None, cc=None, kx=2.0, ky=2.0, kz=2.0, tol=1e-5, limit=500,
              nip=8, npes=1, numpe=1, round_mode=0, source=10.0, fixed=False, fixed_value=100.0, psize=None):
    """In-memory p12meshgen box for p123 (p12meshgen.f90:658-701)."""
    aa = 1.0 / nxe if aa is None else aa
    bb = 1.0 / nye if bb is None else bb
    cc = 1.0 / nze if cc is None else cc
    L = lib()
    nn, nr, nres = c_i64(), c_i64(), c_i64()
    check(L.pf_p123_sizes(nxe, nye, nze, C.byref(nn), C.byref(nr), C.byref(nres)), what="pf_p123_sizes")
    nn, nr, nres = nn.value, nr.value, nres.value
    nels = nxe * nye * nze
    nels_pp, iel_start = calc_nels_pp(nels, npes, numpe, psize)
    g_num = np.empty((nels_pp, 8), np.int32)
    g_coord = np.empty((nels_pp, 3, 8), np.float64)
    check(L.pf_cube_elements(nxe, nze, 8, aa, bb, cc, iel_start, nels_pp, round_mode, ptr(g_num), ptr(g_coord)),
          what="pf_cube_elements")
    rest = np.zeros((2, nr), np.int32)
    check(L.pf_cube_rest(1, nxe, nye, nze, 8, nr, ptr(rest)), what="pf_cube_rest")
    nf, g_g, neq = _steer(nn, 1, rest, g_num, 8)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    r = np.zeros(neq_pp, np.float64)
    no_f = np.zeros(0, np.int32)
    val_f = np.zeros(0, np.float64)
    total = 0.0
    if fixed:
        # p12meshgen's fixed_freedoms branch (p12meshgen.f90:770-781): freedom nres held at a value
        if ieq_start <= nres < ieq_start + neq_pp:
            no_f = np.array([nres], np.int32)
            val_f = np.array([fixed_value], np.float64)
    else:
        # the .lds column is used directly as a global equation number (p123.f90:111-116)
        if ieq_start <= nres < ieq_start + neq_pp:
            r[nres - ieq_start] = source
        total = source
    return Problem(123, 8, 1, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                   g_num, g_coord, g_g, nf, r, kx=kx, ky=ky, kz=kz, tol=tol, limit=limit, nres=nres,
                   no_f=no_f, val_f=val_f, total_load=total)


def cube_p124(nxe, nye, nze, aa=None, bb=None, cc=None, kx=1.0, ky=1.0, kz=1.0, rho=1.0, cp=1.0, dtim=0.01,
              nstep=150, theta=0.5, npri=10, tol=1e-4, limit=100, val0=100.0, nip=8, npes=1, numpe=1,
              round_mode=0, fixed=False, fixed_value=100.0, psize=None):
    """In-memory p12meshgen box for p124 (p12meshgen.f90:827-905): the p123 box (geometry_8bxz, box_bc8,
    nres) with the transient data of the .mg file (defaults = the shipped p124_*.mg)."""
    p = cube_p123(nxe, nye, nze, aa, bb, cc, kx, ky, kz, tol, limit, nip, npes, numpe, round_mode, 0.0, fixed,
                  fixed_value, psize)
    p.program = 124
    p.r_pp[:] = 0.0
    p.rho, p.cp, p.theta, p.dtim, p.nstep, p.npri, p.val0 = rho, cp, theta, dtim, nstep, npri, val0
    return p


def cube_p125(nxe, nye, nze, aa=None, bb=None, cc=None, kx=1.0, ky=1.0, kz=1.0, dtim=2e-4, nstep=5000, npri=500,
              val0=100.0, nip=8, npes=1, numpe=1, round_mode=0, psize=None):
    """In-memory p12meshgen box for p125 (p12meshgen.f90:1017-1170): the p123 box with the explicit time
    stepping data of the .mg file (defaults = the shipped p125_*.mg)."""
    p = cube_p123(nxe, nye, nze, aa, bb, cc, kx, ky, kz, 0.0, 0, nip, npes, numpe, round_mode, 0.0, False, 0.0, psize)
    p.program = 125
    p.r_pp[:] = 0.0
    p.dtim, p.nstep, p.npri, p.val0 = dtim, nstep, npri, val0
    return p


def write_geo_bin(job, g_coord, g_num_sg):
    """mesh_ensi_geo_bin (input.f90:7986-8164), what p12meshgenbin writes beside the ASCII deck: <job>.bin.ensi.geo,
    EnSight Gold "C Binary" geometry -- single-precision coordinates, connectivity in EnSight's node order."""
    gc, gn = f64(g_coord), i32(g_num_sg)
    check(lib().pf_write_geo_bin(job.encode(), gn.shape[1], gc.shape[0], gn.shape[0], ptr(gc), ptr(gn)), what="pf_write_geo_bin")


def _read_mesh(job, nn, nels, nod, meshgen, binary):
    """g_coord (nn,3), g_num (nels,nod) in S&G order: read_g_num_pp + abaqus2sg + read_g_coord_pp from <job>.d, or
    their _be variants (input.f90:632-790, 1254-1420) from <job>.bin.ensi.geo when ``binary``."""
    L = lib()
    g_coord = np.empty((nn, 3), np.float64)
    g_num = np.empty((nels, nod), np.int32)
    if binary:
        check(L.pf_read_geo_bin(job.encode(), nn, nels, nod, ptr(g_coord), ptr(g_num)), what="pf_read_geo_bin")
        check(L.pf_ensi2sg(nod, nels, ptr(g_num)), what="pf_ensi2sg")     # for 8-node bricks == abaqus2sg (xx12.f90:118-121)
    else:
        check(L.pf_read_d(job.encode(), nn, nels, nod, ptr(g_coord), ptr(g_num)), what="pf_read_d")
        if meshgen == 2:
            check(L.pf_abaqus2sg(nod, nels, ptr(g_num)), what="pf_abaqus2sg")
    return g_coord, g_num


def read_deck_p121(job, npes=1, numpe=1, binary=False):
    """read_p121 + read_g_num_pp + abaqus2sg + read_g_coord_pp + read_rest + steering +
    read_loads + load (p121.f90:28-49, 79-85) for one rank.  binary: mesh from <job>.bin.ensi.geo."""
    L = lib()
    info = DeckInfo()
    check(L.pf_read_dat(job.encode(), 121, C.byref(info)), what="pf_read_dat")
    nod, nn, nels, nr, loaded = info.nod, info.nn, info.nels, info.nr, info.loaded
    g_coord, g_num = _read_mesh(job, nn, nels, nod, info.meshgen, binary)
    if info.partitioner == 2:      # external partition: <job>.psize, elements pre-sorted by rank
        nels_pp, iel_start = read_psize(job, npes, numpe)
    elif info.partitioner == 1:
        nels_pp, iel_start = calc_nels_pp(nels, npes, numpe)
    else:
        raise PfError("partitioner must be 1 (internal) or 2 (.psize)")
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_read_bnd(job.encode(), nr, 3, ptr(rest)), what="pf_read_bnd")
    nf, g_g, neq = _steer(nn, 3, rest, g_num_pp, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    node = np.empty(loaded, np.int32)
    val = np.empty((loaded, 3), np.float64)
    check(L.pf_read_lds(job.encode(), loaded, 3, ptr(node), ptr(val)), what="pf_read_lds")
    r = np.empty(neq_pp, np.float64)
    check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
    p = Problem(121, nod, 3, info.nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                g_num_pp, g_coord_pp, g_g, nf, r, e=info.e, v=info.v, tol=info.tol, limit=info.limit,
                total_load=float(val.sum()))
    p.rest = rest
    p.g_coord = g_coord
    return p


def read_deck_p124(job, npes=1, numpe=1):
    """Input section of p124.f90:27-49 (read_p124, read_elements, read_material): the p123 reader with the transient
    data of the .dat and the one material of the .mat."""
    p = read_deck_p123(job, npes, numpe, program=124)
    return p


def read_deck_p125(job, npes=1, numpe=1):
    """Input section of p125.f90:19-33 (read_p125)."""
    return read_deck_p123(job, npes, numpe, program=125)


def read_deck_p123(job, npes=1, numpe=1, program=123, binary=False):
    """Input section of p123.f90:27-55,94-131 (and of programs/dev/xx11/xx11.f90, which shares its deck format)
    for one rank: read_p123, read_g_num_pp, abaqus2sg, read_g_coord_pp, read_rest + rearrange_2/find_g4 -- or
    g_g_pp = g_num_pp when nr = 0 (p123.f90:54) --, read_loads (first column = global equation number,
    p123.f90:111-116) and read_fixed + find_no2 + reindex (node, sense -> this rank's fixed equations)."""
    L = lib()
    info = DeckInfo()
    check(L.pf_read_dat(job.encode(), program, C.byref(info)), what="pf_read_dat")
    nod, nn, nels, nr = info.nod, info.nn, info.nels, info.nr
    if nod not in (8, 4):
        raise PfError("p123 decks hold 8-node bricks (or, for xx11, 4-node tetrahedra)")
    g_coord, g_num = _read_mesh(job, nn, nels, nod, info.meshgen, binary)
    nels_pp, iel_start = read_psize(job, npes, numpe) if info.partitioner == 2 else calc_nels_pp(nels, npes, numpe)
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    if nr > 0:
        rest = np.zeros((2, nr), np.int32)
        check(L.pf_read_bnd(job.encode(), nr, 1, ptr(rest)), what="pf_read_bnd")
        nf, g_g, neq = _steer(nn, 1, rest, g_num_pp, nod)
    else:                                   # "When nr = 0, g_num_pp and g_g_pp are identical"
        nf = np.arange(1, nn + 1, dtype=np.int32).reshape(nn, 1)
        g_g, neq = g_num_pp.copy(), int(nn)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    r = np.zeros(neq_pp, np.float64)
    total = 0.0
    if info.loaded:
        eqn = np.empty(info.loaded, np.int32)
        val = np.empty((info.loaded, 1), np.float64)
        check(L.pf_read_lds(job.encode(), info.loaded, 1, ptr(eqn), ptr(val)), what="pf_read_lds")
        mine = (eqn >= ieq_start) & (eqn < ieq_start + neq_pp)
        r[eqn[mine] - ieq_start] = val[mine, 0]
        total = float(val.sum())
    no_f, val_f = np.zeros(0, np.int32), np.zeros(0, np.float64)
    if info.fixed:
        node = np.empty(info.fixed, np.int32)
        sense = np.empty(info.fixed, np.int32)
        valf = np.empty(info.fixed, np.float64)
        check(L.pf_read_fix(job.encode(), info.fixed, ptr(node), ptr(sense), ptr(valf)), what="pf_read_fix")
        if node.min() < 1 or node.max() > nn or sense.min() < 1 or sense.max() > nf.shape[1]:
            raise PfError(f"{job}.fix names a node outside 1..{nn} or a freedom outside 1..{nf.shape[1]}")
        if info.loaded and (eqn.min() < 1 or eqn.max() > neq):
            raise PfError(f"{job}.lds names an equation outside 1..{neq}")
        eq = nf[node - 1, sense - 1]        # find_no2 (loading.f90:742-828): the equation of (node, sense)
        mine = (eq >= ieq_start) & (eq < ieq_start + neq_pp)
        no_f, val_f = np.ascontiguousarray(eq[mine]), np.ascontiguousarray(valf[mine])
    p = Problem(program, nod, 1, info.nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                g_num_pp, g_coord_pp, g_g, nf, r, kx=info.kx, ky=info.ky, kz=info.kz, tol=info.tol,
                limit=info.limit, nres=info.nres, no_f=no_f, val_f=val_f, total_load=total)
    if program in (124, 125):
        p.dtim, p.nstep, p.npri, p.val0 = info.dtim, info.nstep, info.npri, info.val0
    if program == 124:
        p.theta = info.theta
        prop = np.empty((info.np_types, 5), np.float64)
        check(L.pf_read_mat(job.encode(), 5, info.np_types, ptr(prop)), what="pf_read_mat")
        p.kx, p.ky, p.kz, p.rho, p.cp = (float(v) for v in prop[0])      # one material (p12meshgen.f90:837)
    p.g_coord = g_coord
    return p


def read_deck_p122(job, npes=1, numpe=1):
    """Input section of p122.f90:36-57,97-114 for one rank: read_p122 (element, meshgen, partitioner / nels nn nr nip
    nod fixed_freedoms loaded_nodes / phi c psi e v / incs plasits cjits plastol cjtol), read_qinc, the mesh and
    restraint readers of p121, read_loads + load (ld0_pp) and read_fixed + find_no + reindex (this rank's fixed
    equations and their values).  -> Problem(program 122) with phi, c, psi, qinc, plasits, cjits, plastol, cjtol."""
    L = lib()
    tk = open(job + ".dat").read().split()
    if len(tk) < 20:
        raise PfError(f"{job}.dat: too few values for read_p122")
    meshgen, partitioner = int(tk[1]), int(tk[2])
    nels, nn, nr, nip, nod, fixed, loaded = (int(v) for v in tk[3:10])
    phi, c, psi, e, v = (float(t.replace("D", "E").replace("d", "e")) for t in tk[10:15])
    incs, plasits, cjits = (int(t) for t in tk[15:18])
    plastol, cjtol = float(tk[18]), float(tk[19])
    qinc = [float(t) for t in tk[20:20 + incs]]
    if min(nels, nn) < 1 or nr < 0 or nr > nn or nod not in (8, 20) or nip != 8 or len(qinc) != incs or fixed < 0 or loaded < 0:
        raise PfError(f"{job}.dat: sizes outside what p122 takes (hexahedra with 8 or 20 nodes, nip = 8)")
    g_coord = np.empty((nn, 3), np.float64)
    g_num = np.empty((nels, nod), np.int32)
    check(L.pf_read_d(job.encode(), nn, nels, nod, ptr(g_coord), ptr(g_num)), what="pf_read_d")
    if meshgen == 2:
        check(L.pf_abaqus2sg(nod, nels, ptr(g_num)), what="pf_abaqus2sg")
    nels_pp, iel_start = read_psize(job, npes, numpe) if partitioner == 2 else calc_nels_pp(nels, npes, numpe)
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_read_bnd(job.encode(), nr, 3, ptr(rest)), what="pf_read_bnd")
    nf, g_g, neq = _steer(nn, 3, rest, g_num_pp, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    r = np.zeros(neq_pp, np.float64)
    if loaded:
        node = np.empty(loaded, np.int32)
        val = np.empty((loaded, 3), np.float64)
        check(L.pf_read_lds(job.encode(), loaded, 3, ptr(node), ptr(val)), what="pf_read_lds")
        check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
    no_f, val_f = np.zeros(0, np.int32), np.zeros(0, np.float64)
    if fixed:
        node = np.empty(fixed, np.int32)
        sense = np.empty(fixed, np.int32)
        valf = np.empty(fixed, np.float64)
        check(L.pf_read_fix(job.encode(), fixed, ptr(node), ptr(sense), ptr(valf)), what="pf_read_fix")
        if node.min() < 1 or node.max() > nn or sense.min() < 1 or sense.max() > 3:
            raise PfError(f"{job}.fix names a node outside 1..{nn} or a freedom outside 1..3")
        eq = nf[node - 1, sense - 1]            # find_no (new_library.f90): the equation of (node, sense)
        mine = (eq >= ieq_start) & (eq < ieq_start + neq_pp)
        no_f, val_f = np.ascontiguousarray(eq[mine]), np.ascontiguousarray(valf[mine])
    p = Problem(122, nod, 3, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                g_num_pp, g_coord_pp, g_g, nf, r, e=e, v=v, tol=cjtol, limit=cjits, no_f=no_f, val_f=val_f)
    p.phi, p.c, p.psi, p.qinc, p.plasits, p.cjits, p.plastol, p.cjtol = phi, c, psi, qinc, plasits, cjits, plastol, cjtol
    p.loaded_nodes = loaded
    p.g_coord, p.rest = g_coord, rest
    return p


def cube_p129(nxe, nye, nze, aa, bb, cc, rho=2000.0, e=1.0e5, v=0.3, alpha1=0.0008, beta1=0.5, nstep=40, npri=1, theta=1.0,
              omega=0.01, tol=1e-4, limit=3000, nip=27, npes=1, numpe=1):
    """In-memory p12meshgen cantilever for p129 (p12meshgen.f90 CASE('p129')): 20-node bricks, the nodes of the plane
    y = 0 (the first nr node numbers) fully fixed, 2*nxe+1 loaded nodes at the far end, nres the monitored equation."""
    L = lib()
    nels = nxe * nye * nze
    nr = 3 * nxe * nze + 2 * nxe + 2 * nze + 1
    nn = ((2 * nxe + 1) * (nze + 1) + (nxe + 1) * nze) * (nye + 1) + (nxe + 1) * (nze + 1) * nye
    nels_pp, iel_start = calc_nels_pp(nels, npes, numpe)
    g_num = np.empty((nels_pp, 20), np.int32)
    g_coord = np.empty((nels_pp, 3, 20), np.float64)
    check(L.pf_cube_elements(nxe, nze, 20, aa, bb, cc, iel_start, nels_pp, 0, ptr(g_num), ptr(g_coord)), what="pf_cube_elements")
    rest = np.zeros((4, nr), np.int32)
    rest[0] = np.arange(1, nr + 1)
    nf, g_g, neq = _steer(nn, 3, rest, g_num, 20)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    loaded = 2 * nxe + 1
    k = np.arange(1, loaded + 1)
    node = (nn - loaded + k).astype(np.int32)
    val = np.zeros((loaded, 3))
    val[:, 2] = np.where((k == 1) | (k == loaded), 25.0 / 12.0, np.where(k % 2 == 0, 25.0 / 3.0, 25.0 / 6.0))
    r = np.empty(neq_pp, np.float64)
    check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
    p = Problem(129, 20, 3, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start, g_num, g_coord, g_g,
                nf, r, e=e, v=v, tol=tol, limit=limit, rho=rho, theta=theta, nstep=nstep, npri=npri,
                nres=3 * (nye * (nxe + 1) * (nze + 1) + nr * (nye - 1) + (nxe + 1)), total_load=float(val.sum()))
    p.alpha1, p.beta1, p.omega, p.rest = alpha1, beta1, omega, rest
    return p


def read_deck_p129(job, npes=1, numpe=1):
    """Input section of p129.f90:27-44,106-111 for one rank: read_p129 (input.f90:4968-4970: element, mesh, partitioner,
    nels nn nr nip nod loaded_nodes nres / rho e v alpha1 beta1 / nstep npri theta omega tol limit), the mesh and
    restraint readers of p121, read_loads + load (fext_pp)."""
    L = lib()
    tk = open(job + ".dat").read().split()
    if len(tk) < 21:
        raise PfError(f"{job}.dat: too few values for read_p129")
    meshgen, partitioner = int(tk[1]), int(tk[2])
    nels, nn, nr, nip, nod, loaded, nres = (int(v) for v in tk[3:10])
    rho, e, v, alpha1, beta1 = (float(t.replace("D", "E").replace("d", "e")) for t in tk[10:15])
    nstep, npri = int(tk[15]), int(tk[16])
    theta, omega, tol = (float(t.replace("D", "E").replace("d", "e")) for t in tk[17:20])
    limit = int(tk[20])
    if min(nels, nn) < 1 or nr < 0 or nr > nn or nod != 20 or nip not in (8, 27) or loaded < 0 or loaded > nn:
        raise PfError(f"{job}.dat: sizes outside what p129 takes (20-node hexahedra, nip = 8 or 27)")
    g_coord, g_num = _read_mesh(job, nn, nels, nod, meshgen, False)
    nels_pp, iel_start = read_psize(job, npes, numpe) if partitioner == 2 else calc_nels_pp(nels, npes, numpe)
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_read_bnd(job.encode(), nr, 3, ptr(rest)), what="pf_read_bnd")
    nf, g_g, neq = _steer(nn, 3, rest, g_num_pp, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    r = np.zeros(neq_pp, np.float64)
    total = 0.0
    if loaded:
        node = np.empty(loaded, np.int32)
        val = np.empty((loaded, 3), np.float64)
        check(L.pf_read_lds(job.encode(), loaded, 3, ptr(node), ptr(val)), what="pf_read_lds")
        check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
        total = float(val.sum())
    p = Problem(129, nod, 3, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start, g_num_pp, g_coord_pp,
                g_g, nf, r, e=e, v=v, tol=tol, limit=limit, rho=rho, theta=theta, nstep=nstep, npri=npri, nres=nres,
                total_load=total)
    p.alpha1, p.beta1, p.omega, p.rest, p.g_coord = alpha1, beta1, omega, rest, g_coord
    return p


def cube_p1210(nxe, nye, nze, aa=1.0, bb=1.0, cc=1.0, e=100.0, v=0.3, rho=1.0, sbary=4.0, dtim=2.0e-3, pload=1.0,
               nstep=240, npri=80, npes=1, numpe=1, form=0):
    """A p1210 problem on p12meshgen's p121 cube of 20-node bricks (restrained sides and base, 100 units of load on the
    top patch): the reference ships only the five-element p1210_tiny deck, this is the synthetic workload for sizes
    beyond it.  Defaults: inside the explicit stability limit for unit bricks, and a yield stress the Gauss points under
    the load exceed within the first hundred steps.  form: pf_vm_explicit_set_form."""
    p = cube_p121(nxe, nye, nze, 20, aa=aa, bb=bb, cc=cc, e=e, v=v, npes=npes, numpe=numpe)
    p.program, p.rho, p.sbary, p.dtim, p.pload, p.nstep, p.npri, p.nres, p.form = 1210, rho, sbary, dtim, pload, nstep, npri, 1, form
    return p


def read_deck_p1210(job, npes=1, numpe=1):
    """Input section of p1210.f90:27-52,106-111 for one rank.  read_p1210 (input.f90:5107-5109) reads
    element, meshgen, partitioner, nels nip nn nr nod loaded_nodes nres, rho e v sbary, dtim nstep npri pload;
    the one deck the reference ships (examples/5th_ed/p1210/p1210_tiny.dat, written for the program's 2010 form
    "p1210_5") has no nres and ends `pload dtim nstep npri` with the counts written as reals (3.e+5 3.e+3): both
    layouts are taken, told apart by the token after loaded_nodes (an integer nres or the real rho)."""
    L = lib()
    tk = open(job + ".dat").read().split()
    num = lambda t: float(t.replace("D", "E").replace("d", "e"))
    meshgen, partitioner = int(tk[1]), int(tk[2])
    nels, nip, nn, nr, nod, loaded = (int(v) for v in tk[3:9])
    if len(tk) >= 18 and tk[9].lstrip("+-").isdigit():
        nres = int(tk[9])
        rho, e, v, sbary, dtim = (num(t) for t in tk[10:15])
        nstep, npri, pload = int(tk[15]), int(tk[16]), num(tk[17])
    elif len(tk) >= 17:
        nres = 1                                               # p1210.f90:13 default
        rho, e, v, sbary, pload, dtim = (num(t) for t in tk[9:15])
        nstep, npri = int(round(num(tk[15]))), int(round(num(tk[16])))
    else:
        raise PfError(f"{job}.dat: too few values for read_p1210")
    if min(nels, nn) < 1 or nr < 0 or nr > nn or nod != 20 or nip != 8 or loaded < 0 or loaded > nn or nstep < 0 or npri < 1:
        raise PfError(f"{job}.dat: sizes outside what p1210 takes (20-node hexahedra, nip = 8)")
    g_coord, g_num = _read_mesh(job, nn, nels, nod, meshgen, False)
    nels_pp, iel_start = read_psize(job, npes, numpe) if partitioner == 2 else calc_nels_pp(nels, npes, numpe)
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_read_bnd(job.encode(), nr, 3, ptr(rest)), what="pf_read_bnd")
    nf, g_g, neq = _steer(nn, 3, rest, g_num_pp, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    r = np.zeros(neq_pp, np.float64)
    total = 0.0
    if loaded:
        node = np.empty(loaded, np.int32)
        val = np.empty((loaded, 3), np.float64)
        check(L.pf_read_lds(job.encode(), loaded, 3, ptr(node), ptr(val)), what="pf_read_lds")
        check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
        total = float(val.sum())
    p = Problem(1210, nod, 3, nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start, g_num_pp, g_coord_pp,
                g_g, nf, r, e=e, v=v, rho=rho, nstep=nstep, npri=npri, nres=nres, total_load=total)
    p.sbary, p.dtim, p.pload, p.rest, p.g_coord = sbary, dtim, pload, rest, g_coord
    return p


def read_deck_xx2(job, npes=1, numpe=1):
    """Input section of programs/dev/xx2/xx2.f90:60-160 for one rank: read_xx2, read_elements (connectivity +
    material number of every element), abaqus2sg, read_g_coord_pp, read_rest, read_materialValue, steering,
    read_loads + load.  -> Problem(program 121) with prop / etype_pp set."""
    L = lib()
    info = DeckInfo()
    check(L.pf_read_dat(job.encode(), 2, C.byref(info)), what="pf_read_dat")
    nod, nn, nels, nr, loaded = info.nod, info.nn, info.nels, info.nr, info.loaded
    if info.fixed:
        raise PfError("xx2 decks with fixed_freedoms > 0: pass no_f / val_f to Solver.build_precon yourself")
    g_coord = np.empty((nn, 3), np.float64)
    g_num = np.empty((nels, nod), np.int32)
    etype = np.empty(nels, np.int32)
    check(L.pf_read_d_mat(job.encode(), nn, nels, nod, ptr(g_coord), ptr(g_num), ptr(etype)), what="pf_read_d_mat")
    if info.meshgen == 2:
        check(L.pf_abaqus2sg(nod, nels, ptr(g_num)), what="pf_abaqus2sg")
    nels_pp, iel_start = read_psize(job, npes, numpe) if info.partitioner == 2 else calc_nels_pp(nels, npes, numpe)
    g_num_pp = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_coord_pp = np.empty((nels_pp, 3, nod), np.float64)
    check(L.pf_coords_pp(nod, nels_pp, nn, ptr(g_num_pp), ptr(g_coord), ptr(g_coord_pp)), what="pf_coords_pp")
    rest = np.zeros((4, nr), np.int32)
    check(L.pf_read_bnd(job.encode(), nr, 3, ptr(rest)), what="pf_read_bnd")
    prop = np.empty((info.np_types, 2), np.float64)
    check(L.pf_read_mat(job.encode(), 2, info.np_types, ptr(prop)), what="pf_read_mat")
    nf, g_g, neq = _steer(nn, 3, rest, g_num_pp, nod)
    neq_pp, ieq_start = calc_neq_pp(neq, npes, numpe)
    node = np.empty(loaded, np.int32)
    val = np.empty((loaded, 3), np.float64)
    check(L.pf_read_lds(job.encode(), loaded, 3, ptr(node), ptr(val)), what="pf_read_lds")
    r = np.empty(neq_pp, np.float64)
    check(L.pf_load(3, loaded, nn, ptr(node), ptr(val), ptr(nf), ieq_start, neq_pp, ptr(r)), what="pf_load")
    p = Problem(121, nod, 3, info.nip, nels, nn, nr, neq, npes, numpe, nels_pp, iel_start, neq_pp, ieq_start,
                g_num_pp, g_coord_pp, g_g, nf, r, tol=info.tol, limit=info.limit, total_load=float(val.sum()))
    p.prop = prop
    p.etype_pp = np.ascontiguousarray(etype[iel_start - 1:iel_start - 1 + nels_pp])
    p.rest, p.g_coord = rest, g_coord
    return p


def write_deck_scalar(job, prob, g_coord, g_num, rest, loaded=0, fixed=0):
    """p12meshgen's output side for p123 / p124 / p125 (prob.program): <job>.d/.bnd/.dat (+ .mat for p124)."""
    info = DeckInfo()
    info.program, info.meshgen, info.partitioner, info.nip, info.nod, info.limit = prob.program, 2, 1, prob.nip, 8, prob.limit
    info.nels, info.nn, info.nr, info.loaded, info.fixed, info.nres = prob.nels, prob.nn, prob.nr, loaded, fixed, prob.nres
    info.kx, info.ky, info.kz, info.tol = prob.kx, prob.ky, prob.kz, prob.tol
    info.np_types, info.nstep, info.npri = 1, prob.nstep, prob.npri
    info.val0, info.dtim, info.theta, info.rho, info.cp = prob.val0, prob.dtim, prob.theta, prob.rho, prob.cp
    g_coord, g_num, rest = f64(g_coord), i32(g_num), i32(rest)
    check(lib().pf_write_deck_scalar(str(job).encode(), C.byref(info), ptr(g_coord), ptr(g_num), ptr(rest)),
          what="pf_write_deck_scalar")


def make_ggl(prob):
    """Gather tables of one rank (pf_make_ggl): (ggl_pp, halo_eq, halo_cnt)."""
    L = lib()
    nh = c_i64()
    cnt = np.zeros(prob.npes, np.int64)
    check(L.pf_make_ggl(prob.ntot, prob.nels_pp, ptr(prob.g_g_pp), prob.neq, prob.npes, prob.numpe,
                        None, 0, None, ptr(cnt), C.byref(nh)), what="pf_make_ggl(size)")
    ggl = np.empty_like(prob.g_g_pp)
    halo = np.empty(max(nh.value, 1), np.int32)
    check(L.pf_make_ggl(prob.ntot, prob.nels_pp, ptr(prob.g_g_pp), prob.neq, prob.npes, prob.numpe,
                        ptr(ggl), halo.size, ptr(halo), ptr(cnt), C.byref(nh)), what="pf_make_ggl")
    return ggl, halo[:nh.value], cnt


def nodal_values(prob, x_pp):
    """Nodal field of this rank's node range (calc_nodes_pp + what scatter_nodes yields)."""
    L = lib()
    a, b = c_i64(), c_i64()
    L.pf_calc_nodes_pp(prob.nn, prob.npes, prob.numpe, C.byref(a), C.byref(b))
    out = np.zeros((a.value, prob.nodof))
    check(L.pf_nodal_values(prob.nodof, prob.nn, ptr(prob.nf), prob.ieq_start, prob.neq_pp, ptr(f64(x_pp)),
                            b.value, a.value, ptr(out)), what="pf_nodal_values")
    return out


def write_ensi(path, values, decimals=5):
    """dismsh_ensi_p's EnSight Gold ASCII file: values (nn, numvar), component-major on disk
    (numvar = 1 gets the "Scalar per-node" header p123/p124 write, p124.f90:183-186)."""
    v = f64(values)
    check(lib().pf_write_ensi(str(path).encode(), v.shape[1], v.shape[0], ptr(v), decimals), what="pf_write_ensi")


def write_deck_p121(job, nod, nip, e, v, tol, limit, g_coord, g_num, rest, node, val):
    """p12meshgen's output side: <job>.d/.bnd/.lds/.dat (g_num in S&G order, rest (4, nr))."""
    g_coord, g_num, rest, node, val = f64(g_coord), i32(g_num), i32(rest), i32(node), f64(val)
    check(lib().pf_write_deck_p121(str(job).encode(), nod, g_num.shape[0], g_coord.shape[0], rest.shape[1], nip,
                                   node.size, e, v, tol, limit, ptr(g_coord), ptr(g_num), ptr(rest), ptr(node),
                                   ptr(val)), what="pf_write_deck_p121")
