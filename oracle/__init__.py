"""ctypes wrapper of the CPU oracle (oracle/pf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by parafem_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpf_oracle.so")
_lib = None

vp, i64, cint, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double


def build(force=False):
    src = os.path.join(_HERE, "pf_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "CC=gcc"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


REF_SRC = "/root/reference/parafem/src/programs/dev/xx3/cuda_helpers.cu"
REF_DIR = os.path.join(_HERE, "_ref")


def build_ref():
    """oracle/_ref: the reference's own CUDA mat-vec (xx3/cuda_helpers.cu) compiled from where it lies -- only
    where /root/reference exists (this container); the GPU box uses the shipped binaries."""
    if os.path.exists(REF_SRC):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return ref_available()


def ref_tool(name="metout2pf"):
    """Path of a reference tool built into oracle/_ref (None when it was not built)."""
    path = os.path.join(REF_DIR, name)
    return path if os.path.exists(path) else None


def ref_available():
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in ("libxx3_cuda_helpers.so", "libxx3_cuda_helpers_nofma.so"))


def ref_lib(nofma=False):
    """ctypes handle of the reference's compiled cuda_helpers (RTLD_LOCAL: its symbols have the same names as the
    product's xx3-compatible entry points)."""
    name = "libxx3_cuda_helpers_nofma.so" if nofma else "libxx3_cuda_helpers.so"
    return C.CDLL(os.path.join(REF_DIR, name), mode=os.RTLD_LOCAL | os.RTLD_NOW)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_dot_blocked.restype = dbl
        L.orc_dot_blocked.argtypes = [vp, vp, i64]
        L.orc_dot_ranks.restype = dbl
        L.orc_dot_ranks.argtypes = [vp, vp, i64, cint, cint]
        L.orc_determinant3.restype = dbl
        L.orc_form_km_elastic.argtypes = [i64, cint, cint, vp, dbl, dbl, vp]
        L.orc_form_kc_laplace.argtypes = [i64, cint, cint, vp, dbl, dbl, dbl, vp]
        L.orc_centroid_stress.argtypes = [cint, vp, vp, dbl, dbl, vp]
        L.orc_point_stress.argtypes = [cint, vp, vp, dbl, dbl, dbl, dbl, dbl, vp]
        L.orc_gather.argtypes = [cint, i64, vp, vp, vp]
        L.orc_matvec.argtypes = [cint, i64, vp, vp, vp]
        L.orc_scatter.argtypes = [cint, i64, vp, i64, cint, vp, vp]
        L.orc_rearrange.argtypes = [i64, cint, vp]
        L.orc_rearrange_2.argtypes = [i64, vp]
        L.orc_find_g3.argtypes = [cint, cint, vp, vp, i64, vp]
        L.orc_find_g4.argtypes = [cint, vp, vp, i64, vp]
        L.orc_partition.argtypes = [i64, cint, cint, C.POINTER(i64), C.POINTER(i64)]
        L.orc_pcg.argtypes = [cint, i64, vp, vp, i64, vp, i64, vp, vp, dbl, cint, cint, dbl, cint, vp,
                              C.POINTER(cint), C.POINTER(cint), vp, vp, C.POINTER(dbl), vp, cint, cint, dbl, dbl]
        L.orc_apply_mf.argtypes = [i64, cint, cint, vp, dbl, dbl, vp, vp]
        L.orc_set_element_partition.argtypes = [vp, cint]
        L.orc_set_element_partition.restype = None
        L.orc_sample_hex.argtypes = [cint, vp, vp]
        L.orc_shape_der.argtypes = [cint, vp, cint, cint, vp]
        L.orc_set_threads.argtypes = [cint]
        L.orc_shape_fun.argtypes = [cint, vp, cint, cint, vp]
        L.orc_form_k_transient.argtypes = [i64, cint, cint, vp, dbl, dbl, dbl, dbl, dbl, dbl, dbl, vp, vp, vp, vp]
        L.orc_apply.argtypes = [cint, i64, vp, vp, i64, cint, vp, vp]
        L.orc_p122_elements.argtypes = [i64, cint, cint, vp, dbl, dbl, dbl, dbl, dbl, dbl, cint, vp, vp, vp, vp]
        L.orc_form_mass.argtypes = [i64, cint, cint, vp, dbl, vp]
        L.orc_cube_elements.argtypes = [cint, cint, cint, dbl, dbl, dbl, i64, i64, vp, vp]
        L.orc_cube_rest.argtypes = [cint, cint, cint, cint, i64, vp]
        L.orc_cube_rest.restype = i64
        L.orc_load_p121.argtypes = [cint, cint, cint, dbl, dbl, vp, vp]
        L.orc_load_p121.restype = i64
        L.orc_find_g_all.argtypes = [cint, cint, i64, vp, vp, i64, vp]
        L.orc_find_g_all.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_threads(n):
    lib().orc_set_threads(int(n))


def set_element_partition(counts=None):
    """Partitioner 2 (read_nels_pp, input.f90:3108-3196): elements per emulated rank from a .psize
    list instead of calc_nels_pp; None restores partitioner 1.  Applies when npes == len(counts)."""
    if counts is None:
        lib().orc_set_element_partition(None, 0)
    else:
        c = np.ascontiguousarray(counts, np.int64)
        lib().orc_set_element_partition(_p(c), len(c))


def max_threads():
    return lib().orc_max_threads()


def host_cores():
    """The cores this process may run on -- NOT the launcher's OMP_NUM_THREADS (torchrun sets it to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_cores():
    n = host_cores()
    set_threads(n)
    return n


def _round_sig(a, digits):
    """What a value becomes after a trip through a Fortran Ew.d field (d significant digits)."""
    u, inv = np.unique(a, return_inverse=True)
    r = np.array([float(f"{x:.{digits - 1}e}") for x in u])
    return r[inv].reshape(a.shape)


class Mesh:
    """What p121 / p123 hold after read_* + rearrange + find_g (global arrays, one rank)."""


def cube_p121(nxe, nye, nze, nod=20, aa=None, bb=None, cc=None, e=100.0, v=0.3, tol=1e-5, limit=2000, nip=8,
              deck_rounding=False):
    """p12meshgen's p121 cube (p12meshgen.f90:118-236) followed by p121.f90:30-45,79-84: connectivity,
    element coordinates, restraints -> steering array, loads -> starting residual.  Built from the oracle's
    own restatement of geometry_*bxz / cube_bc* / load_p121 / rearrange / find_g3 -- independent of the
    product's host library.  ``deck_rounding``: coordinates through E14.6 and loads through E16.8, as
    when the mesh travels through the deck files."""
    aa = 10.0 / nxe if aa is None else aa
    bb = 10.0 / nye if bb is None else bb
    cc = 10.0 / nze if cc is None else cc
    L = lib()
    nels = nxe * nye * nze
    if nod == 20:
        nr = (((2 * nxe + 1) * (nze + 1) + (nxe + 1) * nze) * 2 + ((2 * nye - 1) * nze + (nye - 1) * nze) * 2
              + (2 * nye - 1) * (nxe + 1) + (nye - 1) * nxe)
        nn = ((2 * nxe + 1) * (nze + 1) + (nxe + 1) * nze) * (nye + 1) + (nxe + 1) * (nze + 1) * nye
    else:
        nr = (nxe + 1) * (nze + 1) * 2 + (nye - 1) * (nze + 1) * 2 + (nxe - 1) * (nze - 1)
        nn = (nxe + 1) * (nye + 1) * (nze + 1)
    m = Mesh()
    m.program, m.nod, m.nodof, m.nip, m.nels, m.nn, m.nr = 121, nod, 3, nip, nels, nn, nr
    m.e, m.v, m.tol, m.limit = e, v, tol, limit
    m.g_num_pp = np.empty((nels, nod), np.int32)
    m.g_coord_pp = np.empty((nels, 3, nod))
    assert L.orc_cube_elements(nod, nxe, nze, aa, bb, cc, 0, nels, _p(m.g_num_pp), _p(m.g_coord_pp)) == 0
    if deck_rounding:
        m.g_coord_pp = _round_sig(m.g_coord_pp, 6)
    # nod = 8: p12meshgen.f90:178's nr equals the rows cube_bc8 emits only on nxe == nye == nze boxes; the
    # rows cube_bc8 emits are what a deck can hold, so they are counted (first call writes nothing)
    emitted = L.orc_cube_rest(nod, nxe, nye, nze, 0, None)
    assert emitted == nr or (nod == 8 and not nxe == nye == nze), (emitted, nr)
    m.nr = nr = emitted
    rest = np.zeros((4, nr), np.int32)
    assert L.orc_cube_rest(nod, nxe, nye, nze, nr, _p(rest)) == nr
    m.rest = rest.copy()
    L.orc_rearrange(nr, 3, _p(rest))
    m.g_g_pp = np.zeros((nels, 3 * nod), np.int32)
    L.orc_find_g_all(nod, 3, nels, _p(m.g_num_pp), _p(m.g_g_pp), nr, _p(rest))
    m.neq = int(m.g_g_pp.max())
    loaded = L.orc_load_p121(nod, nxe, nze, aa, bb, None, None)
    node = np.empty(loaded, np.int32)
    val = np.empty(loaded)
    assert L.orc_load_p121(nod, nxe, nze, aa, bb, _p(node), _p(val)) == loaded
    if deck_rounding:
        val = _round_sig(val, 8)
    # load() + scatter_noadd (loading.f90:60-140): the z-load of every loaded node lands on its equation
    nf_z = np.zeros(nn + 1, np.int32)
    nf_z[m.g_num_pp.ravel()] = m.g_g_pp.reshape(nels, nod, 3)[:, :, 2].ravel()
    m.r_pp = np.zeros(m.neq)
    eq = nf_z[node]
    m.r_pp[eq[eq > 0] - 1] = val[eq > 0]
    m.loaded_nodes, m.load_node, m.load_val = loaded, node, val
    m.total_load = float(val.sum())
    return m


def cube_p123(nxe, nye, nze, aa=None, bb=None, cc=None, kx=2.0, ky=2.0, kz=2.0, tol=1e-5, limit=500, nip=8,
              source=10.0):
    """p12meshgen's p123 box (p12meshgen.f90:658-701) followed by p123.f90's rearrange_2 / find_g4 and the
    loaded freedom nres (used directly as an equation number, p123.f90:111-116)."""
    aa = 1.0 / nxe if aa is None else aa
    bb = 1.0 / nye if bb is None else bb
    cc = 1.0 / nze if cc is None else cc
    L = lib()
    nels = nxe * nye * nze
    nr = (nxe + 1) * (nye + 1) + (nxe + 1) * nze + nye * nze
    nn = (nxe + 1) * (nye + 1) * (nze + 1)
    m = Mesh()
    m.program, m.nod, m.nodof, m.nip, m.nels, m.nn, m.nr = 123, 8, 1, nip, nels, nn, nr
    m.kx, m.ky, m.kz, m.tol, m.limit, m.nres = kx, ky, kz, tol, limit, nxe * (nze - 1) + 1
    m.g_num_pp = np.empty((nels, 8), np.int32)
    m.g_coord_pp = np.empty((nels, 3, 8))
    assert L.orc_cube_elements(8, nxe, nze, aa, bb, cc, 0, nels, _p(m.g_num_pp), _p(m.g_coord_pp)) == 0
    rest = np.zeros((2, nr), np.int32)
    assert L.orc_cube_rest(1, nxe, nye, nze, nr, _p(rest)) == nr
    L.orc_rearrange_2(nr, _p(rest))
    m.g_g_pp = np.zeros((nels, 8), np.int32)
    L.orc_find_g_all(8, 1, nels, _p(m.g_num_pp), _p(m.g_g_pp), nr, _p(rest))
    m.neq = int(m.g_g_pp.max())
    m.r_pp = np.zeros(m.neq)
    m.r_pp[m.nres - 1] = source
    m.total_load = source
    return m


def form_km_elastic(g_coord_pp, nod, nip, e, v):
    """g_coord_pp (nels,3,nod) -> storkm (nels, ntot, ntot) [Fortran storkm_pp(i,j,iel) = out[iel,j,i]]."""
    g = _f64(g_coord_pp)
    nels = g.shape[0]
    out = np.empty((nels, 3 * nod, 3 * nod))
    rc = lib().orc_form_km_elastic(nels, nod, nip, _p(g), e, v, _p(out))
    assert rc == 0
    return out


def form_km_elastic_mat(g_coord_pp, nod, nip, prop, etype):
    """xx2.f90:169-193: e, v = prop(:, etype(iel)) per element (prop (np_types, 2), etype 1-based)."""
    g, et = _f64(g_coord_pp), _i32(etype)
    out = np.empty((g.shape[0], 3 * nod, 3 * nod))
    for m in np.unique(et):
        sel = np.flatnonzero(et == m)
        out[sel] = form_km_elastic(g[sel], nod, nip, float(prop[m - 1][0]), float(prop[m - 1][1]))
    return out


def form_kc_laplace(g_coord_pp, nip, kx, ky, kz):
    """p123.f90:70-84; 8-node bricks or (nip = 1) 4-node tetrahedra, from the last axis of g_coord_pp."""
    g = _f64(g_coord_pp)
    nels, nod = g.shape[0], g.shape[2]
    out = np.empty((nels, nod, nod))
    rc = lib().orc_form_kc_laplace(nels, nod, nip, _p(g), kx, ky, kz, _p(out))
    assert rc == 0
    return out


def centroid_stress(nod, coord, eld, e, v):
    sig = np.empty(6)
    rc = lib().orc_centroid_stress(nod, _p(_f64(coord)), _p(_f64(eld)), e, v, _p(sig))
    assert rc == 0
    return sig


def point_stress(nod, coord, eld, e, v, xi, eta, zeta):
    sig = np.empty(6)
    rc = lib().orc_point_stress(nod, _p(_f64(coord)), _p(_f64(eld)), e, v, xi, eta, zeta, _p(sig))
    assert rc == 0
    return sig


def gather(g_g, p):
    g = _i32(g_g)
    out = np.empty(g.shape)
    lib().orc_gather(g.shape[1], g.shape[0], _p(g), _p(_f64(p)), _p(out))
    return out


def matvec(storkm, pmul):
    k = _f64(storkm)
    out = np.empty((k.shape[0], k.shape[1]))
    lib().orc_matvec(k.shape[1], k.shape[0], _p(k), _p(_f64(pmul)), _p(out))
    return out


def scatter(g_g, utemp, neq, npes=1):
    g = _i32(g_g)
    u = np.zeros(neq)
    lib().orc_scatter(g.shape[1], g.shape[0], _p(g), neq, npes, _p(_f64(utemp)), _p(u))
    return u


def dot_blocked(a, b):
    a, b = _f64(a), _f64(b)
    return lib().orc_dot_blocked(_p(a), _p(b), a.size)


def dot_ranks(a, b, npes=1, red_mode=1):
    a, b = _f64(a), _f64(b)
    return lib().orc_dot_ranks(_p(a), _p(b), a.size, npes, red_mode)


def find_g3(g_num, rest, nodof=3):
    """rearrange + find_g3 over all elements; rest (nodof+1, nr) as read from .bnd (copied)."""
    rest = _i32(rest).copy()
    nr = rest.shape[1]
    lib().orc_rearrange(nr, nodof, _p(rest))
    g_num = _i32(g_num)
    nels, nod = g_num.shape
    g = np.zeros((nels, nod * nodof), np.int32)
    for e in range(nels):
        lib().orc_find_g3(nod, nodof, _p(g_num[e]), _p(g[e]), nr, _p(rest))
    return g


def find_g4(g_num, rest):
    rest = _i32(rest).copy()
    nr = rest.shape[1]
    lib().orc_rearrange_2(nr, _p(rest))
    g_num = _i32(g_num)
    nels, nod = g_num.shape
    g = np.zeros((nels, nod), np.int32)
    for e in range(nels):
        lib().orc_find_g4(nod, _p(g_num[e]), _p(g[e]), nr, _p(rest))
    return g


def set_mf_order(order):
    """2: the summation order of k_apply_mf4 / k_apply_mf3 (FP64 tensor-core kernels, the default for both bricks: one
    fma chain in the k order of their mma instructions, deemat's structural zeros left out); 1: k_apply_mf2
    (PF_MF=2lane: two 12-term chains added); 0: k_apply_mf (PF_MF=1lane: one 24-term chain, Gauss points ascending)."""
    lib().orc_set_mf_order(int(order))


def default_mf_order(nod):
    """The order of the kernel the library launches for this element type (PF_MF in the environment selects the older
    kernels for A/B runs; the tests follow it)."""
    sel = os.environ.get("PF_MF", "")
    if sel == "1lane":
        return 0
    if sel == "2lane":
        return 1 if nod == 20 else 0
    return 2


def apply_mf(g_coord_pp, nod, nip, e, v, pmul, mode=2):
    """Matrix-free element products (config E): utemp = sum_gp B^T D B p det w in the operation order of the device
    kernel in use (k_apply_mf4 unless PF_MF selects an older one; both matrix-free modes of a
    kernel give the same bits, ``mode`` is accepted for symmetry with the device call)."""
    g, pm = _f64(g_coord_pp), _f64(pmul)
    out = np.empty(pm.shape)
    set_mf_order(default_mf_order(nod))
    rc = lib().orc_apply_mf(g.shape[0], nod, nip, _p(g), e, v, _p(pm), _p(out))
    assert rc == 0
    return out


def pcg(storkm, g_g, neq, r, tol, limit, npes=1, red_mode=0, no_f=None, val_f=None, penalty=1e20, mf=None):
    """p121.f90:65-104 / p123.f90:86-151 on global arrays.  Returns dict(x, iters, converged,
    ratio, diag, seconds)."""
    k, g, r = _f64(storkm), _i32(g_g), _f64(r)
    nels, ntot = g.shape
    nfixed = 0 if no_f is None else len(no_f)
    no_f = _i32(no_f) if nfixed else None
    val_f = _f64(val_f) if (nfixed and val_f is not None) else None
    x = np.empty(neq)
    diag = np.empty(neq)
    ratio = np.zeros(limit)
    it, conv, secs = cint(), cint(), dbl()
    mfc = _f64(mf["g_coord_pp"]) if mf else None     # mf = dict(g_coord_pp, nod, nip, e, v[, mode]): matrix-free products
    if mf:
        set_mf_order(default_mf_order(mf["nod"]))
    rc = lib().orc_pcg(ntot, nels, _p(g), _p(k), neq, _p(r), nfixed, _p(no_f), _p(val_f), penalty, npes,
                       red_mode, tol, limit, _p(x), C.byref(it), C.byref(conv), _p(ratio), _p(diag),
                       C.byref(secs), _p(mfc), mf["nod"] if mf else 0, mf["nip"] if mf else 0,
                       mf["e"] if mf else 0.0, mf["v"] if mf else 0.0)
    assert rc == 0
    return dict(x=x, iters=it.value, converged=bool(conv.value), ratio=ratio[:it.value], diag=diag,
                seconds=secs.value)


def form_k_transient(g_coord_pp, nip, kx, ky, kz, rho, cp, theta, dtim, raw=False):
    """p124.f90:81-95: (storka, storkb) for 8-node bricks; raw=True also returns (kc, pm)."""
    g = _f64(g_coord_pp)
    nels = g.shape[0]
    a, b = np.empty((nels, 8, 8)), np.empty((nels, 8, 8))
    kc = np.empty((nels, 8, 8)) if raw else None
    pm = np.empty((nels, 8, 8)) if raw else None
    rc = lib().orc_form_k_transient(nels, 8, nip, _p(g), kx, ky, kz, rho, cp, theta, dtim, _p(a), _p(b), _p(kc), _p(pm))
    assert rc == 0
    return (a, b, kc, pm) if raw else (a, b)


def apply(storkm, g_g, neq, x, npes=1):
    """u = scatter(MATMUL(storkm, gather(x))) over npes emulated ranks."""
    k, g = _f64(storkm), _i32(g_g)
    u = np.zeros(neq)
    lib().orc_apply(g.shape[1], g.shape[0], _p(g), _p(k), neq, npes, _p(_f64(x)), _p(u))
    return u


def p124(storka, storkb, g_g, neq, val0, nstep, tol, limit, npes=1, red_mode=0, loads=None, keep=(),
         no_f=None, val_f=None, penalty=1e20):
    """The time-stepping loop of p124.f90:139-232: per step the right-hand side loads + B*x (A*x0 on the
    first step), then PCG from x = 0 on storka.  `loads` = val*dtim at the loaded freedoms (neq) or None;
    no_f / val_f = fixed freedoms (global equation numbers, values), handled line by line as the reference
    writes them (:155-160, :170-173, :193-199, :209-212).  Returns dict(iters[nstep], x (last), fields
    {step: x} for the steps listed in `keep` (0 = the initial field))."""
    nfix = 0 if no_f is None else len(no_f)
    if nfix:
        no_f, val_f = _i32(no_f), _f64(val_f)
        ntot = storka.shape[1]
        diag_tmp = np.ascontiguousarray(storka[:, np.arange(ntot), np.arange(ntot)])
        store = scatter(g_g, diag_tmp, neq, npes)[no_f - 1] + penalty
    x = np.full(neq, float(val0))
    if nfix:
        x[no_f - 1] = val_f
    fields = {0: x.copy()} if 0 in keep else {}
    iters, conv = [], []
    for j in range(1, nstep + 1):
        rhs = np.zeros(neq) if loads is None else _f64(loads).copy()
        u = apply(storka if j == 1 else storkb, g_g, neq, x, npes)
        if nfix and j != 1:
            u[no_f - 1] = store * val_f
        rhs = rhs + u
        if nfix:
            rhs[no_f - 1] = rhs[no_f - 1] - store * val_f
        res = pcg(storka, g_g, neq, rhs, tol, limit, npes=npes, red_mode=red_mode, no_f=no_f if nfix else None,
                  val_f=None, penalty=penalty)
        x = res["x"]
        iters.append(res["iters"])
        conv.append(res["converged"])
        if j in keep:
            fields[j] = x.copy()
    return dict(iters=iters, converged=conv, x=x, fields=fields)


def form_mass(g_coord_pp, nod, nip, rho):
    """p129.f90:85-98, consistent mass: store_mm_pp (nels, ntot, ntot)."""
    g = _f64(g_coord_pp)
    out = np.empty((g.shape[0], 3 * nod, 3 * nod))
    assert lib().orc_form_mass(g.shape[0], nod, nip, _p(g), rho, _p(out)) == 0
    return out


def p129(km, mm, g_g, neq, fext, theta, omega, alpha1, beta1, nstep, tol, limit, npes=1, red_mode=1, keep=()):
    """The time-stepping loop of p129.f90:60-69,99-150 (forced vibration, theta method, consistent mass, Rayleigh
    damping alpha1 / beta1; dtim = period/20): per step the two right-hand-side products, the harmonic load, one PCG
    solve from x = 0 on store_mm_pp*c3 + store_km_pp*c4, then the velocity / acceleration updates.
    -> dict(rows [(time, cos(omega t), iters)], x (last), d1x, d2x, fields {step: x})."""
    import math
    pi = math.acos(-1.0)
    period = 2.0 * pi / omega
    dtim = period / 20.0
    c1 = (1.0 - theta) * dtim
    c2 = beta1 - c1
    c3 = alpha1 + 1.0 / (theta * dtim)
    c4 = beta1 + theta * dtim
    a_mat = mm * c3 + km * c4                 # temp_pp of the iterations loop (p129.f90:128)
    b_mat = km * c2 + mm * c3                 # temp_pp of elements_3 (:114)
    m_th = mm / theta                         # :120
    ntot = km.shape[1]
    diag_tmp = np.ascontiguousarray(mm[:, np.arange(ntot), np.arange(ntot)] * c3 + km[:, np.arange(ntot), np.arange(ntot)] * c4)
    x0 = np.zeros(neq); d1x0 = np.zeros(neq); d2x0 = np.zeros(neq)
    real_time, rows, fields = 0.0, [], {}
    for j in range(1, nstep + 1):
        real_time = real_time + dtim
        u = apply(b_mat, g_g, neq, x0, npes)
        vu = apply(m_th, g_g, neq, d1x0, npes)
        loads = fext * (theta * dtim * math.cos(omega * real_time) + c1 * math.cos(omega * (real_time - dtim)))
        loads = u + vu + loads
        res = pcg(a_mat, g_g, neq, loads, tol, limit, npes=npes, red_mode=red_mode)
        assert np.array_equal(res["diag"], 1.0 / scatter(g_g, diag_tmp, neq, npes))     # :99-105
        x1 = res["x"]
        d1x1 = (x1 - x0) / (theta * dtim) - d1x0 * (1.0 - theta) / theta
        d2x1 = (d1x1 - d1x0) / (theta * dtim) - d2x0 * (1.0 - theta) / theta
        x0, d1x0, d2x0 = x1, d1x1, d2x1
        rows.append((real_time, math.cos(omega * real_time), res["iters"]))
        if j in keep:
            fields[j] = x1.copy()
    return dict(rows=rows, x=x0, d1x=d1x0, d2x=d2x0, fields=fields, dtim=dtim)


def p1210(g_coord_pp, g_g, neq, fext, e, v, sbary, rho, dtim, pload, nstep, npri, npes=1, form=0):
    """p1210.f90 on global arrays (orc_p1210_run): explicit elasto-plastic (von Mises) dynamics with the lumped mass of
    :93-104.  form 0: elements_2 as the reference writes it (the arithmetic of k_p1210_elements); form 1: the operator
    form of the tensor-core kernel k_p1210_mf (another rounding).  -> dict(mm, snaps = [(step, x1, d1x1, d2x1)] every
    npri steps)."""
    lib().orc_set_p1210_form(int(form))
    g, gg, fe = _f64(g_coord_pp), _i32(g_g), _f64(fext)
    nels, nod = g.shape[0], g.shape[2]
    nout = nstep // npri
    snap = np.zeros((max(nout, 1), 3, neq))
    mm = np.zeros(neq)
    L = lib()
    rc = L.orc_p1210_run(C.c_int64(nels), nod, 8, _p(g), _p(gg), C.c_int64(neq), _p(fe), C.c_double(e), C.c_double(v),
                         C.c_double(sbary), C.c_double(rho), C.c_double(dtim), C.c_double(pload), int(nstep), int(npri),
                         int(npes), _p(mm), _p(snap))
    assert rc == 0
    return dict(mm=mm, snaps=[((k + 1) * npri, snap[k, 0], snap[k, 1], snap[k, 2]) for k in range(nout)])


def cube_p129(nxe, nye, nze, aa, bb, cc, rho=2000.0, e=1.0e5, v=0.3, alpha1=0.0008, beta1=0.5, nstep=40, npri=1, theta=1.0,
              omega=0.01, tol=1e-4, limit=3000, nip=27, deck_rounding=False):
    """p12meshgen's p129 cantilever (p12meshgen.f90, CASE('p129')): geometry_20bxz bricks, the first nr nodes (the plane
    y = 0) fully fixed, 2*nxe+1 loaded nodes on the last line of the mesh, nres the monitored equation."""
    L = lib()
    nels = nxe * nye * nze
    nr = 3 * nxe * nze + 2 * nxe + 2 * nze + 1
    nn = ((2 * nxe + 1) * (nze + 1) + (nxe + 1) * nze) * (nye + 1) + (nxe + 1) * (nze + 1) * nye
    m = Mesh()
    m.program, m.nod, m.nodof, m.nip, m.nels, m.nn, m.nr = 129, 20, 3, nip, nels, nn, nr
    m.rho, m.e, m.v, m.alpha1, m.beta1, m.nstep, m.npri, m.theta, m.omega, m.tol, m.limit = rho, e, v, alpha1, beta1, nstep, npri, theta, omega, tol, limit
    m.nres = 3 * (nye * (nxe + 1) * (nze + 1) + nr * (nye - 1) + (nxe + 1))
    m.g_num_pp = np.empty((nels, 20), np.int32)
    m.g_coord_pp = np.empty((nels, 3, 20))
    assert L.orc_cube_elements(20, nxe, nze, aa, bb, cc, 0, nels, _p(m.g_num_pp), _p(m.g_coord_pp)) == 0
    if deck_rounding:                             # the deck holds the coordinates as F12.4
        m.g_coord_pp = np.round(m.g_coord_pp, 4)
    rest = np.zeros((4, nr), np.int32)
    rest[0] = np.arange(1, nr + 1)
    m.rest = rest.copy()
    L.orc_rearrange(nr, 3, _p(rest))
    m.g_g_pp = np.zeros((nels, 60), np.int32)
    L.orc_find_g_all(20, 3, nels, _p(m.g_num_pp), _p(m.g_g_pp), nr, _p(rest))
    m.neq = int(m.g_g_pp.max())
    loaded = 2 * nxe + 1
    node = nn - loaded + np.arange(1, loaded + 1)
    val = np.where((np.arange(1, loaded + 1) == 1) | (np.arange(1, loaded + 1) == loaded), 25.0 / 12.0,
                   np.where(np.arange(1, loaded + 1) % 2 == 0, 25.0 / 3.0, 25.0 / 6.0))
    if deck_rounding:
        val = _round_sig(val, 8)
    nf_z = np.zeros(nn + 1, np.int32)
    nf_z[m.g_num_pp.ravel()] = m.g_g_pp.reshape(nels, 20, 3)[:, :, 2].ravel()
    m.r_pp = np.zeros(m.neq)
    eq = nf_z[node]
    m.r_pp[eq[eq > 0] - 1] = val[eq > 0]
    m.loaded_nodes, m.load_node, m.load_val, m.total_load = loaded, node.astype(np.int32), val, float(val.sum())
    return m


def form_k_explicit(g_coord_pp, nip, kx, ky, kz, dtim):
    """p125.f90:66-79: (store_pm (nels,8,8), mass (nels,8)); mass(i) = SUM(pm(i,:)), j ascending."""
    _, _, kc, pm = form_k_transient(g_coord_pp, nip, kx, ky, kz, 1.0, 1.0, 0.5, dtim, raw=True)
    mass = 0.0 + pm[:, 0, :]
    for j in range(1, 8):
        mass = mass + pm[:, j, :]
    store = 0.0 - kc * dtim
    idx = np.arange(8)
    store[:, idx, idx] = mass - kc[:, idx, idx] * dtim
    return store, mass


def p125(store_pm, mass, g_g, neq, val0, nstep, npes=1, keep=()):
    """The recursion of p125.f90:82-99: globma = 1/scatter(mass); loads = scatter(store_pm*gather(loads))*globma."""
    globma = 1.0 / scatter(g_g, mass, neq, npes)
    x = np.full(neq, float(val0))
    fields = {}
    for j in range(1, nstep + 1):
        x = apply(store_pm, g_g, neq, x, npes) * globma
        if j in keep:
            fields[j] = x.copy()
    return dict(x=x, fields=fields, globma=globma)
