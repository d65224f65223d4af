"""Test helper: 4-node tetrahedra from a generated brick mesh (six tetrahedra per 8-node brick around the
(-,-,-) -> (+,+,+) diagonal, oriented so that det(jac) > 0 with the reference's shape_der for nod = 4)."""
import dataclasses
import itertools

import numpy as np

from parafem_b200 import host

# S&G local node order of the 8-node brick as (xi, eta, zeta) signs (shape_fun, new_library.f90:414-420)
_SIGNS = [(-1, -1, -1), (-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1)]
_LOCAL = {s: k for k, s in enumerate(_SIGNS)}


def _kuhn():
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [-1, -1, -1]
        path = [tuple(v)]
        for a in perm:
            v[a] = 1
            path.append(tuple(v))
        tets.append([_LOCAL[p] for p in path])
    return np.array(tets)            # (6, 4) local brick nodes


def tets_of(base):
    """(g_num (6*nels, 4) int32, g_coord (nn, 3)) from a hex8 host.Problem holding the whole mesh."""
    assert base.nod == 8 and base.npes == 1
    g_coord = np.zeros((base.nn, 3))
    g_coord[base.g_num_pp - 1] = np.transpose(base.g_coord_pp, (0, 2, 1))
    t = base.g_num_pp[:, _kuhn()].reshape(-1, 4).astype(np.int32)
    x = g_coord[t - 1]                                         # (ntet, 4, 3)
    det = np.linalg.det(x[:, :3, :] - x[:, 3:4, :])            # rows x1-x4, x2-x4, x3-x4 = jac
    flip = det < 0
    t[flip, 0], t[flip, 1] = t[flip, 1].copy(), t[flip, 0].copy()
    return np.ascontiguousarray(t), g_coord


def tet_problem(base, npes=1, numpe=1, free_all=False):
    """The brick problem `base` (whole mesh) re-meshed with tetrahedra; rank numpe's share.  free_all: no
    restrained nodes (nf = 1..nn, as p123.f90:54 does for nr = 0) -- the patch test fixes the boundary itself."""
    g_num, g_coord = tets_of(base)
    nels = g_num.shape[0]
    nf = np.arange(1, base.nn + 1, dtype=np.int32).reshape(-1, 1) if free_all else base.nf
    neq = int(nf.max())
    nels_pp, iel_start = host.calc_nels_pp(nels, npes, numpe)
    neq_pp, ieq_start = host.calc_neq_pp(neq, npes, numpe)
    gn = np.ascontiguousarray(g_num[iel_start - 1:iel_start - 1 + nels_pp])
    g_g = np.ascontiguousarray(nf[gn - 1].reshape(nels_pp, -1))
    coords = np.ascontiguousarray(np.transpose(g_coord[gn - 1], (0, 2, 1)))
    r = np.zeros(neq_pp) if free_all else np.ascontiguousarray(base.r_pp[ieq_start - 1:ieq_start - 1 + neq_pp])
    p = dataclasses.replace(base, nod=4, nip=1, nels=nels, npes=npes, numpe=numpe, nels_pp=nels_pp, iel_start=iel_start,
                            neq=neq, neq_pp=neq_pp, ieq_start=ieq_start, g_num_pp=gn, g_coord_pp=coords, g_g_pp=g_g,
                            nf=nf, r_pp=r, nr=0 if free_all else base.nr)
    p.g_coord = g_coord
    return p


def patch_problem(n=5, coef=(1.5, -2.0, 0.75, 3.0)):
    """Scalar patch test on tetrahedra: every boundary node held at the linear field T = a x + b y + c z + d
    (fixed freedoms, penalty rows of p123.f90:120-131); the interior must reproduce it."""
    base = host.cube_p123(n, n + 1, n - 1, aa=.3, bb=.2, cc=.25, kx=1., ky=1., kz=1., tol=1e-12, limit=2000)
    p = tet_problem(base, free_all=True)
    c = p.g_coord
    field = coef[0] * c[:, 0] + coef[1] * c[:, 1] + coef[2] * c[:, 2] + coef[3]
    on = np.zeros(p.nn, bool)
    for a in range(3):
        on |= np.isclose(c[:, a], c[:, a].min()) | np.isclose(c[:, a], c[:, a].max())
    p.no_f = (np.flatnonzero(on) + 1).astype(np.int32)
    p.val_f = np.ascontiguousarray(field[on])
    p.program = 123
    return p, field
