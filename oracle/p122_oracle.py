"""CPU oracle for program p122 (3-D elasto-plasticity, Mohr-Coulomb, viscoplastic strain method, PCG); the device
side is pf_plastic_begin / pf_plastic_increment (tests/test_gpu_plastic.py).

TEST INFRASTRUCTURE ONLY.  Restates programs/5th_ed/p122/p122.f90:66-233 with numpy (the element integrals
vectorised over elements and Gauss points) on top of the C oracle's storkm / gather-matvec-scatter, and the
library routines it calls: invar (new_library.f90:1813-1916, nst = 6), mocouf (:2364-2417), mocouq (:2423-2490),
formm (:85-198, nst = 6), checon_par (maths.f90:999-1069).  Pinned against examples/5th_ed/p122/demo/p122_demo.res
(tests/test_oracle_golden.py): displacement, stresses and iteration counts of the ten load increments.
"""
import ctypes as C

import numpy as np

from . import _f64, _i32, _p, apply, dot_blocked, form_km_elastic, lib, scatter


def _bee_detw(g_coord_pp, nip=8):
    """bee(6,24) and det*w of every (element, Gauss point): p122.f90:203-207."""
    L = lib()
    pts, wts = np.zeros((3, nip)), np.zeros(nip)
    L.orc_sample_hex(nip, _p(pts), _p(wts))
    co = np.transpose(_f64(g_coord_pp), (0, 2, 1))                    # (nels, nod, 3)
    nels, nod = co.shape[:2]
    bee = np.zeros((nels, nip, 6, 3 * nod))
    detw = np.zeros((nels, nip))
    for ig in range(nip):
        der = np.zeros((nod, 3))
        L.orc_shape_der(nod, _p(pts), nip, ig, _p(der))               # der(a,m) at [m,a]
        jac = np.einsum("ma,emb->eab", der, co)                        # jac(a,b) = sum_m der(a,m) coord(m,b)
        det = np.linalg.det(jac)
        deriv = np.einsum("eab,mb->eam", np.linalg.inv(jac), der)      # deriv(a,m) = sum_b inv(a,b) der(b,m)
        x, y, z = deriv[:, 0], deriv[:, 1], deriv[:, 2]
        b = bee[:, ig]
        b[:, 0, 0::3] = x; b[:, 1, 1::3] = y; b[:, 2, 2::3] = z       # beemat, new_library.f90:976-993
        b[:, 3, 0::3] = y; b[:, 3, 1::3] = x
        b[:, 4, 1::3] = z; b[:, 4, 2::3] = y
        b[:, 5, 0::3] = z; b[:, 5, 2::3] = x
        detw[:, ig] = det * wts[ig]
    return bee, detw


def _deemat(e, v):
    v1, c = v / (1 - v), e * (1 - v) / ((1 + v) * (1 - 2 * v))
    vv = (1 - 2 * v) * .5 / (1 - v)
    d = np.zeros((6, 6))
    d[:3, :3] = v1 * c
    d[np.arange(3), np.arange(3)] = c
    d[np.arange(3, 6), np.arange(3, 6)] = vv * c
    return d


def _invar(s):
    sigm = (s[..., 0] + s[..., 1] + s[..., 2]) / 3.0
    d2 = ((s[..., 0] - s[..., 1]) ** 2 + (s[..., 1] - s[..., 2]) ** 2 + (s[..., 2] - s[..., 0]) ** 2) / 6.0 \
        + s[..., 3] ** 2 + s[..., 4] ** 2 + s[..., 5] ** 2
    ds = s[..., :3] - sigm[..., None]
    d3 = ds[..., 0] * ds[..., 1] * ds[..., 2] - ds[..., 0] * s[..., 4] ** 2 - ds[..., 1] * s[..., 5] ** 2 \
        - ds[..., 2] * s[..., 3] ** 2 + 2.0 * s[..., 3] * s[..., 4] * s[..., 5]
    dsbar = np.sqrt(3.0) * np.sqrt(d2)
    with np.errstate(divide="ignore", invalid="ignore"):
        sine = np.clip(-3.0 * np.sqrt(3.0) * d3 / (2.0 * np.sqrt(d2) ** 3), -1.0, 1.0)
    theta = np.where(dsbar < 1e-10, 0.0, np.arcsin(np.where(dsbar < 1e-10, 0.0, sine)) / 3.0)
    return sigm, dsbar, theta


def _mocouf(phi, c, sigm, dsbar, theta):
    phir = phi * 4.0 * np.arctan(1.0) / 180.0
    return np.sin(phir) * sigm + dsbar * (np.cos(theta) / np.sqrt(3.0) - np.sin(theta) * np.sin(phir) / 3.0) - c * np.cos(phir)


def _mocouq(psi, dsbar, theta):
    psir = psi * 4.0 * np.arctan(1.0) / 180.0
    snth, snps, sq3 = np.sin(theta), np.sin(psir), np.sqrt(3.0)
    c1 = np.where(snth < 0, -1.0, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        csth, cs3, tn3 = np.cos(theta), np.cos(3 * theta), np.tan(3 * theta)
        tnth = snth / csth
        dq2b = sq3 * csth / dsbar * ((1 + tnth * tn3) + snps * (tn3 - tnth) / sq3) * .5
        dq3b = .5 * 3 * (sq3 * snth + snps * csth) / (cs3 * dsbar * dsbar)
        dq2a = (sq3 * .5 - c1 * snps * .5 / sq3) * sq3 * .5 / dsbar
    corner = np.abs(snth) > .49
    return np.full_like(dsbar, snps), np.where(corner, dq2a, dq2b), np.where(corner, 0.0, dq3b)


def _formm(s):
    """m1, m2, m3 (..., 6, 6) of formm, nst = 6."""
    sx, sy, sz, txy, tyz, tzx = (s[..., k] for k in range(6))
    sigm = (sx + sy + sz) / 3.0
    dx, dy, dz = sx - sigm, sy - sigm, sz - sigm
    m1 = np.zeros(s.shape[:-1] + (6, 6)); m2 = np.zeros_like(m1); m3 = np.zeros_like(m1)
    with np.errstate(divide="ignore", invalid="ignore"):
        m1[..., :3, :3] = (1.0 / (3.0 * sigm))[..., None, None]
    for i in range(3):
        m2[..., i, i] = 2.0; m2[..., i + 3, i + 3] = 6.0
    m2[..., 0, 1] = m2[..., 0, 2] = m2[..., 1, 2] = -1.0
    up = {(0, 0): dx, (0, 1): dz, (0, 2): dy, (0, 3): txy, (0, 4): -2 * tyz, (0, 5): tzx, (1, 1): dy, (1, 2): dx, (1, 3): txy,
          (1, 4): tyz, (1, 5): -2 * tzx, (2, 2): dz, (2, 3): -2 * txy, (2, 4): tyz, (2, 5): tzx, (3, 3): -3 * dz, (3, 4): 3 * tzx,
          (3, 5): 3 * tyz, (4, 4): -3 * dx, (4, 5): 3 * txy, (5, 5): -3 * dy}
    for (i, j), v in up.items():
        m3[..., i, j] = v
    for m in (m1, m2, m3):
        for i in range(6):
            for j in range(i + 1, 6):
                m[..., j, i] = m[..., i, j]
    return m1 / 3.0, m2 / 3.0, m3 / 3.0


def _checon(new, old, tol):
    """checon_par: converged = max|new-old| / max|new| <= tol; old <- new."""
    big = np.abs(new).max()
    conv = bool(np.abs(new - old).max() / big <= tol) if big > 0 else False
    old[:] = new
    return conv


def p122(g_coord_pp, g_g_pp, neq, phi, c, psi, e, v, qinc, plasits, cjits, plastol, cjtol, no_f=None, valf=None, ld0=None,
         npes=1, penalty=1e20, c_elements=True, red_mode=0):
    """-> list of dict(disp1 = totd(1), sigma = tensor(:,1,1), cjtot, plasiters) per load increment, and totd.
    c_elements: the Gauss-point update through orc_p122_elements (pf_oracle.c: the reference's loop with defined
    summation orders -- what the device kernel is held to); False: the vectorised numpy form of the same.
    red_mode 1: DOT_PRODUCT_P through the fixed blocked tree of the CUDA kernels (orc_dot_blocked) instead of numpy's
    dot -- another legal summation order, and the one a GPU-vs-oracle comparison needs."""
    dot = (lambda a, b: float(dot_blocked(a, b))) if red_mode else (lambda a, b: float(np.dot(a, b)))
    g_g = _i32(g_g_pp)
    nels, ntot = g_g.shape
    km = form_km_elastic(g_coord_pp, ntot // 3, 8, e, v)
    dee = _deemat(e, v)
    bee, detw = _bee_detw(g_coord_pp)
    nip = bee.shape[1]
    diag = scatter(g_g, np.ascontiguousarray(km[:, np.arange(ntot), np.arange(ntot)]), neq, npes)
    nfix = 0 if no_f is None else len(no_f)
    if nfix:
        no_f = np.asarray(no_f) - 1
        diag[no_f] += penalty
        store = diag[no_f].copy()
    diag = 1.0 / diag
    snph = np.sin(phi * np.pi / 180.0)
    dt = 4.0 * (1 + v) * (1 - 2 * v) / (e * (1 - 2 * v + snph * snph))
    tensor = np.zeros((nels, nip, 6)); totd = np.zeros(neq); x = np.zeros(neq); xnew = np.zeros(neq); oldis = np.zeros(neq)
    safe = np.maximum(g_g, 1) - 1
    out = []
    for q in qinc:
        plasiters, cjtot = 0, 0
        bdylds = np.zeros(neq); evpt = np.zeros((nels, nip, 6))
        while True:
            plasiters += 1
            loads = np.zeros(neq)
            if plasiters == 1:
                if nfix:
                    loads[no_f] = store * valf * q
                if ld0 is not None:
                    loads = ld0 * q + bdylds
            else:
                if ld0 is not None:
                    loads = ld0 * q
                loads = loads + bdylds
                if nfix:
                    loads[no_f] = 0.0
            r = loads - apply(km, g_g, neq, x, npes)
            d = diag * r
            p = d.copy()
            cjiters = 0
            while True:
                cjiters += 1
                u = apply(km, g_g, neq, p, npes)
                if nfix:
                    u[no_f] = p[no_f] * store if plasiters == 1 else 0.0
                up = dot(r, d)
                alpha = up / dot(p, u)
                xnew = x + p * alpha
                r = r - u * alpha
                d = diag * r
                beta = dot(r, d) / up
                p = d + p * beta
                if _checon(xnew, x, cjtol) or cjiters == cjits:
                    break
            cjtot += cjiters
            loads = xnew.copy()
            conv = _checon(loads, oldis, plastol)
            if plasiters == 1:
                conv = False
            last = conv or plasiters == plasits
            if last:
                bdylds = np.zeros(neq)
            eld = np.where(g_g > 0, loads[safe], 0.0)                                 # gather(loads_pp,pmul_pp)
            if c_elements:
                bload = np.empty((nels, ntot))
                coord = _f64(g_coord_pp)
                eldc = _f64(eld)
                rc = lib().orc_p122_elements(nels, ntot // 3, nip, _p(coord), e, v, phi, c, psi, dt, int(last), _p(eldc),
                                             _p(evpt), _p(tensor), _p(bload))
                assert rc == 0
                bdylds = bdylds + scatter(g_g, bload, neq, npes)
                if last:
                    break
                continue
            eps = np.einsum("egsc,ec->egs", bee, eld) - evpt
            sigma = eps @ dee.T
            stress = sigma + tensor
            sigm, dsbar, theta = _invar(stress)
            f = _mocouf(phi, c, sigm, dsbar, theta)
            yielding = f >= 0.0
            if last:
                devp = stress
            else:
                dq1, dq2, dq3 = _mocouq(psi, dsbar, theta)
                m1, m2, m3 = _formm(stress)
                flow = f[..., None, None] * (m1 * dq1[..., None, None] + m2 * dq2[..., None, None] + m3 * dq3[..., None, None])
                evp = np.einsum("egij,egj->egi", flow, stress) * dt
                evp = np.where(yielding[..., None], np.nan_to_num(evp), 0.0)
                evpt = evpt + evp
                devp = evp @ dee.T
            eload = np.einsum("egsc,egs->egc", bee, np.where(yielding[..., None], devp, 0.0))
            bload = np.einsum("egc,eg->ec", eload, detw)
            if last:
                tensor = stress.copy()
            bdylds = bdylds + scatter(g_g, bload, neq, npes)          # scatter ADDS into bdylds_pp (gather_scatter.f90:759-773)
            if last:
                break
        totd = totd + loads
        out.append(dict(disp1=float(totd[0]), sigma=tensor[0, 0].copy(), cjtot=cjtot, plasiters=plasiters))
        if plasiters == plasits:
            break
    return out, totd
