"""Python mirror of the device API (section A of include/parafem_b200.h).

``Solver`` holds one pf_handle = one rank = one B200.  Every method is a thin
ctypes call into libparafem_b200.so; there is no PyTorch or CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PfError, c_i64, check, f64, i32, lib, ptr


def nccl_unique_id():
    """128-byte NCCL id created by rank 0; broadcast it and pass it to every Solver."""
    buf = np.zeros(128, np.uint8)
    check(lib().pf_nccl_unique_id(ptr(buf)), what="pf_nccl_unique_id")
    return buf


class Solver:
    def __init__(self, rank=0, nranks=1, device=0, nccl_id=None):
        self._h = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        idp = ptr(np.ascontiguousarray(nccl_id, np.uint8)) if nccl_id is not None else None
        check(lib().pf_init(rank, nranks, device, idp, C.byref(self._h)), what="pf_init")
        self.prob = None

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if self._h:
            lib().pf_finalize(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        check(rc, self._h, what)

    # -- setup ------------------------------------------------------------------
    def setup_mesh(self, prob):
        """pf_setup_mesh from a host.Problem (this rank's g_coord_pp / g_g_pp)."""
        self.prob = prob
        self._ck(lib().pf_setup_mesh(self._h, prob.nod, prob.nodof, prob.nip, prob.nels_pp,
                                     ptr(f64(prob.g_coord_pp)), ptr(i32(prob.g_g_pp)), prob.neq,
                                     prob.ieq_start, prob.neq_pp), "pf_setup_mesh")

    def form_km_elastic(self, e, v):
        self._ck(lib().pf_form_km_elastic(self._h, e, v), "pf_form_km_elastic")

    def form_km_elastic_mat(self, prop, etype_pp):
        """xx2.f90:169-193: prop (np_types, 2) = (e, v) per material, etype_pp 1-based material of each element."""
        pr, et = f64(prop), i32(etype_pp)
        assert pr.ndim == 2 and pr.shape[1] == 2 and et.size == self.prob.nels_pp
        self._ck(lib().pf_form_km_elastic_mat(self._h, pr.shape[0], ptr(pr), ptr(et)), "pf_form_km_elastic_mat")

    def form_kc_laplace(self, kx, ky, kz):
        self._ck(lib().pf_form_kc_laplace(self._h, kx, ky, kz), "pf_form_kc_laplace")

    def set_storkm(self, storkm):
        k = f64(storkm)
        assert k.size == self.prob.nels_pp * self.prob.ntot ** 2
        self._ck(lib().pf_set_storkm(self._h, ptr(k)), "pf_set_storkm")

    def set_matrix_free(self, on=True):
        """BASELINE config E: call before form_km_elastic."""
        self._ck(lib().pf_set_matrix_free(self._h, int(on)), "pf_set_matrix_free")

    def set_storkm_layout(self, layout):
        """0: the reference's storkm_pp; 1: packed lower triangles (half the stream). Before forming."""
        self._ck(lib().pf_set_storkm_layout(self._h, int(layout)), "pf_set_storkm_layout")

    def measure_fp64(self):
        t = C.c_double()
        self._ck(lib().pf_measure_fp64(self._h, C.byref(t)), "pf_measure_fp64")
        return t.value

    def measure_fp64_tensor(self):
        """TFLOP/s of the FP64 tensor pipe (DMMA.8x8x4), the denominator of the tensor-core matrix-free kernel."""
        t = C.c_double()
        self._ck(lib().pf_measure_fp64_tensor(self._h, C.byref(t)), "pf_measure_fp64_tensor")
        return t.value

    def measure_matvec(self, reps=50):
        """ms per launch of the current problem's mat-vec kernel, ``reps`` launches back to back between one event pair."""
        t = C.c_double()
        self._ck(lib().pf_measure_matvec(self._h, int(reps), C.byref(t)), "pf_measure_matvec")
        return t.value

    def measure_hbm_read(self):
        """GB/s of a read-only stream through the mat-vec's bulk-copy ring (no arithmetic)."""
        t = C.c_double()
        self._ck(lib().pf_measure_hbm_read(self._h, C.byref(t)), "pf_measure_hbm_read")
        return t.value

    def get_storkm(self, iel0=0, n=None):
        n = self.prob.nels_pp - iel0 if n is None else n
        nt = self.prob.ntot
        out = np.empty((n, nt, nt))
        self._ck(lib().pf_get_storkm(self._h, iel0, n, ptr(out)), "pf_get_storkm")
        return out

    def build_precon(self, no_f=None, penalty=1e20):
        no_f = np.zeros(0, np.int32) if no_f is None else i32(no_f)
        self._nfixed = no_f.size
        self._ck(lib().pf_build_precon(self._h, no_f.size, ptr(no_f) if no_f.size else None, penalty),
                 "pf_build_precon")

    def diag_precon(self):
        out = np.empty(self.prob.neq_pp)
        self._ck(lib().pf_get_diag_precon(self._h, ptr(out)), "pf_get_diag_precon")
        return out

    def store(self):
        out = np.empty(self._nfixed)
        self._ck(lib().pf_get_store(self._h, ptr(out)), "pf_get_store")
        return out

    # -- solve --------------------------------------------------------------------
    def pcg_solve(self, r_pp, tol, limit):
        """p121.f90:87-104 through one C-ABI call with host buffers. -> (xnew_pp, iters, converged)"""
        r = f64(r_pp)
        x = np.empty(self.prob.neq_pp)
        it, cv = C.c_int(), C.c_int()
        self._ck(lib().pf_pcg_solve(self._h, ptr(r), tol, limit, ptr(x), C.byref(it), C.byref(cv)), "pf_pcg_solve")
        return x, it.value, bool(cv.value)

    def pcg_km(self, km, diag_precon_pp, r_pp, tol, limit):
        """PCG_KM (maths.f90:1152-1323): one km(ntot,ntot) for every element. -> (xnew_pp, iters, converged)"""
        k, dg, r = f64(km), f64(diag_precon_pp), f64(r_pp)
        assert k.shape == (self.prob.ntot, self.prob.ntot) and dg.size == r.size == self.prob.neq_pp
        x = np.empty(self.prob.neq_pp)
        it, cv = C.c_int(), C.c_int()
        self._ck(lib().pf_pcg_km(self._h, ptr(k), ptr(dg), ptr(r), tol, limit, ptr(x), C.byref(it), C.byref(cv)), "pf_pcg_km")
        return x, it.value, bool(cv.value)

    def pcg_load_rhs(self, r_pp):
        self._ck(lib().pf_pcg_load_rhs(self._h, ptr(f64(r_pp))), "pf_pcg_load_rhs")

    def pcg_run(self, tol, limit):
        """Device-resident solve. -> (iters, converged, elapsed_ms on the solver stream)"""
        it, cv, ms = C.c_int(), C.c_int(), C.c_double()
        self._ck(lib().pf_pcg_run(self._h, tol, limit, C.byref(it), C.byref(cv), C.byref(ms)), "pf_pcg_run")
        return it.value, bool(cv.value), ms.value

    def pcg_get_x(self):
        x = np.empty(self.prob.neq_pp)
        self._ck(lib().pf_pcg_get_x(self._h, ptr(x)), "pf_pcg_get_x")
        return x

    def ratio_history(self, maxn=100000):
        out = np.zeros(maxn)
        n = C.c_int()
        self._ck(lib().pf_get_ratio_history(self._h, ptr(out), maxn, C.byref(n)), "pf_get_ratio_history")
        return out[:n.value].copy()

    # -- p124: transient conduction ---------------------------------------------------
    def form_k_transient(self, kx, ky, kz, rho, cp, theta, dtim):
        """p124.f90:81-95: storka_pp (the PCG matrix) and storkb_pp on the device."""
        self._ck(lib().pf_form_k_transient(self._h, kx, ky, kz, rho, cp, theta, dtim), "pf_form_k_transient")

    def get_storkb(self, iel0=0, n=None):
        n = self.prob.nels_pp - iel0 if n is None else n
        nt = self.prob.ntot
        out = np.empty((n, nt, nt))
        self._ck(lib().pf_get_storkb(self._h, iel0, n, ptr(out)), "pf_get_storkb")
        return out

    def transient_start(self, val0, val_f=None):
        v = f64(val_f) if val_f is not None and len(val_f) else None
        self._ck(lib().pf_transient_start(self._h, val0, ptr(v)), "pf_transient_start")

    def transient_step(self, tol, limit, loads_pp=None):
        """One pass of p124's timesteps loop. -> (iters, converged, elapsed_ms)"""
        it, cv, ms = C.c_int(), C.c_int(), C.c_double()
        l = f64(loads_pp) if loads_pp is not None else None
        self._ck(lib().pf_transient_step(self._h, ptr(l), tol, limit, C.byref(it), C.byref(cv), C.byref(ms)),
                 "pf_transient_step")
        return it.value, bool(cv.value), ms.value

    # -- p125: explicit transient conduction -----------------------------------------------
    def form_k_explicit(self, kx, ky, kz, dtim):
        """p125.f90:66-82: store_pm_pp and the inverted lumped mass globma_pp on the device."""
        self._ck(lib().pf_form_k_explicit(self._h, kx, ky, kz, dtim), "pf_form_k_explicit")

    def explicit_start(self, val0):
        self._ck(lib().pf_explicit_start(self._h, val0), "pf_explicit_start")

    def explicit_steps(self, nsteps):
        """nsteps passes of p125's recursion on the device. -> elapsed_ms"""
        ms = C.c_double()
        self._ck(lib().pf_explicit_steps(self._h, nsteps, C.byref(ms)), "pf_explicit_steps")
        return ms.value

    # -- p129: forced vibration, theta method ---------------------------------------------------
    def form_dynamic(self, e, v, rho, alpha1, beta1, theta, dtim):
        """p129.f90:80-98: store_km_pp, consistent store_mm_pp and the three matrix sets of the time loop."""
        self._ck(lib().pf_form_dynamic(self._h, e, v, rho, alpha1, beta1, theta, dtim), "pf_form_dynamic")

    def dynamic_start(self, fext_pp):
        self._ck(lib().pf_dynamic_start(self._h, ptr(f64(fext_pp))), "pf_dynamic_start")

    def dynamic_step(self, load_factor, tol, limit):
        """One time step (p129.f90:113-149). -> (iters, converged, elapsed_ms)"""
        it, cv, ms = C.c_int(), C.c_int(), C.c_double()
        self._ck(lib().pf_dynamic_step(self._h, load_factor, tol, limit, C.byref(it), C.byref(cv), C.byref(ms)), "pf_dynamic_step")
        return it.value, bool(cv.value), ms.value

    def dynamic_get(self):
        """-> (x1_pp, d1x1_pp, d2x1_pp) of the last step"""
        out = [np.empty(self.prob.neq_pp) for _ in range(3)]
        self._ck(lib().pf_dynamic_get(self._h, ptr(out[0]), ptr(out[1]), ptr(out[2])), "pf_dynamic_get")
        return tuple(out)

    # -- p1210: explicit elasto-plastic (von Mises) dynamics -------------------------------------
    def vm_explicit_begin(self, e, v, sbary, rho, dtim, pload, fext_pp):
        """p1210.f90:93-112 after setup_mesh: element tables, lumped mass, external loads, zero state."""
        self._ck(lib().pf_vm_explicit_begin(self._h, e, v, sbary, rho, dtim, pload, ptr(f64(fext_pp))), "pf_vm_explicit_begin")

    def vm_explicit_set_form(self, form):
        """0: elements_2 as the reference writes it; 1: operator form on the FP64 tensor cores."""
        self._ck(lib().pf_vm_explicit_set_form(self._h, int(form)), "pf_vm_explicit_set_form")

    def vm_explicit_steps(self, nsteps):
        """nsteps passes of time_steps (p1210.f90:114-150) on the device. -> elapsed_ms"""
        ms = C.c_double()
        self._ck(lib().pf_vm_explicit_steps(self._h, int(nsteps), C.byref(ms)), "pf_vm_explicit_steps")
        return ms.value

    def vm_explicit_get(self, mass=False):
        """-> (x1_pp, d1x1_pp, d2x1_pp[, mm_pp])"""
        out = [np.empty(self.prob.neq_pp) for _ in range(4 if mass else 3)]
        self._ck(lib().pf_vm_explicit_get(self._h, ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]) if mass else None),
                 "pf_vm_explicit_get")
        return tuple(out)

    # -- p122: elasto-plasticity ---------------------------------------------------------------
    def plastic_begin(self, phi, c, psi, e, v):
        """p122.f90:88-93 after form_km_elastic + build_precon. -> the critical time step dt"""
        dt = C.c_double()
        self._ck(lib().pf_plastic_begin(self._h, phi, c, psi, e, v, C.byref(dt)), "pf_plastic_begin")
        return dt.value

    def plastic_increment(self, qinc, plasits, plastol, cjits, cjtol, ld0_pp=None, valf_pp=None):
        """One load increment (p122.f90:115-205). -> (plasiters, cjtot, elapsed_ms)"""
        pl, cj, ms = C.c_int(), C.c_int(), C.c_double()
        l0 = f64(ld0_pp) if ld0_pp is not None else None
        vf = f64(valf_pp) if valf_pp is not None and len(valf_pp) else None
        self._ck(lib().pf_plastic_increment(self._h, qinc, ptr(l0), ptr(vf), plasits, plastol, cjits, cjtol, C.byref(pl),
                                            C.byref(cj), C.byref(ms)), "pf_plastic_increment")
        return pl.value, cj.value, ms.value

    def plastic_totd(self):
        out = np.empty(self.prob.neq_pp)
        self._ck(lib().pf_plastic_get(self._h, ptr(out), 0, 0, None), "pf_plastic_get")
        return out

    def plastic_tensor(self, iel=0, ig=0):
        out = np.empty(6)
        self._ck(lib().pf_plastic_get(self._h, None, iel, ig, ptr(out)), "pf_plastic_get")
        return out

    # -- fine-grained -----------------------------------------------------------------
    def gather(self, p_pp):
        out = np.empty((self.prob.nels_pp, self.prob.ntot))
        self._ck(lib().pf_gather(self._h, ptr(f64(p_pp)), ptr(out)), "pf_gather")
        return out

    def matvec(self, pmul_pp):
        out = np.empty((self.prob.nels_pp, self.prob.ntot))
        self._ck(lib().pf_matvec(self._h, ptr(f64(pmul_pp)), ptr(out)), "pf_matvec")
        return out

    def scatter(self, utemp_pp):
        out = np.empty(self.prob.neq_pp)
        self._ck(lib().pf_scatter(self._h, ptr(f64(utemp_pp)), ptr(out)), "pf_scatter")
        return out

    def apply(self, p_pp):
        out = np.empty(self.prob.neq_pp)
        self._ck(lib().pf_apply(self._h, ptr(f64(p_pp)), ptr(out)), "pf_apply")
        return out

    def dot(self, a_pp, b_pp):
        res = C.c_double()
        self._ck(lib().pf_dot(self._h, ptr(f64(a_pp)), ptr(f64(b_pp)), C.byref(res)), "pf_dot")
        return res.value

    def norm(self, a_pp):
        res = C.c_double()
        self._ck(lib().pf_norm(self._h, ptr(f64(a_pp)), C.byref(res)), "pf_norm")
        return res.value

    def sum(self, a_pp):
        res = C.c_double()
        self._ck(lib().pf_sum(self._h, ptr(f64(a_pp)), C.byref(res)), "pf_sum")
        return res.value

    def centroid_stress(self, iel, e, v):
        out = np.empty(6)
        self._ck(lib().pf_centroid_stress(self._h, iel, e, v, ptr(out)), "pf_centroid_stress")
        return out

    def point_stress(self, iel, xi, eta, zeta, e, v):
        """sigma at a local point of element iel (0-based, local); see pf_point_stress."""
        out = np.empty(6)
        self._ck(lib().pf_point_stress(self._h, iel, xi, eta, zeta, e, v, ptr(out)), "pf_point_stress")
        return out

    def halo_transport(self):
        return {0: "none", 1: "nccl", 2: "peer"}[lib().pf_halo_transport(self._h)]

    def last_solve_ms(self):
        ms = C.c_double()
        self._ck(lib().pf_get_last_solve_ms(self._h, C.byref(ms)), "pf_get_last_solve_ms")
        return ms.value

    # -- measurement -----------------------------------------------------------------
    def set_profile(self, on):
        self._ck(lib().pf_set_profile(self._h, int(on)), "pf_set_profile")

    def reset_profile(self):
        self._ck(lib().pf_reset_profile(self._h), "pf_reset_profile")

    def kernel_ms(self, which):
        ms, n = C.c_double(), c_i64()
        self._ck(lib().pf_get_kernel_ms(self._h, which, C.byref(ms), C.byref(n)), "pf_get_kernel_ms")
        return ms.value, n.value

    def kernel_launches(self):
        return lib().pf_kernel_launches(self._h)

    def device_info(self):
        sm, fr, tot = C.c_int(), c_i64(), c_i64()
        self._ck(lib().pf_device_info(self._h, C.byref(sm), C.byref(fr), C.byref(tot)), "pf_device_info")
        return sm.value, fr.value, tot.value


def setup_problem(solver, prob, matrix_free=False, layout=0):
    """The device part of p121.f90:49-69,86 / p123.f90:57-92,120-125 for one rank."""
    solver.setup_mesh(prob)
    solver.set_matrix_free(matrix_free)
    solver.set_storkm_layout(layout)
    if prob.program == 121:
        if prob.prop is not None:           # per-element materials (xx2)
            solver.form_km_elastic_mat(prob.prop, prob.etype_pp)
        else:
            solver.form_km_elastic(prob.e, prob.v)
        solver.build_precon(prob.no_f if prob.no_f.size else None, 1e20)
        if prob.no_f.size:
            # r_pp(j) = store_pp(i)*valf(k)   (xx2.f90:294-300)
            prob.r_pp[prob.no_f - prob.ieq_start] = solver.store() * prob.val_f
    elif prob.program == 129:
        # p129.f90:78-105: period / 20 time step, the matrix sets, the preconditioner of store_mm*c3 + store_km*c4
        import math
        prob.dtim = 2.0 * math.acos(-1.0) / prob.omega / 20.0
        solver.form_dynamic(prob.e, prob.v, prob.rho, prob.alpha1, prob.beta1, prob.theta, prob.dtim)
        solver.build_precon(None, 1e20)
        solver.dynamic_start(prob.r_pp)
    elif prob.program == 1210:
        # p1210.f90:93-112: no element matrices, no preconditioner
        solver.vm_explicit_begin(prob.e, prob.v, prob.sbary, prob.rho, prob.dtim, prob.pload, prob.r_pp)
        solver.vm_explicit_set_form(getattr(prob, "form", 0))
    elif prob.program == 122:
        # p122.f90:94-114: storkm_pp, the preconditioner with the penalty on this rank's fixed freedoms, zero stresses
        solver.form_km_elastic(prob.e, prob.v)
        solver.build_precon(prob.no_f if prob.no_f.size else None, 1e20)
        prob.dt = solver.plastic_begin(prob.phi, prob.c, prob.psi, prob.e, prob.v)
    elif prob.program == 125:
        solver.form_k_explicit(prob.kx, prob.ky, prob.kz, prob.dtim)
    elif prob.program == 124:
        solver.form_k_transient(prob.kx, prob.ky, prob.kz, prob.rho, prob.cp, prob.theta, prob.dtim)
        solver.build_precon(prob.no_f, 1e20)
    else:
        solver.form_kc_laplace(prob.kx, prob.ky, prob.kz)
        solver.build_precon(prob.no_f, 1e20)
        if prob.no_f.size:
            # r_pp(j) = store_pp(i)*val_f(k)   (p123.f90:127-131)
            st = solver.store()
            prob.r_pp[prob.no_f - prob.ieq_start] = st * prob.val_f
    return solver
