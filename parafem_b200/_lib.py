"""ctypes binding of libparafem_b200.so (include/parafem_b200.h).

The library is built in-tree by ``parafem_b200.build.build()`` (nvcc, sm_100a).
There is no fallback: if the shared object is missing, importing this module
raises, and every device entry point returns an error code (raised as
``PfError``) when no B200 is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libparafem_b200.so")


class PfError(RuntimeError):
    pass


c_i64 = C.c_int64
c_int = C.c_int
c_dbl = C.c_double
vp = C.c_void_p


class DeckInfo(C.Structure):
    _fields_ = [("program", c_int), ("meshgen", c_int), ("partitioner", c_int), ("nip", c_int),
                ("nod", c_int), ("limit", c_int), ("nels", c_i64), ("nn", c_i64), ("nr", c_i64),
                ("loaded", c_i64), ("fixed", c_i64), ("nres", c_i64), ("e", c_dbl), ("v", c_dbl),
                ("kx", c_dbl), ("ky", c_dbl), ("kz", c_dbl), ("tol", c_dbl),
                ("np_types", c_int), ("nstep", c_int), ("npri", c_int), ("pad_", c_int),
                ("val0", c_dbl), ("dtim", c_dbl), ("theta", c_dbl), ("rho", c_dbl), ("cp", c_dbl)]


P = C.POINTER
# name -> (restype, argtypes); every symbol include/parafem_b200.h declares
SIGNATURES = {
    # A. device API
    "pf_nccl_unique_id": (c_int, [vp]),
    "pf_init": (c_int, [c_int, c_int, c_int, vp, P(vp)]),
    "pf_finalize": (c_int, [vp]),
    "pf_last_error": (c_int, [vp, C.c_char_p, c_int]),
    "pf_version": (c_int, []),
    "pf_setup_mesh": (c_int, [vp, c_int, c_int, c_int, c_i64, vp, vp, c_i64, c_i64, c_i64]),
    "pf_form_km_elastic": (c_int, [vp, c_dbl, c_dbl]),
    "pf_form_kc_laplace": (c_int, [vp, c_dbl, c_dbl, c_dbl]),
    "pf_form_km_elastic_mat": (c_int, [vp, c_int, vp, vp]),
    "pf_set_storkm": (c_int, [vp, vp]),
    "pf_get_storkm": (c_int, [vp, c_i64, c_i64, vp]),
    "pf_set_matrix_free": (c_int, [vp, c_int]),
    "pf_set_storkm_layout": (c_int, [vp, c_int]),
    "pf_build_precon": (c_int, [vp, c_i64, vp, c_dbl]),
    "pf_get_diag_precon": (c_int, [vp, vp]),
    "pf_get_store": (c_int, [vp, vp]),
    "pf_pcg_solve": (c_int, [vp, vp, c_dbl, c_int, vp, P(c_int), P(c_int)]),
    "pf_pcg_load_rhs": (c_int, [vp, vp]),
    "pf_pcg_run": (c_int, [vp, c_dbl, c_int, P(c_int), P(c_int), P(c_dbl)]),
    "pf_pcg_get_x": (c_int, [vp, vp]),
    "pf_pcg_km": (c_int, [vp, vp, vp, vp, c_dbl, c_int, vp, P(c_int), P(c_int)]),
    "pf_get_ratio_history": (c_int, [vp, vp, c_int, P(c_int)]),
    "pf_form_k_transient": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl]),
    "pf_get_storkb": (c_int, [vp, c_i64, c_i64, vp]),
    "pf_transient_start": (c_int, [vp, c_dbl, vp]),
    "pf_transient_step": (c_int, [vp, vp, c_dbl, c_int, P(c_int), P(c_int), P(c_dbl)]),
    "pf_form_k_explicit": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl]),
    "pf_explicit_start": (c_int, [vp, c_dbl]),
    "pf_explicit_steps": (c_int, [vp, c_int, P(c_dbl)]),
    "pf_form_dynamic": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl]),
    "pf_dynamic_start": (c_int, [vp, vp]),
    "pf_dynamic_step": (c_int, [vp, c_dbl, c_dbl, c_int, P(c_int), P(c_int), P(c_dbl)]),
    "pf_dynamic_get": (c_int, [vp, vp, vp, vp]),
    "pf_plastic_begin": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, P(c_dbl)]),
    "pf_plastic_increment": (c_int, [vp, c_dbl, vp, vp, c_int, c_dbl, c_int, c_dbl, P(c_int), P(c_int), P(c_dbl)]),
    "pf_plastic_get": (c_int, [vp, vp, c_i64, c_int, vp]),
    "pf_gather": (c_int, [vp, vp, vp]),
    "pf_matvec": (c_int, [vp, vp, vp]),
    "pf_scatter": (c_int, [vp, vp, vp]),
    "pf_apply": (c_int, [vp, vp, vp]),
    "pf_dot": (c_int, [vp, vp, vp, P(c_dbl)]),
    "pf_norm": (c_int, [vp, vp, P(c_dbl)]),
    "pf_sum": (c_int, [vp, vp, P(c_dbl)]),
    "pf_centroid_stress": (c_int, [vp, c_i64, c_dbl, c_dbl, vp]),
    "pf_point_stress": (c_int, [vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, vp]),
    "pf_halo_transport": (c_int, [vp]),
    "pf_get_last_solve_ms": (c_int, [vp, P(c_dbl)]),
    "pf_set_profile": (c_int, [vp, c_int]),
    "pf_reset_profile": (c_int, [vp]),
    "pf_get_kernel_ms": (c_int, [vp, c_int, P(c_dbl), P(c_i64)]),
    "pf_kernel_launches": (c_i64, [vp]),
    "pf_vm_explicit_begin": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, vp]),
    "pf_vm_explicit_set_form": (c_int, [vp, c_int]),
    "pf_vm_explicit_steps": (c_int, [vp, c_int, P(c_dbl)]),
    "pf_vm_explicit_get": (c_int, [vp, vp, vp, vp, vp]),
    "pf_measure_fp64": (c_int, [vp, P(c_dbl)]),
    "pf_measure_fp64_tensor": (c_int, [vp, P(c_dbl)]),
    "pf_measure_matvec": (c_int, [vp, c_int, P(c_dbl)]),
    "pf_measure_hbm_read": (c_int, [vp, P(c_dbl)]),
    "pf_device_info": (c_int, [vp, P(c_int), P(c_i64), P(c_i64)]),
    # B. host helpers
    "pf_calc_nels_pp": (None, [c_i64, c_int, c_int, P(c_i64), P(c_i64)]),
    "pf_calc_neq_pp": (None, [c_i64, c_int, c_int, P(c_i64), P(c_i64)]),
    "pf_read_psize": (c_int, [C.c_char_p, c_int, c_int, P(c_i64), P(c_i64)]),
    "pf_p121_sizes": (c_int, [c_int, c_int, c_int, c_int, P(c_i64), P(c_i64), P(c_i64)]),
    "pf_p123_sizes": (c_int, [c_int, c_int, c_int, P(c_i64), P(c_i64), P(c_i64)]),
    "pf_cube_elements": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_i64, c_int, vp, vp]),
    "pf_cube_rest": (c_int, [c_int, c_int, c_int, c_int, c_int, c_i64, vp]),
    "pf_p121_loads": (c_int, [c_int, c_int, c_int, c_dbl, c_dbl, c_int, vp, vp]),
    "pf_form_nf": (c_int, [c_i64, c_int, c_i64, vp, vp, P(c_i64)]),
    "pf_find_g": (c_int, [c_int, c_int, c_i64, c_i64, vp, vp, vp]),
    "pf_load": (c_int, [c_int, c_i64, c_i64, vp, vp, vp, c_i64, c_i64, vp]),
    "pf_abaqus2sg": (c_int, [c_int, c_i64, vp]),
    "pf_read_dat": (c_int, [C.c_char_p, c_int, P(DeckInfo)]),
    "pf_read_d": (c_int, [C.c_char_p, c_i64, c_i64, c_int, vp, vp]),
    "pf_read_d_mat": (c_int, [C.c_char_p, c_i64, c_i64, c_int, vp, vp, vp]),
    "pf_read_mat": (c_int, [C.c_char_p, c_int, c_int, vp]),
    "pf_read_bnd": (c_int, [C.c_char_p, c_i64, c_int, vp]),
    "pf_read_lds": (c_int, [C.c_char_p, c_i64, c_int, vp, vp]),
    "pf_read_fix": (c_int, [C.c_char_p, c_i64, vp, vp, vp]),
    "pf_coords_pp": (c_int, [c_int, c_i64, c_i64, vp, vp, vp]),
    "pf_write_deck_p121": (c_int, [C.c_char_p, c_int, c_i64, c_i64, c_i64, c_int, c_i64, c_dbl, c_dbl, c_dbl, c_int,
                                   vp, vp, vp, vp, vp]),
    "pf_write_deck_scalar": (c_int, [C.c_char_p, P(DeckInfo), vp, vp, vp]),
    "pf_calc_nodes_pp": (None, [c_i64, c_int, c_int, P(c_i64), P(c_i64)]),
    "pf_calc_npes_pp": (c_int, [c_int]),
    "pf_nodal_values": (c_int, [c_int, c_i64, vp, c_i64, c_i64, vp, c_i64, c_i64, vp]),
    "pf_write_ensi": (c_int, [C.c_char_p, c_int, c_i64, vp, c_int]),
    "pf_make_ggl": (c_int, [c_int, c_i64, vp, c_i64, c_int, c_int, vp, c_i64, vp, vp, P(c_i64)]),
    "pf_write_geo_bin": (c_int, [C.c_char_p, c_int, c_i64, c_i64, vp, vp]),
    "pf_geo_bin_sizes": (c_int, [C.c_char_p, P(c_i64), P(c_i64), P(c_int)]),
    "pf_read_geo_bin": (c_int, [C.c_char_p, c_i64, c_i64, c_int, vp, vp]),
    "pf_ensi2sg": (c_int, [c_int, c_i64, vp]),
    "pf_make_put_tables": (c_int, [c_int, c_i64, vp, vp, vp, vp, vp, vp, vp, vp, P(c_i64)]),
    "pf_make_acc_chunks": (c_int, [c_i64, c_int, c_i64, vp, vp]),
}

# the reference's existing CUDA boundary (include/parafem_xx3_compat.h = xx3.f90:56-148): scalars by reference
PI = P(c_int)
XX3_SIGNATURES = {
    "set_gpu": (c_int, [PI]),
    "allocate_memory_on_gpu": (c_int, [PI, PI, P(vp)]),
    "free_memory_on_gpu": (c_int, [P(vp)]),
    "copy_data_to_gpu": (c_int, [PI, PI, vp, P(vp)]),
    "copy_data_from_gpu": (c_int, [PI, PI, vp, P(vp)]),
    "matrix_vector_multiplies": (c_int, [PI, PI, PI, P(vp), P(vp), P(vp)]),
}

_lib = None


def lib():
    """Load the shared object once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PfError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the device path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(XX3_SIGNATURES.items()):
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def ptr(a):
    """Raw pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be contiguous"
    return a.ctypes.data_as(vp)


def check(rc, handle=None, what=""):
    if rc != 0:
        buf = C.create_string_buffer(1024)
        lib().pf_last_error(handle, buf, 1024)
        raise PfError(f"{what} failed (status {rc}): {buf.value.decode(errors='replace')}")


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
