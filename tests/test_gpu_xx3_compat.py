"""The reference's existing CUDA boundary (xx3.f90:56-148 -> xx3/cuda_helpers.cu:183-368), exported by
libparafem_b200.so under the same names: the call sequence of xx3.f90:409-530 (allocate, copy the element
matrices once, then per iteration copy pmul in / multiply / copy utemp out) against the oracle's MATMUL."""
import ctypes as C

import numpy as np
import pytest

import oracle
from parafem_b200._lib import XX3_SIGNATURES, lib, ptr

pytestmark = pytest.mark.gpu


def ci(v):
    return C.byref(C.c_int(v))


class Xx3Gpu:
    """What xx3.f90 does with its three device pointers: pmul_pp -> device_lhs_vector, the products come back from
    device_rhs_vector (xx3.f90:496-529).  L = the library that provides the six symbols."""

    def __init__(self, n_mat, n_row, n_col, L=None):
        self.L, self.shape = (L if L is not None else lib()), (n_mat, n_row, n_col)
        self.d_km, self.d_rhs, self.d_lhs = C.c_void_p(), C.c_void_p(), C.c_void_p()
        dev = C.c_int(0)
        assert self.L.set_gpu(C.byref(dev)) == 0
        assert self.L.allocate_memory_on_gpu(ci(n_mat * n_row * n_col), ci(8), C.byref(self.d_km)) == 0      # xx3.f90:423
        assert self.L.allocate_memory_on_gpu(ci(n_mat * n_col), ci(8), C.byref(self.d_lhs)) == 0              # :433
        assert self.L.allocate_memory_on_gpu(ci(n_mat * n_row), ci(8), C.byref(self.d_rhs)) == 0              # :443

    def upload(self, km):
        n_mat, n_row, n_col = self.shape
        assert self.L.copy_data_to_gpu(ci(n_mat * n_row * n_col), ci(8), ptr(km), C.byref(self.d_km)) == 0

    def multiply(self, pmul):
        n_mat, n_row, n_col = self.shape
        out = np.empty((n_mat, n_row))
        assert self.L.copy_data_to_gpu(ci(n_mat * n_col), ci(8), ptr(pmul), C.byref(self.d_lhs)) == 0         # :496
        assert self.L.matrix_vector_multiplies(ci(n_mat), ci(n_row), ci(n_col), C.byref(self.d_lhs), C.byref(self.d_km),
                                               C.byref(self.d_rhs)) == 0                                      # :509
        assert self.L.copy_data_from_gpu(ci(n_mat * n_row), ci(8), ptr(out), C.byref(self.d_rhs)) == 0        # :525
        return out

    def close(self):
        for d in (self.d_km, self.d_rhs, self.d_lhs):
            assert self.L.free_memory_on_gpu(C.byref(d)) == 0


@pytest.mark.parametrize("n_mat,ntot", [(1000, 60), (125, 60), (4099, 24), (7, 24), (50001, 8), (3, 8)])
def test_element_sizes_of_p121_p123_equal_matmul(n_mat, ntot):
    rng = np.random.RandomState(ntot + n_mat)
    km = rng.randn(n_mat, ntot, ntot)            # storkm_pp(i,j,iel) = km[iel, j, i]
    g = Xx3Gpu(n_mat, ntot, ntot)
    g.upload(km)
    for _ in range(2):                            # two "iterations": the matrices stay resident
        pmul = rng.randn(n_mat, ntot)
        assert np.array_equal(g.multiply(pmul), oracle.matvec(km, pmul))
    g.close()


@pytest.mark.parametrize("n_mat,n_row,n_col", [(513, 12, 7), (100, 30, 30), (9, 1, 5)])
def test_any_shape_equals_the_column_sweep(n_mat, n_row, n_col):
    rng = np.random.RandomState(1)
    km = rng.randn(n_mat, n_col, n_row)           # column-major (n_row, n_col) per matrix
    pmul = rng.randn(n_mat, n_col)
    g = Xx3Gpu(n_mat, n_row, n_col)
    g.upload(km)
    ref = np.zeros((n_mat, n_row))
    for j in range(n_col):                         # u(i) = u(i) + a(i,j)*p(j), j ascending
        ref = ref + km[:, j, :] * pmul[:, j:j + 1]
    assert np.array_equal(g.multiply(pmul), ref)
    g.close()


def _ref_handle(nofma):
    L = oracle.ref_lib(nofma=nofma)
    for name, (res, args) in XX3_SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


@pytest.mark.parametrize("n_mat,ntot", [(2000, 60), (3001, 24), (10007, 8), (77, 12)])
def test_against_the_reference_binary(n_mat, ntot):
    """oracle/_ref: the reference's OWN cuda_helpers.cu (xx3's MultiMatVecMultiply1, cuda_helpers.cu:144-177),
    compiled by oracle/Makefile from where it lies, run on this GPU through the same six symbols.  Built with
    --fmad=false (the rounding of the Fortran MATMUL loop the GPU path of xx3 replaces) it must give the SAME BITS as
    this library; built as the reference builds it (nvcc default: the multiply-add is contracted to an FMA) the two
    agree to rounding."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    rng = np.random.RandomState(ntot)
    km = rng.randn(n_mat, ntot, ntot)
    pmul = rng.randn(n_mat, ntot)
    ours = Xx3Gpu(n_mat, ntot, ntot)
    ours.upload(km)
    u = ours.multiply(pmul)
    ours.close()
    assert np.array_equal(u, oracle.matvec(km, pmul))
    for nofma in (True, False):
        ref = Xx3Gpu(n_mat, ntot, ntot, L=_ref_handle(nofma))
        ref.upload(km)
        u_ref = ref.multiply(pmul)
        ref.close()
        if nofma:
            assert np.array_equal(u, u_ref)
        else:
            assert np.abs(u - u_ref).max() <= 1e-13 * np.abs(u_ref).max()       # FMA contraction: last-bit differences


def test_failures_return_exit_failure():
    L = lib()
    bad = C.c_int(99)
    assert L.set_gpu(C.byref(bad)) == 1            # EXIT_FAILURE + a printf message, as the reference
    d = C.c_void_p()
    assert L.matrix_vector_multiplies(ci(0), ci(8), ci(8), C.byref(d), C.byref(d), C.byref(d)) == 1
    ok = C.c_int(0)
    assert L.set_gpu(C.byref(ok)) == 0
