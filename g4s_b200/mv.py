"""The four entry points of the reference's mv/ driver (mv/mv.c:6-27), same names and argument meaning:
(A, B, C, dim) with A a dim*dim row-major-filled buffer read column-major by the BLAS call, B the input vector
(overwritten by dtrmv) and C the output.  Host numpy buffers; the work runs on the GPU in libg4s_b200.so."""
import ctypes as C

import numpy as np

from ._lib import f64p, lib


def _call(name, A, B, Cv, dim):
    for a in (A, B, Cv):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
            raise TypeError("A, B, C must be C-contiguous float64 numpy arrays (they are modified in place)")
    if A.size != dim * dim or B.size != dim or Cv.size != dim:
        raise ValueError("A must hold dim*dim doubles, B and C dim doubles")
    getattr(lib(), name)(A.ctypes.data_as(f64p), B.ctypes.data_as(f64p), Cv.ctypes.data_as(f64p), C.c_int(dim))


def matrix_multiply_dsymv(A, B, Cv, dim):
    _call("matrix_multiply_dsymv", A, B, Cv, dim)


def matrix_multiply_dtrmv(A, B, Cv, dim):
    _call("matrix_multiply_dtrmv", A, B, Cv, dim)


def matrix_multiply_sspmv(A, B, Cv, dim):
    _call("matrix_multiply_sspmv", A, B, Cv, dim)


def matrix_multiply_dgemv(A, B, Cv, dim):
    _call("matrix_multiply_dgemv", A, B, Cv, dim)
