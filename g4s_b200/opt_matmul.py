"""Host-side mirror of DeePMD-kit's OptMatmul op (deepmd/source/op/opt_matmul.cc:24-62; used at
deepmd/deepmd/utils/network.py:234,239) and of its registered gradient (deepmd/source/op/_opt_matmul_grad.py)."""
import ctypes as C

import numpy as np

from ._lib import check, f64p, lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def opt_matmul(xx, w):
    """res = op_module.opt_matmul(xx, w): xx [M,N] times w [N,K] -> [M,K], float64 (host arrays; copies inside the call)."""
    xx, w = _f64(xx), _f64(w)
    if xx.ndim != 2 or w.ndim != 2 or xx.shape[1] != w.shape[0]:
        raise ValueError("opt_matmul: xx [M,N] and w [N,K] expected")
    M, N = xx.shape
    K = w.shape[1]
    res = np.empty((M, K), dtype=np.float64)
    check(lib().g4s_opt_matmul(C.c_int(M), C.c_int(N), C.c_int(K), xx.ctypes.data_as(f64p), w.ctypes.data_as(f64p),
                               res.ctypes.data_as(f64p)))
    return res


def opt_matmul_device(M, N, K, xx_ptr, w_ptr, res_ptr, stream=None):
    check(lib().g4s_opt_matmul_device(C.c_int(M), C.c_int(N), C.c_int(K), C.c_void_p(xx_ptr), C.c_void_p(w_ptr),
                                      C.c_void_p(res_ptr), C.c_void_p(stream.cuda_stream if stream is not None else 0)))


def opt_matmul_grad_device(M, N, K, xx_ptr, w_ptr, grad_ptr, dxx_ptr, dw_ptr, stream=None):
    """_opt_matmul_grad: dxx = grad w^T, dw = xx^T grad (device pointers; either output may be 0)."""
    check(lib().g4s_opt_matmul_grad_device(C.c_int(M), C.c_int(N), C.c_int(K), C.c_void_p(xx_ptr), C.c_void_p(w_ptr),
                                           C.c_void_p(grad_ptr), C.c_void_p(dxx_ptr), C.c_void_p(dw_ptr),
                                           C.c_void_p(stream.cuda_stream if stream is not None else 0)))
