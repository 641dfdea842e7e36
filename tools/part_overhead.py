"""Developer tool: cost of the partitioned-x kernel variant when every chunk is local (world = 1)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
from g4s_b200._lib import check  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 318
A = g4s_b200.CSR.laplacian3d27(n)
x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
parts = (C.c_void_p * 1)(x.data_ptr())
cuts = (C.c_int * 2)(0, A.rows)
L = g4s_b200.lib()


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


nbytes, _ = A.spmv_cost()
t0 = timeit(lambda: A.spmv_device(x.data_ptr(), y.data_ptr()))
t1 = timeit(lambda: check(L.g4s_spmv_partitioned_device(A.handle, C.c_int(1), C.c_int(0), parts, cuts,
                                                        C.c_void_p(y.data_ptr()), C.c_void_p(0), C.c_ulonglong(0),
                                                        None, C.c_void_p(0))))
print("n=%d plain %.4f ms (%.0f GB/s)   partitioned(world=1) %.4f ms (%.0f GB/s)" %
      (n, t0, nbytes / t0 / 1e6, t1, nbytes / t1 / 1e6))
