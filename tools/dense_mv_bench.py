"""Developer tool: the four dense mv entry points on device-resident buffers (g4s_dense_mv_device), CUDA-event times."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
from g4s_b200._lib import check  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = g4s_b200.lib()
g = torch.Generator(device="cuda").manual_seed(5)
A = torch.rand(dim * dim, dtype=torch.float64, device="cuda", generator=g)
B = torch.ones(dim, dtype=torch.float64, device="cuda")
Cv = torch.empty(dim, dtype=torch.float64, device="cuda")
for op, name, nbytes in ((0, "dgemv", 8.0 * dim * dim + 16.0 * dim), (1, "dsymv", 4.0 * dim * (dim + 1) + 16.0 * dim),
                         (2, "dtrmv", 4.0 * dim * (dim + 1) + 16.0 * dim), (3, "dspmv", 4.0 * dim * (dim + 1) + 16.0 * dim)):
    def run():
        check(L.g4s_dense_mv_device(C.c_int(op), C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()),
                                    C.c_void_p(Cv.data_ptr()), C.c_int(dim), C.c_void_p(0)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%s dim %d: %.4f ms  %.0f GB/s" % (name, dim, ms, nbytes / ms / 1e6))
