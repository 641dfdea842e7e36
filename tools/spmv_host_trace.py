"""Developer tool: g4s_spmv_host on the n^3 27-point Laplacian with pinned host vectors; G4S_SPMV_HOST_TRACE=1 prints the
per-block completion times of the upload / product / download streams."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
A = g4s_b200.CSR.laplacian3d27(n)
hx = torch.ones(A.cols, dtype=torch.float64).pin_memory()
hy = torch.empty(A.rows, dtype=torch.float64).pin_memory()
L = g4s_b200.lib()


def step():
    rc = L.g4s_spmv_host(A.handle, C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr()))
    assert rc == 0


for _ in range(8):
    step()
t0 = time.perf_counter()
for _ in range(10):
    step()
print("e2e %.3f ms per product" % ((time.perf_counter() - t0) / 10 * 1e3))
