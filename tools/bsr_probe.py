"""Developer tool: one BSR 3x3 SpMM x 64 columns on a hex-mesh stencil pattern (variant from argv)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
import g4s_b200.dist  # noqa: E402,F401
from g4s_b200._lib import check  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
P = g4s_b200.CSR.laplacian3d27(n)
rp, ci, va = P.device_arrays()
nb = P.nnz
vals = torch.as_tensor(g4s_b200.dist._DevArray(va, nb, "<f8"), device="cuda")
J = torch.ones(3, 3, dtype=torch.float64, device="cuda")
I3 = torch.eye(3, dtype=torch.float64, device="cuda")
diag = (vals > 0).double()[:, None, None]
blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
mb, ncol = P.rows, 64
B = torch.rand(mb * 3 * ncol, dtype=torch.float64, device="cuda") * 2 - 1
Cd = torch.empty(mb * 3 * ncol, dtype=torch.float64, device="cuda")
L = g4s_b200.lib()
check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))


def run():
    check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                C.c_void_p(blocks.data_ptr()), C.c_int(ncol), C.c_void_p(B.data_ptr()),
                                C.c_void_p(Cd.data_ptr()), C.c_void_p(0)))


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
nbytes = 76.0 * nb + 4 * (mb + 1) + 2 * 8.0 * 3 * mb * ncol
print("bsr n=%d variant=%d: %.3f ms  %.0f GB/s algorithmic  %.1f GFLOP/s  checksum %.6e" %
      (n, variant, ms, nbytes / ms / 1e6, 2.0 * 9 * nb * ncol / ms / 1e6, float(Cd.sum())))
