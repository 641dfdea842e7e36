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
pencil = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else None  # ordered kernel: patch of the pencil order
planes = int(os.environ.get("BSR_PLANES", "0"))  # > 0: only the first `planes` mesh planes (one rank's slab of config 5)
P = g4s_b200.CSR.laplacian3d27(n, 0, planes * n * n) if planes else g4s_b200.CSR.laplacian3d27(n)
rp, ci, va = P.device_arrays()
nb = P.nnz
vals = torch.as_tensor(g4s_b200.dist._DevArray(va, nb, "<f8"), device="cuda")
J = torch.ones(3, 3, dtype=torch.float64, device="cuda")
I3 = torch.eye(3, dtype=torch.float64, device="cuda")
diag = (vals > 0).double()[:, None, None]
blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
mb, ncol = P.rows, 64
kb = min(n ** 3, mb + n * n) if planes else mb
B = torch.rand(kb * 3 * ncol, dtype=torch.float64, device="cuda") * 2 - 1
Cd = torch.empty(mb * 3 * ncol, dtype=torch.float64, device="cuda")
L = g4s_b200.lib()
check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))


order = None
if pencil:
    import numpy as np

    oh = np.empty(mb, dtype=np.int32)
    th = np.empty(-(-n // pencil[0]) * -(-n // pencil[1]) + 1, dtype=np.int32)
    nt = C.c_int()
    check(L.g4s_grid_pencil_order(C.c_int(n), C.c_int(n), C.c_int(planes or n), C.c_int(pencil[0]), C.c_int(pencil[1]),
                                  oh.ctypes.data_as(C.c_void_p), th.ctypes.data_as(C.c_void_p), C.byref(nt)))
    order = torch.from_numpy(oh).cuda()
    tiles = torch.from_numpy(th).cuda() if not os.environ.get("BSR_NO_TILES") else None


def run():
    if order is not None:
        check(L.g4s_bsr3_spmm64_ordered_device(C.c_int(mb), C.c_int(mb), C.c_void_p(rp), C.c_void_p(ci),
                                               C.c_void_p(blocks.data_ptr()), C.c_void_p(B.data_ptr()),
                                               C.c_void_p(Cd.data_ptr()), C.c_void_p(order.data_ptr()),
                                               C.c_void_p(tiles.data_ptr() if tiles is not None else 0), C.c_int(nt.value),
                                               C.c_void_p(0)))
        return
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
if order is not None:  # same result as the unordered kernel
    ref = torch.empty_like(Cd)
    check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                C.c_void_p(blocks.data_ptr()), C.c_int(ncol), C.c_void_p(B.data_ptr()),
                                C.c_void_p(ref.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    print("ordered %s: max |C - C_unordered| = %.3e" % (pencil, float((Cd - ref).abs().max())))
print("bsr n=%d variant=%d: %.3f ms  %.0f GB/s algorithmic  %.1f GFLOP/s  checksum %.6e" %
      (n, variant, ms, nbytes / ms / 1e6, 2.0 * 9 * nb * ncol / ms / 1e6, float(Cd.sum())))
