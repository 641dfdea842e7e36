"""Small end-to-end exercise of every kernel family (used under compute-sanitizer memcheck)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import g4s_b200  # noqa: E402
from g4s_b200._lib import check  # noqa: E402
from g4s_b200.dist import GpuOps  # noqa: E402
from matrices import laplacian_2d, laplacian_3d_27, powerlaw_csr, random_csr  # noqa: E402

L = g4s_b200.lib()
rng = np.random.default_rng(0)
# SpMV: every family / shape, host and borrowed-device handles (odd sizes: unaligned tails of the bulk copies)
for A in (laplacian_2d(37), laplacian_3d_27(9), powerlaw_csr(3001, 5), random_csr(777, 333, 0.03, 2, empty_rows=True)):
    M = g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4])
    x = rng.uniform(-1, 1, A[1])
    for v in (0, 6, 12, 15, 8):
        M.set_tuning(0, v)
        M.spmv(x)
    rp, ci, va = (torch.from_numpy(a).cuda() for a in (A[2], A[3], A[4]))
    h = C.c_void_p()
    check(L.g4s_csr_create_device(C.byref(h), C.c_int(A[0]), C.c_int(A[1]), C.c_void_p(rp.data_ptr()), C.c_void_p(ci.data_ptr()),
                                  C.c_void_p(va.data_ptr()), C.c_void_p(0)))
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty(A[0], dtype=torch.float64, device="cuda")
    check(L.g4s_spmv_device(h, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    L.g4s_csr_destroy(h)
# generators
g4s_b200.CSR.laplacian3d27(7, 10, 200).to_host()
g4s_b200.CSR.laplacian2d(9).to_host()
g4s_b200.CSR.rmat(10, 8, seed=3).to_host()
# SpGEMM: all classes
import scipy.sparse as sp  # noqa: E402
from matrices import to_tuple  # noqa: E402

band = to_tuple(sp.diags([rng.uniform(-1, 1, 3000) for _ in range(71)], list(range(-35, 36)), shape=(3000, 3000)))
for A in (laplacian_2d(30), laplacian_3d_27(8), powerlaw_csr(5000, 7, max_deg=3000), band):
    M = g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4])
    g4s_b200.HashSpGEMM(M, M).to_host()
# split / compact / gather
ops = GpuOps()
A = laplacian_3d_27(8)
M = g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4])
d, o = ops.split(M, 100, 300)
ops.compact(o)
# dense mv
dim = 130
Ad, Bd, Cd = rng.uniform(-1, 1, dim * dim), rng.uniform(-1, 1, dim), np.zeros(dim)
for name in ("dgemv", "dsymv", "dtrmv", "sspmv"):
    getattr(g4s_b200.mv, "matrix_multiply_" + name)(Ad, Bd, Cd, dim)
# BSR + EBE
mb = 101
pat = (sp.random(mb, mb, density=0.05, random_state=rng, format="csr") + sp.identity(mb, format="csr")).tocsr()
blocks = torch.from_numpy(rng.uniform(-1, 1, pat.nnz * 9)).cuda()
t = [torch.from_numpy(a.astype(np.int32)).cuda() for a in (pat.indptr, pat.indices)]
Bm = torch.rand(mb * 3 * 64, dtype=torch.float64, device="cuda")
Cm = torch.empty_like(Bm)
for v in (1, 2, 3):
    L.g4s_bsr_spmm_set_variant(C.c_int(v))
    check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(t[0].data_ptr()), C.c_void_p(t[1].data_ptr()),
                                C.c_void_p(blocks.data_ptr()), C.c_int(64), C.c_void_p(Bm.data_ptr()), C.c_void_p(Cm.data_ptr()),
                                C.c_void_p(0)))
L.g4s_bsr_spmm_set_variant(C.c_int(0))
nel, neq = 50, 200
kd = torch.rand(nel * 576, dtype=torch.float64, device="cuda")
dd = torch.randint(0, neq, (nel * 24,), dtype=torch.int32, device="cuda")
ud, Au = torch.rand(neq, dtype=torch.float64, device="cuda"), torch.zeros(neq, dtype=torch.float64, device="cuda")
check(L.g4s_ebe_matvec_device(C.c_int(nel), C.c_int(24), C.c_void_p(kd.data_ptr()), C.c_void_p(dd.data_ptr()),
                              C.c_void_p(ud.data_ptr()), C.c_void_p(Au.data_ptr()), C.c_void_p(0)))
torch.cuda.synchronize()
print("sanity_small: all kernel families ran;", L.g4s_kernel_launch_count(), "launches")
