"""Developer tool: device-timed numbers for the secondary configs.
  python tools/misc_bench.py rmat [scale=24]      SpMV on R-MAT (config 3)
  python tools/misc_bench.py bsr  [nodes=128]     BSR 3x3 SpMM x 64 columns on a hex mesh (config 5), all kernels
  python tools/misc_bench.py lap2d [n=1000]       SpMV on the 2-D 5-point Laplacian (config 1)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
import g4s_b200.dist  # noqa: E402,F401
from g4s_b200._lib import check  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def spmv_report(A, name):
    nbytes, flops = A.spmv_cost()
    x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
    y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
    for variant, lanes in ((0, 0), (9, 0), (0, 1), (0, 4), (0, 32)):
        A.set_tuning(lanes, variant)
        ms = timeit(lambda: A.spmv_device(x.data_ptr(), y.data_ptr()))
        print("%s rows=%d nnz=%d variant=%d lanes=%d : %.4f ms  %.1f GB/s  %.1f GFLOP/s" %
              (name, A.rows, A.nnz, variant, lanes, ms, nbytes / ms / 1e6, flops / ms / 1e6), flush=True)


def main():
    what = sys.argv[1]
    if what == "rmat":
        scale = int(sys.argv[2]) if len(sys.argv) > 2 else 24
        A = g4s_b200.CSR.rmat(scale, 16, seed=20240601)
        torch.cuda.synchronize()
        spmv_report(A, "rmat%d" % scale)
    elif what == "lap2d":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
        spmv_report(g4s_b200.CSR.laplacian2d(n), "lap2d_%d" % n)
    elif what == "bsr":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
        # block pattern = 27-point stencil on n^3 nodes; blocks: diagonal 26 I + J, off-diagonal -I - 0.1 J
        P = g4s_b200.CSR.laplacian3d27(n)
        rp, ci, va = P.device_arrays()
        nb = P.nnz
        vals = torch.as_tensor(g4s_b200.dist._DevArray(va, nb, "<f8"), device="cuda")
        J = torch.ones(3, 3, dtype=torch.float64, device="cuda")
        I3 = torch.eye(3, dtype=torch.float64, device="cuda")
        diag = (vals > 0).double()[:, None, None]
        blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
        mb, ncol = P.rows, 64
        g = torch.Generator(device="cuda").manual_seed(777)
        B = (torch.rand(mb * 3 * ncol, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
        Cd = torch.empty(mb * 3 * ncol, dtype=torch.float64, device="cuda")
        nbytes = 76.0 * nb + 4 * (mb + 1) + 2 * 8.0 * 3 * mb * ncol
        flops = 2.0 * 9 * nb * ncol
        L = g4s_b200.lib()
        for variant, name in ((1, "dfma"), (2, "dmma"), (3, "generic")):
            check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))
            ms = timeit(lambda: check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp),
                                                           C.c_void_p(ci), C.c_void_p(blocks.data_ptr()), C.c_int(ncol),
                                                           C.c_void_p(B.data_ptr()), C.c_void_p(Cd.data_ptr()),
                                                           C.c_void_p(0))), iters=5, warm=2)
            print("bsr3x3 nodes=%d^3 blocks=%d %s : %.3f ms  %.1f GB/s (algorithmic)  %.1f GFLOP/s" %
                  (n, nb, name, ms, nbytes / ms / 1e6, flops / ms / 1e6), flush=True)
        check(L.g4s_bsr_spmm_set_variant(C.c_int(0)))


if __name__ == "__main__":
    main()
