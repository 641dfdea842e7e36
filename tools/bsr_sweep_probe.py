"""Times the BSR SpMM kernels on BASELINE configs[4] at n^3 nodes (default 128): row-wise K-packed kernel vs the
sliding-window sweep.  CUDA events, 3 warm-ups."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
from g4s_b200 import bsr  # noqa: E402
from g4s_b200._lib import check  # noqa: E402
from g4s_b200.dist import _DevArray  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    L = g4s_b200.lib()
    P = g4s_b200.CSR.laplacian3d27(n)
    rp, ci, va = P.device_arrays()
    nb, mb = P.nnz, P.rows
    vals = torch.as_tensor(_DevArray(va, nb, "<f8"), device="cuda")
    J = torch.ones(3, 3, dtype=torch.float64, device="cuda")
    I3 = torch.eye(3, dtype=torch.float64, device="cuda")
    diag = (vals > 0).double()[:, None, None]
    blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
    del diag
    g = torch.Generator(device="cuda").manual_seed(777)
    B = torch.rand(mb * 192, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    C1, C2 = torch.empty_like(B), torch.empty_like(B)
    nbytes = 76.0 * nb + 4 * (mb + 1) + 2 * 8.0 * 192 * mb
    flops = 2.0 * 9 * nb * 64

    def timeit(fn):
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

    out = {"n": n, "blocks": nb, "algorithmic_GB": nbytes / 1e9}
    if os.environ.get("PROBE_SKIP_KPACK") != "1":
        ms = timeit(lambda: check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                                        C.c_void_p(blocks.data_ptr()), C.c_int(64), C.c_void_p(B.data_ptr()),
                                                        C.c_void_p(C1.data_ptr()), C.c_void_p(0))))
        out["kpack_ms"] = ms
        out["kpack_gbs"] = nbytes / ms / 1e6
    t0 = time.perf_counter()
    p0, p1 = int(os.environ.get("PROBE_P0", "4")), int(os.environ.get("PROBE_P1", "4"))
    strips = bsr.grid_pencil_strips(n, n, 0, n, p0, p1)
    out["patch"] = [p0, p1]
    plan = bsr.BsrPlan(mb, mb, rp, ci, strips)
    out["plan_create_s"] = time.perf_counter() - t0
    out["set_values_ms"] = timeit(lambda: plan.set_values(blocks.data_ptr()))
    out["plan"] = plan.info()
    ms = timeit(lambda: plan.spmm(B.data_ptr(), C2.data_ptr()))
    out["sweep_ms"] = ms
    out["sweep_gbs"] = nbytes / ms / 1e6
    out["sweep_tflops"] = flops / ms / 1e9
    if "kpack_ms" in out:
        out["max_abs_diff_vs_kpack"] = float((C1 - C2).abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
