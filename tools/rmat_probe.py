"""Developer tool: SpMV on R-MAT, lane/variant sweep (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
combos = [(0, 0), (11, 0), (15, 0), (8, 0)] if (len(sys.argv) < 3 or sys.argv[2] == 'all') else [(0, 0)]
kind = sys.argv[3] if len(sys.argv) > 3 else 'rmat'
A = g4s_b200.CSR.rmat(scale, 16, seed=20240601) if kind == 'rmat' else (g4s_b200.CSR.laplacian2d(scale) if kind == 'lap2d' else g4s_b200.CSR.laplacian3d27(scale))
nbytes, flops = A.spmv_cost()
x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
for variant, lanes in combos:
    A.set_tuning(lanes, variant)
    for _ in range(3):
        A.spmv_device(x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        A.spmv_device(x.data_ptr(), y.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(kind + "%d nnz=%d variant=%d lanes=%d : %.3f ms %.0f GB/s" % (scale, A.nnz, variant, lanes, ms, nbytes / ms / 1e6), flush=True)
