"""Developer tool: SpMV on R-MAT scale s (configs[2] at 24), device-timed; G4S_SPMV_HOT=0 disables the hot-column staging."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
A = g4s_b200.CSR.rmat(scale, 16, seed=20240601)
nbytes, flops = A.spmv_cost()
x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
for _ in range(3):
    A.spmv_device(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    A.spmv_device(x.data_ptr(), y.data_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("rmat %d nnz %d: %.4f ms %.0f GB/s (HOT=%s) checksum %.6e" % (scale, A.nnz, ms, nbytes / ms / 1e6, os.environ.get("G4S_SPMV_HOT"), float(y.sum())))
