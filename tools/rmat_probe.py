"""R-MAT scale-24 SpMV (BASELINE configs[2]): device time per kernel shape."""
import json
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
out = {"rows": A.rows, "nnz": A.nnz, "GB": nbytes / 1e9}
for variant in [0] + [int(v) for v in sys.argv[2:]]:
    A.set_tuning(0, variant)
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
    out["variant %d" % variant] = {"ms": ms, "gbs": nbytes / ms / 1e6}
print(json.dumps(out))
