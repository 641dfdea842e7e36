"""A*A on the 3-D 27-point Laplacian (rows of 729 products / 125 columns: the warp-per-row hash class), phase times."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
A = g4s_b200.CSR.laplacian3d27(n)
L = g4s_b200.lib()
flop = 2.0 * g4s_b200.compute_flop(A, A)
for _ in range(3):
    g4s_b200.HashSpGEMM(A, A).make_empty()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g4s_b200.HashSpGEMM(A, A).make_empty()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
L.g4s_spgemm_set_phase_timing(1)
g4s_b200.HashSpGEMM(A, A).make_empty()
ph = (C.c_double * 4)()
L.g4s_spgemm_last_phase_ms(ph)
print("27-point n=%d: %.3f ms %.1f GFLOP/s | bin %.3f sym %.3f scan %.3f num %.3f" % (n, ms, flop / ms / 1e6, *ph))
