"""A*A on an R-MAT graph (long rows: the dense-accumulator classes), phase times."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
A = g4s_b200.CSR.rmat(scale, 16, seed=20240601)
L = g4s_b200.lib()
flop = 2.0 * g4s_b200.compute_flop(A, A)
for _ in range(2):
    g4s_b200.HashSpGEMM(A, A).make_empty()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    Cm = g4s_b200.HashSpGEMM(A, A)
    nnzc = Cm.nnz
    Cm.make_empty()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
L.g4s_spgemm_set_phase_timing(1)
g4s_b200.HashSpGEMM(A, A).make_empty()
ph = (C.c_double * 4)()
L.g4s_spgemm_last_phase_ms(ph)
print("R-MAT %d: nnzA %d nnzC %d  %.3f ms %.1f GFLOP/s | bin %.3f sym %.3f scan %.3f num %.3f" % (scale, A.nnz, nnzc, ms, flop / ms / 1e6, *ph))
