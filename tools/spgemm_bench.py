"""Developer tool: time SpGEMM A*A on the 2-D 5-point Laplacian (config 4) and print the phase split."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
A = g4s_b200.CSR.laplacian2d(n)
flop = 2.0 * g4s_b200.compute_flop(A, A)
for _ in range(2):
    g4s_b200.HashSpGEMM(A, A).make_empty()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    Cm = g4s_b200.HashSpGEMM(A, A)
    nnzc = Cm.nnz
    Cm.make_empty()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ph = (C.c_double * 4)()
g4s_b200.lib().g4s_spgemm_last_phase_ms(ph)  # zeros unless g4s_spgemm_set_phase_timing(1) was called
print("n=%d rows=%d nnzA=%d nnzC=%d  %.3f ms  %.1f GFLOP/s  phases(ms): bin %.3f sym %.3f scan+alloc %.3f num %.3f"
      % (n, A.rows, A.nnz, nnzc, ms, flop / ms / 1e6, ph[0], ph[1], ph[2], ph[3]))
