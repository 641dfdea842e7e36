"""Developer tool: SpGEMM A*A on the 2-D 5-point Laplacian (configs[3]) with per-phase event times."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # > 1: the first 1/parts of A's rows times all of A (one rank's share)
B = g4s_b200.CSR.laplacian2d(n)
A = B if parts == 1 else g4s_b200.CSR.laplacian2d(n, 0, n * n // parts)
flop = 2.0 * g4s_b200.compute_flop(A, B)
lib = g4s_b200.lib()
for _ in range(3):
    g4s_b200.HashSpGEMM(A, B).make_empty()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    g4s_b200.HashSpGEMM(A, B).make_empty()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
lib.g4s_spgemm_set_phase_timing(1)
acc = [0.0] * 4
ph = (C.c_double * 4)()
for _ in range(reps):
    g4s_b200.HashSpGEMM(A, B).make_empty()
    lib.g4s_spgemm_last_phase_ms(ph)
    acc = [x + y for x, y in zip(acc, ph)]
lib.g4s_spgemm_set_phase_timing(0)
print("n=%d parts=%d %.4f ms %.1f GFLOP/s | bin %.4f sym %.4f scan+alloc %.4f num %.4f | env PF=%s FUSED_SCAN=%s"
      % (n, parts, ms, flop / ms / 1e6, *[x / reps for x in acc], os.environ.get("G4S_SPGEMM_MERGE_PF"), os.environ.get("G4S_SPGEMM_FUSED_SCAN")))
