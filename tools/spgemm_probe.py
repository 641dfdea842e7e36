"""Developer tool: SpGEMM A*A on several matrix families with the phase split."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402


def run(name, A, reps=3):
    flop = 2.0 * g4s_b200.compute_flop(A, A)
    Cm = g4s_b200.HashSpGEMM(A, A)
    nnzc = Cm.nnz
    Cm.make_empty()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g4s_b200.HashSpGEMM(A, A).make_empty()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ph = (C.c_double * 4)()
    g4s_b200.lib().g4s_spgemm_last_phase_ms(ph)  # zeros unless g4s_spgemm_set_phase_timing(1) was called
    print("%-12s rows=%d nnzA=%d intprod=%.3g nnzC=%d : %.3f ms %.1f GFLOP/s  (bin %.3f sym %.3f scan %.3f num %.3f)"
          % (name, A.rows, A.nnz, flop / 2, nnzc, ms, flop / ms / 1e6, ph[0], ph[1], ph[2], ph[3]), flush=True)


if os.environ.get("SPGEMM_ESC"):  # time the expand - sort - compress path instead
    g4s_b200.HashSpGEMM = g4s_b200.OuterSpGEMM
which = sys.argv[1:] or ["lap3d64", "lap3d100", "rmat16", "rmat18", "lap2d2048"]
for w in which:
    if w.startswith("lap3d"):
        run(w, g4s_b200.CSR.laplacian3d27(int(w[5:])))
    elif w.startswith("lap2d"):
        run(w, g4s_b200.CSR.laplacian2d(int(w[5:])))
    elif w.startswith("rmat"):
        run(w, g4s_b200.CSR.rmat(int(w[4:]), 16, seed=1))
