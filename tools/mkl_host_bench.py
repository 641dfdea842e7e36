"""Developer tool: the host-pointer mkl() twin on configs[3] (A x A, 2-D 5-point Laplacian n = 2048), phase timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
A = g4s_b200.CSR.laplacian2d(n).to_host()
B = g4s_b200.CSR(A.rows, A.cols, A.rowptr.copy(), A.colids.copy(), A.values.copy())  # a deep copy, as `B = A` in the driver
for name, rhs in (("same arrays", A), ("copied B", B)):
    for rep in range(4):
        t = g4s_b200.Timings()
        g4s_b200.mkl(A, rhs, t)
    print("%-12s total %.1f ms: create %.1f, spmm %.2f, export %.1f, destroy %.2f" %
          (name, t.total * 1e3, t.create * 1e3, t.spmm * 1e3, t.export_csr * 1e3, t.destroy * 1e3), flush=True)
