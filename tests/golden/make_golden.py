"""Generates tests/golden/*.npz and sym_pattern.mtx by running the REFERENCE's own code
(oracle/_ref/libg4s_ref.so = /root/reference/mm/inc compiled unmodified, see oracle/Makefile).
Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
The GPU box has no reference tree; these fixtures are what pins the oracle (and the CUDA path) there."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from matrices import laplacian_2d, laplacian_3d_27, powerlaw_csr, random_csr, tridiag3  # noqa: E402
from oracle.binding import Ref, build  # noqa: E402


def main():
    build()
    ref = Ref()
    assert ref.available, "needs /root/reference"
    cases = {
        "tridiag3": (tridiag3(),) * 2,
        "lap2d_24": (laplacian_2d(24),) * 2,
        "lap3d_5": (laplacian_3d_27(5),) * 2,
        "rand_rect": (random_csr(90, 140, 0.06, 31, empty_rows=True), random_csr(140, 70, 0.05, 32, empty_rows=True)),
        "powerlaw": (powerlaw_csr(1500, 33, max_deg=120),) * 2,
    }
    out = {}
    for name, (A, B) in cases.items():
        crpt, ccol, cval, _ = ref.hash_spgemm(A, B, variant=0)
        out.update({name + "__am": A[0], name + "__ak": A[1], name + "__bn": B[1],
                    name + "__arpt": A[2], name + "__acol": A[3], name + "__aval": A[4],
                    name + "__brpt": B[2], name + "__bcol": B[3], name + "__bval": B[4],
                    name + "__crpt": crpt, name + "__ccol": ccol, name + "__cval": cval})
    np.savez_compressed(os.path.join(HERE, "spgemm_golden.npz"), **out)

    rng = np.random.default_rng(2024)
    n = 12
    lower = [(i, j) for i in range(n) for j in range(i + 1) if rng.random() < 0.3 or i == j]
    path = os.path.join(HERE, "sym_pattern.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n% golden input for CSR::construct\n")
        f.write("%d %d %d\n" % (n, n, len(lower)))
        for i, j in lower:
            f.write("%d %d\n" % (i + 1, j + 1))
    _, _, rpt, col, val = ref.csr_construct(path)
    gn, gm = 40, 400
    start = np.sort(rng.integers(0, gn, gm))
    end = rng.integers(0, gn, gm)
    w = rng.uniform(0, 1, gm)
    _, _, grpt, gcol, gval = ref.csr_from_graph(gn, start, end, w)
    np.savez_compressed(os.path.join(HERE, "formats_golden.npz"), mtx_rpt=rpt, mtx_col=col, mtx_val=val,
                        g_n=gn, g_start=start, g_end=end, g_w=w, g_rpt=grpt, g_col=gcol, g_val=gval)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
