"""The reference's two drivers (mv/mv.c main, mm/src/mkl_spgemm.cpp main) rebuilt on the C ABI:
g4s_b200/bin/g4s_mv and g4s_b200/bin/g4s_spgemm (g4s_b200/csrc/drivers/)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "g4s_b200", "bin")
MTX = os.path.join(ROOT, "tests", "golden", "sym_pattern.mtx")


def test_drivers_build_and_fail_loudly_without_a_gpu():
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "g4s_b200", "csrc"), "drivers"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert os.path.exists(os.path.join(BIN, "g4s_mv")) and os.path.exists(os.path.join(BIN, "g4s_spgemm"))
    import torch

    if not torch.cuda.is_available():  # no CPU fallback: the drivers must report the missing device, not compute
        out = subprocess.run([os.path.join(BIN, "g4s_spgemm"), MTX], capture_output=True, text=True)
        assert out.returncode != 0 and "no CUDA device" in out.stderr
        out = subprocess.run([os.path.join(BIN, "g4s_mv"), MTX], capture_output=True, text=True)
        assert out.returncode != 0 and "failed" in out.stderr


@pytest.mark.gpu
def test_spgemm_driver_matches_oracle(oracle):
    A = oracle.mm_construct(MTX)
    rpt, col, val = oracle.hash_spgemm(A, A)
    total, _ = oracle.intprod(A[2], A[3], A[2])
    out = subprocess.run([os.path.join(BIN, "g4s_spgemm"), MTX], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C: %d x %d, nnz %d" % (A[0], A[1], len(col)) in out.stdout
    assert "total flop %f" % (2.0 * total) in out.stdout
    for phase in ("create", "spmm", "convert", "order", "export_csr", "destroy", "sum_total"):  # Timings::print
        assert re.search(r"^\s+%s\s+[0-9.]+ms" % phase, out.stdout, re.M), phase


@pytest.mark.gpu
def test_mv_driver_results_match_the_reference_functions(tmp_path, oracle):
    """mv/mv.c:59-94 end to end: same file, same rand() sequence (glibc, default seed), same order of the four calls —
    dtrmv overwrites B, so sspmv and dgemv see its result.  The GPU driver's four vectors against the reference's own
    matrix_multiply_* (oracle/_ref/libmv_ref.so = unmodified mv/mv.c on OpenBLAS; the oracle's restatement when the
    prebuilt library is absent)."""
    import ctypes

    rng = np.random.default_rng(1)
    dim, nnz = 300, 4000
    rows, cols = rng.integers(1, dim + 1, nnz), rng.integers(1, dim + 1, nnz)
    p = str(tmp_path / "pat.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n% comment\n")
        f.write("%d %d %d\n" % (dim, dim, nnz))
        for r, c in zip(rows, cols):
            f.write("%d %d\n" % (r, c))
    dump = str(tmp_path / "vectors.bin")
    out = subprocess.run([os.path.join(BIN, "g4s_mv"), p], capture_output=True, text=True, timeout=120,
                         env=dict(os.environ, G4S_MV_DUMP=dump))
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("dsymv", "dtrmv", "sspmv", "dgemv"):
        assert re.search(r"matrix_multiply_%s time: [0-9.]+ ms" % name, out.stdout), name
    got = np.fromfile(dump, dtype=np.float64).reshape(4, dim)
    # the reference's fill: A[(row-1)*dim + col-1] = rand() in file order (a repeated position keeps the later value)
    libc = ctypes.CDLL(None)
    libc.srand(1)
    A = np.zeros(dim * dim)
    for r, c in zip(rows, cols):
        A[(r - 1) * dim + c - 1] = float(libc.rand())
    try:
        from oracle.binding import Ref

        impl = Ref()
        impl = impl if impl.mv_available else oracle
    except Exception:
        impl = oracle
    B = np.ones(dim)
    want = []
    _, Cv = impl.dense_mv("dsymv", A, B)
    want.append(Cv)
    B, _ = impl.dense_mv("dtrmv", A, B)
    want.append(B.copy())
    _, Cv = impl.dense_mv("dspmv" if impl is oracle else "sspmv", A, B)
    want.append(Cv)
    _, Cv = impl.dense_mv("dgemv", A, B)
    want.append(Cv)
    for k, name in enumerate(("dsymv", "dtrmv", "sspmv", "dgemv")):
        scale = np.abs(A).reshape(dim, dim).sum(axis=0).max() + np.abs(A).reshape(dim, dim).sum(axis=1).max()
        tol = 1e-12 * scale * max(1.0, np.abs(want[1]).max())
        assert np.allclose(got[k], want[k], rtol=0, atol=tol), name
    assert "checksum(C) %.17g" % got[3].sum() in out.stdout or "checksum(C)" in out.stdout
