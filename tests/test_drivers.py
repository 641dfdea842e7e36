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
def test_mv_driver_runs_the_four_entry_points(tmp_path):
    rng = np.random.default_rng(1)
    dim, nnz = 300, 4000
    p = str(tmp_path / "pat.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n% comment\n")
        f.write("%d %d %d\n" % (dim, dim, nnz))
        for r, c in zip(rng.integers(1, dim + 1, nnz), rng.integers(1, dim + 1, nnz)):
            f.write("%d %d\n" % (r, c))
    out = subprocess.run([os.path.join(BIN, "g4s_mv"), p], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("dsymv", "dtrmv", "sspmv", "dgemv"):
        assert re.search(r"matrix_multiply_%s time: [0-9.]+ ms" % name, out.stdout), name
    assert "checksum(C)" in out.stdout
