"""Pins the CPU oracle (oracle/liboracle.so) before anything is compared against it:
 - against the reference's own code compiled unmodified (oracle/_ref/libg4s_ref.so, libmv_ref.so),
 - against the closed forms and the 3x3 known-answer case recorded in SURVEY.md §4,
 - against the committed golden vectors in tests/golden/ (generated from the reference by make_golden.py).
CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from matrices import (laplacian_2d, laplacian_3d_27, powerlaw_csr, random_csr, to_scipy, to_tuple, tridiag3)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same_csr(got, want, exact_values=True):
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    if exact_values:
        np.testing.assert_array_equal(got[2], want[2])
    else:
        # other summation orders: rounding only (entries that cancel are compared on the row's scale)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-12, atol=1e-14 * np.abs(want[2]).max())


# ---------------------------------------------------------------- SpGEMM vs the reference ---------
CASES = {
    "lap2d_16": lambda: (laplacian_2d(16),) * 2,
    "lap2d_64": lambda: (laplacian_2d(64),) * 2,
    "lap3d_6": lambda: (laplacian_3d_27(6),) * 2,
    "rand_sq": lambda: (random_csr(300, 300, 0.03, 1),) * 2,
    "rand_rect": lambda: (random_csr(120, 200, 0.05, 2, empty_rows=True), random_csr(200, 90, 0.04, 3, empty_rows=True)),
    "rand_dense_rows": lambda: (random_csr(64, 64, 0.6, 4), random_csr(64, 64, 0.6, 5)),
    "powerlaw": lambda: (powerlaw_csr(3000, 7, max_deg=200),) * 2,
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_hash_spgemm_matches_reference_bit_exact(oracle, ref, name):
    A, B = CASES[name]()
    want = ref.hash_spgemm(A, B, variant=0)
    got = oracle.hash_spgemm(A, B)
    same_csr(got, want[:3])  # rowptr, colids AND fp64 values identical: same accumulation order
    got_mt = oracle.hash_spgemm(A, B, threads=3)
    same_csr(got_mt[:3], want[:3])
    # the reference's other variants agree on the pattern (values to rounding)
    heap = ref.heap_spgemm(A, B)
    same_csr(heap, want[:3], exact_values=False)
    vec = ref.hash_spgemm(A, B, variant=1)
    same_csr(vec[:3], want[:3], exact_values=False)
    # and so does scipy
    sc = to_scipy(A) @ to_scipy(B)
    sc.sort_indices()
    np.testing.assert_array_equal(sc.indptr, want[0])
    np.testing.assert_array_equal(sc.indices, want[1])
    np.testing.assert_allclose(sc.data, want[2], rtol=1e-12, atol=1e-14 * np.abs(want[2]).max())


def test_tridiag3_known_answer(oracle):
    A = tridiag3()
    rpt, col, val = oracle.hash_spgemm(A, A)
    dense = sp.csr_matrix((val, col, rpt), shape=(3, 3)).toarray()
    np.testing.assert_array_equal(dense, np.array([[5, -4, 1], [-4, 6, -4], [1, -4, 5]], dtype=float))
    assert len(col) == 9


@pytest.mark.parametrize("n", [64, 256])
def test_laplacian_closed_forms(oracle, n):
    """nnz(A) = 5n^2-4n, intprod = 25(n-2)^2+64(n-2)+36, nnz(A^2) = 13n^2-20n+4, sum = 4n+8."""
    A = laplacian_2d(n)
    assert len(A[3]) == 5 * n * n - 4 * n
    total, row_nz = oracle.intprod(A[2], A[3], A[2])
    assert total == 25 * (n - 2) ** 2 + 64 * (n - 2) + 36
    rpt, col, val = oracle.hash_spgemm(A, A)
    assert len(col) == 13 * n * n - 20 * n + 4
    assert {64: 51972, 256: 846852}[n] == len(col)
    assert val.sum() == 4 * n + 8


def test_bin_matches_reference(oracle, ref):
    for A, B in [(laplacian_2d(40),) * 2, (powerlaw_csr(2500, 11, max_deg=300),) * 2,
                 (random_csr(120, 200, 0.05, 2, empty_rows=True), random_csr(200, 90, 0.04, 3))]:
        for threads in (1, 3, 8):
            total_r, row_nz_r, offs_r, bid_r = ref.bin(threads, A, B)
            total, row_nz = oracle.intprod(A[2], A[3], B[2])
            assert total == total_r == ref.get_flop(A, B)
            np.testing.assert_array_equal(row_nz, row_nz_r)
            np.testing.assert_array_equal(oracle.rows_offset(row_nz, total, threads), offs_r)
            np.testing.assert_array_equal(oracle.bin_id(row_nz, B[1]), bid_r)


# ---------------------------------------------------------------- loaders vs the reference --------
def write_mtx(path, header, size_line, entries, comments=("% a comment",)):
    with open(path, "w") as f:
        f.write(header + "\n")
        for c in comments:
            f.write(c + "\n")
        f.write(size_line + "\n")
        for e in entries:
            f.write(" ".join(str(v) for v in e) + "\n")


def test_mm_construct_matches_reference(oracle, ref, tmp_path):
    rng = np.random.default_rng(5)
    # general real, unsorted entries
    m = sp.random(37, 29, density=0.1, random_state=rng, format="coo")
    perm = rng.permutation(m.nnz)
    ent = [(int(m.row[k]) + 1, int(m.col[k]) + 1, repr(float(m.data[k]))) for k in perm]
    p = str(tmp_path / "gen.mtx")
    write_mtx(p, "%%MatrixMarket matrix coordinate real general", "37 29 %d" % m.nnz, ent)
    got, want = oracle.mm_construct(p), ref.csr_construct(p)
    assert got[:2] == want[:2]
    same_csr(got[2:], want[2:])
    np.testing.assert_array_equal(to_scipy(got).toarray(), m.toarray())
    # symmetric (lower triangle stored), skew-symmetric, pattern, integer, complex
    low = sp.tril(sp.random(25, 25, density=0.2, random_state=rng), format="coo")
    ent = [(int(r) + 1, int(c) + 1, repr(float(v))) for r, c, v in zip(low.row, low.col, low.data)]
    for sym in ("symmetric", "skew-symmetric"):
        p = str(tmp_path / (sym + ".mtx"))
        write_mtx(p, "%%MatrixMarket matrix coordinate real " + sym, "25 25 %d" % low.nnz, ent)
        got, want = oracle.mm_construct(p), ref.csr_construct(p)
        same_csr(got[2:], want[2:])
    p = str(tmp_path / "pat.mtx")
    write_mtx(p, "%%MatrixMarket matrix coordinate pattern symmetric", "25 25 %d" % low.nnz, [e[:2] for e in ent])
    got, want = oracle.mm_construct(p), ref.csr_construct(p)
    same_csr(got[2:], want[2:])
    assert set(got[4]) == {1.0}
    p = str(tmp_path / "int.mtx")
    write_mtx(p, "%%MatrixMarket matrix coordinate integer general", "4 4 3", [(1, 1, 3), (4, 2, -7), (2, 3, 5)])
    same_csr(oracle.mm_construct(p)[2:], ref.csr_construct(p)[2:])
    p = str(tmp_path / "cplx.mtx")
    write_mtx(p, "%%MatrixMarket matrix coordinate complex general", "3 3 2", [(1, 2, 1.5, 9.0), (3, 1, -2.0, 4.0)])
    same_csr(oracle.mm_construct(p)[2:], ref.csr_construct(p)[2:])


@pytest.mark.parametrize("header,size,msg", [
    ("%%MatrixMarket matrix array real general", "3 3 1", "array"),
    ("%%MatrixMarket vector coordinate real general", "3 3 1", "banner"),
    ("%%MatrixMarket matrix coordinate real hermitian", "3 3 1", "hermitian"),
    ("%%MatrixMarket matrix coordinate quaternion general", "3 3 1", "data type"),
    ("%%MatrixMarket matrix coordinate real general", "3 3", "coordinate format"),
    # truncated body: the reference throws too, but only after it has set rows/nnz on a CSR whose pointers
    # are still uninitialised, so ~CSR frees garbage while unwinding (segfault) -- not run through the ref
    ("%%MatrixMarket matrix coordinate real general", "3 3 5", "nnz"),
])
def test_mm_construct_rejects_what_reference_rejects(oracle, ref, tmp_path, header, size, msg):
    p = str(tmp_path / "bad.mtx")
    write_mtx(p, header, size, [(1, 1, 1.0)])
    if msg != "nnz":
        with pytest.raises(RuntimeError):
            ref.csr_construct(p)
    with pytest.raises(RuntimeError, match=msg):
        oracle.mm_construct(p)
    with pytest.raises(RuntimeError):
        oracle.mm_construct(str(tmp_path / "missing.mtx"))


def test_csr_from_graph_matches_reference(oracle, ref):
    rng = np.random.default_rng(9)
    n, m = 50, 600
    start = np.sort(rng.integers(0, n, m))
    end = rng.integers(0, n, m)  # plenty of duplicate (start,end) pairs -> summed
    w = rng.uniform(0, 1, m)
    got, want = oracle.csr_from_graph(n, start, end, w), ref.csr_from_graph(n, start, end, w)
    same_csr(got[2:], want[2:])
    assert len(got[3]) < m  # duplicates really were merged
    dense = np.zeros((n, n))
    np.add.at(dense, (start, end), w)
    np.testing.assert_allclose(to_scipy(got).toarray(), dense, rtol=1e-13)
    # a start vertex that re-appears in a second run keeps both runs (reference behaviour)
    start2 = np.array([0, 0, 1, 0, 0]); end2 = np.array([3, 3, 2, 3, 1]); w2 = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    got, want = oracle.csr_from_graph(4, start2, end2, w2), ref.csr_from_graph(4, start2, end2, w2)
    same_csr(got[2:], want[2:])
    assert got[2].tolist() == [0, 3, 4, 4, 4] and got[3].tolist() == [3, 1, 3, 2]


def test_submatrix(oracle):
    A = random_csr(60, 70, 0.1, 21)
    sub = oracle.csr_submatrix(A, 40, 30, 5, 10)
    np.testing.assert_array_equal(to_scipy(sub).toarray(), to_scipy(A).toarray()[5:45, 10:40])
    with pytest.raises(ValueError):
        oracle.csr_submatrix(A, 61, 30)


# ---------------------------------------------------------------- mv ------------------------------
def test_dense_mv_matches_reference_call_sites(oracle, ref):
    """mv/mv.c's four entry points, linked to OpenBLAS in place of MKL (parity vs MKL itself is unpinned)."""
    if not ref.mv_available:
        pytest.skip("oracle/_ref/libmv_ref.so not built")
    rng = np.random.default_rng(3)
    for dim in (1, 7, 64, 257):
        A = rng.uniform(-1, 1, (dim, dim))
        B = rng.uniform(-1, 1, dim)
        for name, oname in (("dgemv", "dgemv"), ("dsymv", "dsymv"), ("dtrmv", "dtrmv"), ("sspmv", "dspmv")):
            Br, Cr = ref.dense_mv(name, A, B)
            Bo, Co = oracle.dense_mv(oname, A, B)
            scale = np.abs(A).sum() if name == "sspmv" else (np.abs(A) @ np.abs(B)).max() + np.abs(A.T) @ np.abs(B)
            np.testing.assert_allclose(Co, Cr, rtol=0, atol=1e-13 * np.max(scale))
            np.testing.assert_allclose(Bo, Br, rtol=0, atol=1e-13 * np.max(scale))
    # what the calls mean for the row-major-filled file matrix
    dim = 33
    A = rng.uniform(-1, 1, (dim, dim)); B = rng.uniform(-1, 1, dim)
    np.testing.assert_allclose(oracle.dense_mv("dgemv", A, B)[1], A.T @ B, atol=1e-13)
    L = np.tril(A); S = L + L.T - np.diag(np.diag(A))
    np.testing.assert_allclose(oracle.dense_mv("dsymv", A, B)[1], S @ B, atol=1e-13)
    np.testing.assert_allclose(oracle.dense_mv("dtrmv", A, B)[0], np.tril(A) @ B, atol=1e-13)


def test_spmv_csr_restatement(oracle):
    for A in (laplacian_2d(30), laplacian_3d_27(7), random_csr(200, 150, 0.05, 8, empty_rows=True),
              powerlaw_csr(2000, 5)):
        rng = np.random.default_rng(12345)
        x = rng.uniform(-1, 1, A[1])
        y = oracle.spmv_csr(A[2], A[3], A[4], x)
        np.testing.assert_array_equal(y, oracle.spmv_csr(A[2], A[3], A[4], x, omp=True))
        tol = 1e-13 * oracle.spmv_csr_abs(A[2], A[3], A[4], x) + 1e-300
        assert np.all(np.abs(y - to_scipy(A) @ x) <= tol)
    # cross-check against the dense reference semantics: CSR of M^T times x == dgemv on row-major M
    M = random_csr(40, 40, 0.2, 77)
    dense = to_scipy(M).toarray()
    x = np.random.default_rng(1).uniform(-1, 1, 40)
    Mt = to_tuple(to_scipy(M).T)
    np.testing.assert_allclose(oracle.spmv_csr(Mt[2], Mt[3], Mt[4], x), oracle.dense_mv("dgemv", dense, x)[1],
                               atol=1e-13)
    # Laplacian times ones: interior rows are exactly zero, the total is the boundary deficit
    A = laplacian_2d(50)
    y = oracle.spmv_csr(A[2], A[3], A[4], np.ones(A[1]))
    assert y.sum() == 4 * 50 and np.count_nonzero(y) == 4 * 50 - 4


def test_bsr_spmm_restatement(oracle):
    rng = np.random.default_rng(4)
    pat = sp.random(20, 20, density=0.2, random_state=rng, format="csr") + sp.identity(20, format="csr")
    pat.sort_indices()
    blocks = rng.uniform(-1, 1, (pat.nnz, 3, 3))
    bsr = sp.bsr_matrix((blocks, pat.indices, pat.indptr), shape=(60, 60))
    Bd = rng.uniform(-1, 1, (60, 8))
    got = oracle.bsr_spmm(pat.indptr, pat.indices, blocks.reshape(-1), 3, Bd)
    np.testing.assert_allclose(got, bsr @ Bd, atol=1e-13)


# ---------------------------------------------------------------- golden vectors ------------------
def test_golden_vectors(oracle):
    """Outputs of the reference's HashSpGEMM<false,true>, CSR::construct and CSR(graph&) captured by
    tests/golden/make_golden.py (needs /root/reference); they travel to boxes where the reference does not."""
    g = np.load(os.path.join(GOLDEN, "spgemm_golden.npz"))
    for name in sorted({k.split("__")[0] for k in g.files}):
        A = (int(g[name + "__am"]), int(g[name + "__ak"]), g[name + "__arpt"], g[name + "__acol"], g[name + "__aval"])
        B = (int(g[name + "__ak"]), int(g[name + "__bn"]), g[name + "__brpt"], g[name + "__bcol"], g[name + "__bval"])
        got = oracle.hash_spgemm(A, B)
        same_csr(got, (g[name + "__crpt"], g[name + "__ccol"], g[name + "__cval"]))
    f = np.load(os.path.join(GOLDEN, "formats_golden.npz"))
    got = oracle.mm_construct(os.path.join(GOLDEN, "sym_pattern.mtx"))
    same_csr(got[2:], (f["mtx_rpt"], f["mtx_col"], f["mtx_val"]))
    got = oracle.csr_from_graph(int(f["g_n"]), f["g_start"], f["g_end"], f["g_w"])
    same_csr(got[2:], (f["g_rpt"], f["g_col"], f["g_val"]))
