"""CPU-only tests of the host side of the product: the C-ABI library loads and exports every symbol that
include/g4s_b200.h declares, loaders / partitioner / Timings behave like the reference's (checked against the
pinned oracle and the golden vectors), and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from matrices import laplacian_2d, powerlaw_csr, random_csr, to_scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def g4s():
    import g4s_b200

    return g4s_b200


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g4s_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n.startswith(("g4s_", "matrix_multiply_", "compute_flop_host"))
                      and n != "g4s_alloc_fn"))


def test_library_exports_every_declared_symbol(g4s):
    L = g4s.lib()
    names = declared_symbols()
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/g4s_b200.h but not exported: %s" % missing
    assert b"sm_100a" in L.g4s_version()


def test_no_cpu_fallback(g4s):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = laplacian_2d(4)
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    with pytest.raises(g4s.G4SError) as e:
        M.spmv(np.ones(A[1]))
    assert e.value.status == -2  # G4S_ERR_CUDA
    with pytest.raises(g4s.G4SError):
        g4s.CSR.laplacian3d27(4)
    with pytest.raises(g4s.G4SError):
        g4s.mkl(M, M)


def test_timings_struct_matches_reference_layout(g4s):
    # class Timings {bool,bool,double x7} (mm/inc/Timings.h:4-22): 2 bytes + padding to 8, then 7 doubles
    assert C.sizeof(g4s.Timings) == 64 and g4s.Timings.create.offset == 8 and g4s.Timings.total.offset == 56
    L = g4s.lib()
    t, acc = g4s.Timings(), g4s.Timings()
    L.g4s_timings_init(C.byref(t))
    L.g4s_timings_init(C.byref(acc))
    assert t.measure_separate == 1 and t.measure_total == 1 and t.total == 0.0
    t.create, t.spmm, t.total = 1.0, 2.0, 4.0
    for _ in range(10):
        L.g4s_timings_add(C.byref(acc), C.byref(t))
    L.g4s_timings_div(C.byref(acc), C.c_double(10))
    assert (acc.create, acc.spmm, acc.total) == (1.0, 2.0, 4.0)


def test_matrix_market_reader(g4s, oracle, tmp_path):
    got = g4s.CSR.construct(os.path.join(GOLDEN, "sym_pattern.mtx"))
    f = np.load(os.path.join(GOLDEN, "formats_golden.npz"))  # the reference's own CSR::construct output
    np.testing.assert_array_equal(got.rowptr, f["mtx_rpt"])
    np.testing.assert_array_equal(got.colids, f["mtx_col"])
    np.testing.assert_array_equal(got.values, f["mtx_val"])
    rng = np.random.default_rng(3)
    import scipy.sparse as sp

    m = sp.random(40, 31, density=0.15, random_state=rng, format="coo")
    for header, sym in (("real general", False), ("real symmetric", True), ("real skew-symmetric", True),
                        ("integer general", False), ("complex general", False), ("pattern general", False)):
        mm = sp.tril(sp.random(30, 30, density=0.2, random_state=rng), format="coo") if sym else m
        p = str(tmp_path / (header.replace(" ", "_") + ".mtx"))
        with open(p, "w") as fh:
            fh.write("%%MatrixMarket matrix coordinate " + header + "\n% comment\n%another\n")
            fh.write("%d %d %d\n" % (mm.shape[0], mm.shape[1], mm.nnz))
            for r, c, v in zip(mm.row, mm.col, mm.data):
                val = {"real": repr(float(v)), "integer": str(int(v * 100)), "complex": "%r 0.5" % float(v),
                       "pattern": ""}[header.split()[0]]
                fh.write("%d %d %s\n" % (r + 1, c + 1, val))
        got, want = g4s.CSR.construct(p), oracle.mm_construct(p)
        assert (got.rows, got.cols) == want[:2]
        np.testing.assert_array_equal(got.rowptr, want[2])
        np.testing.assert_array_equal(got.colids, want[3])
        np.testing.assert_array_equal(got.values, want[4])


@pytest.mark.parametrize("header,size,body,status", [
    ("%%MatrixMarket matrix array real general", "3 3 1", "1 1 1.0", -5),
    ("%%MatrixMarket vector coordinate real general", "3 3 1", "1 1 1.0", -5),
    ("%%MatrixMarket matrix coordinate real hermitian", "3 3 1", "1 1 1.0", -5),
    ("%%MatrixMarket matrix coordinate quaternion general", "3 3 1", "1 1 1.0", -5),
    ("%%MatrixMarket matrix coordinate real general", "3 3", "1 1 1.0", -5),
    ("%%MatrixMarket matrix coordinate real general", "3 3 5", "1 1 1.0", -5),
    ("%%MatrixMarket matrix coordinate real general", "3 3 1", "4 1 1.0", -5),
])
def test_matrix_market_reader_rejects_malformed_input(g4s, tmp_path, header, size, body, status):
    p = str(tmp_path / "bad.mtx")
    open(p, "w").write(header + "\n" + size + "\n" + body + "\n")
    with pytest.raises(g4s.G4SError) as e:
        g4s.CSR.construct(p)
    assert e.value.status == status
    with pytest.raises(g4s.G4SError) as e:
        g4s.CSR.construct(str(tmp_path / "missing.mtx"))
    assert e.value.status == -4


def test_edge_list_constructor(g4s, oracle):
    f = np.load(os.path.join(GOLDEN, "formats_golden.npz"))  # the reference's own CSR(graph&) output
    got = g4s.CSR.from_graph(int(f["g_n"]), f["g_start"], f["g_end"], f["g_w"])
    np.testing.assert_array_equal(got.rowptr, f["g_rpt"])
    np.testing.assert_array_equal(got.colids, f["g_col"])
    np.testing.assert_array_equal(got.values, f["g_val"])
    rng = np.random.default_rng(17)
    n, m = 300, 5000
    start, end, w = np.sort(rng.integers(0, n, m)), rng.integers(0, n, m), rng.uniform(0, 1, m)
    got, want = g4s.CSR.from_graph(n, start, end, w), oracle.csr_from_graph(n, start, end, w)
    np.testing.assert_array_equal(got.rowptr, want[2])
    np.testing.assert_array_equal(got.colids, want[3])
    np.testing.assert_array_equal(got.values, want[4])
    # repeated start vertex in a second run: both runs kept (reference behaviour)
    got = g4s.CSR.from_graph(4, [0, 0, 1, 0, 0], [3, 3, 2, 3, 1], [1.0, 2.0, 3.0, 4.0, 5.0])
    assert got.rowptr.tolist() == [0, 3, 4, 4, 4] and got.colids.tolist() == [3, 1, 3, 2]
    with pytest.raises(g4s.G4SError):
        g4s.CSR.from_graph(4, [0, 5], [1, 1], [1.0, 1.0])
    empty = g4s.CSR.from_graph(3, [], [], [])
    assert empty.nnz == 0 and empty.rowptr.tolist() == [0, 0, 0, 0]


def test_submatrix_and_equality(g4s, oracle):
    A = random_csr(60, 70, 0.1, 21)
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    sub = M.submatrix(40, 30, 5, 10)
    want = oracle.csr_submatrix(A, 40, 30, 5, 10)
    np.testing.assert_array_equal(sub.rowptr, want[2])
    np.testing.assert_array_equal(sub.colids, want[3])
    np.testing.assert_array_equal(sub.values, want[4])
    np.testing.assert_array_equal(to_scipy((40, 30, sub.rowptr, sub.colids, sub.values)).toarray(),
                                  to_scipy(A).toarray()[5:45, 10:40])
    with pytest.raises(g4s.G4SError):
        M.submatrix(61, 30)
    # CSR::operator== : structure exact, values within 1e-3 abs-or-rel
    B = g4s.CSR(A[0], A[1], A[2], A[3], A[4] * (1 + 1e-5))
    assert M == B
    assert not (M == g4s.CSR(A[0], A[1], A[2], A[3], A[4] + 0.5))
    assert not (M == sub)


def test_partitioner_matches_reference_cut(g4s, oracle):
    L = g4s.lib()
    for A in (laplacian_2d(40), powerlaw_csr(2500, 11, max_deg=300)):
        total, row_nz = oracle.intprod(A[2], A[3], A[2])
        prefix = np.zeros(A[0] + 1, dtype=np.int64)
        np.cumsum(row_nz, out=prefix[1:])
        for parts in (1, 2, 3, 8):
            cuts = np.zeros(parts + 1, dtype=np.int32)
            assert L.g4s_partition_rows_i64(prefix.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(A[0]),
                                            C.c_int(parts), cuts.ctypes.data_as(C.POINTER(C.c_int))) == 0
            want = np.minimum(oracle.rows_offset(row_nz, total, parts), A[0])  # BIN::set_rows_offset, clamped
            np.testing.assert_array_equal(cuts, want)
            assert cuts[0] == 0 and cuts[-1] == A[0] and np.all(np.diff(cuts) >= 0)
        # nnz balance straight from rowptr (SpMV partition)
        cuts = np.zeros(5, dtype=np.int32)
        rp = np.ascontiguousarray(A[2], dtype=np.int32)
        assert L.g4s_partition_rows_i32(rp.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(A[0]), C.c_int(4),
                                        cuts.ctypes.data_as(C.POINTER(C.c_int))) == 0
        per = np.diff(rp[cuts])
        assert per.sum() == rp[-1] and per.max() <= rp[-1] / 4 + np.diff(rp).max() + 4
    assert L.compute_flop_host(rp.ctypes.data_as(C.POINTER(C.c_int)),
                               np.ascontiguousarray(A[3]).ctypes.data_as(C.POINTER(C.c_int)),
                               rp.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(A[0])) == total


def test_grid_pencil_order_is_a_tile_major_permutation(g4s):
    n0, n1, n2 = 6, 5, 3
    order = np.empty(n0 * n1 * n2, dtype=np.int32)
    tiles = np.empty(2 * 3 + 1, dtype=np.int32)
    nt = C.c_int()
    g4s._lib.check(g4s.lib().g4s_grid_pencil_order(C.c_int(n0), C.c_int(n1), C.c_int(n2), C.c_int(4), C.c_int(2),
                                                   order.ctypes.data_as(C.c_void_p), tiles.ctypes.data_as(C.c_void_p),
                                                   C.byref(nt)))
    assert np.array_equal(np.sort(order), np.arange(n0 * n1 * n2))
    assert nt.value == 6 and list(tiles) == [0, 24, 36, 60, 72, 84, 90]
    node = lambda i, j, k: (k * n1 + j) * n0 + i  # noqa: E731
    # first pencil: patch i in [0,4), j in [0,2), swept along k
    want = [node(i, j, k) for k in range(n2) for j in range(2) for i in range(4)]
    assert list(order[:len(want)]) == want
    # second pencil is the ragged rest of the first axis
    assert list(order[len(want):len(want) + 4]) == [node(4, 0, 0), node(5, 0, 0), node(4, 1, 0), node(5, 1, 0)]
    assert g4s.lib().g4s_grid_pencil_order(C.c_int(0), C.c_int(1), C.c_int(1), C.c_int(1), C.c_int(1),
                                           order.ctypes.data_as(C.c_void_p), None, None) == -1
