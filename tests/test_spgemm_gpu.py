"""GPU parity tests for SpGEMM C = A B (run on the B200 box: pytest -m gpu), through the C ABI.
Bar (BASELINE.json north_star): row pointers, column indices and sparsity pattern bit-exact against the
reference's HashSpGEMM<false,true> (via the pinned oracle and the committed golden vectors); fp64 values
within 1e-12 relative, differences attributable to summation order: |c_gpu - c_ref| <= 1e-12 * (|A||B|)_ij."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from matrices import laplacian_2d, laplacian_3d_27, powerlaw_csr, random_csr, to_scipy, to_tuple, tridiag3

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12


@pytest.fixture(scope="module")
def g4s():
    import g4s_b200

    assert g4s_b200.lib().g4s_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g4s_b200


def as_csr(g4s, t):
    return g4s.CSR(t[0], t[1], t[2], t[3], t[4])


def check_against(got, want_rpt, want_col, want_val, scale):
    np.testing.assert_array_equal(got.rowptr, want_rpt)
    np.testing.assert_array_equal(got.colids, want_col)
    err = np.abs(got.values - want_val)
    assert np.all(err <= RTOL * scale + 1e-300), "max err %g" % err.max()


def oracle_product(oracle, A, B):
    rpt, col, val = oracle.hash_spgemm(A, B)
    Aabs = (A[0], A[1], A[2], A[3], np.abs(A[4]))
    Babs = (B[0], B[1], B[2], B[3], np.abs(B[4]))
    _, _, scale = oracle.hash_spgemm(Aabs, Babs)
    return rpt, col, val, scale


def banded(rows, half_width, seed):
    """Every row has 2*half_width+1 entries: work per row of A*A up to (2h+1)^2 (classes 3-4)."""
    rng = np.random.default_rng(seed)
    diags = [rng.uniform(-1, 1, rows) for _ in range(2 * half_width + 1)]
    return to_tuple(sp.diags(diags, list(range(-half_width, half_width + 1)), shape=(rows, rows)))


def shuffled(t, seed):
    """Permute the stored order inside every row (the reference's CSR::shuffleIds idea, mm/inc/CSR.h:671-689): the
    hash algorithm does not need sorted inputs, and unsorted B rows keep the GPU off its merge fast path."""
    rng = np.random.default_rng(seed)
    rp, ci, va = t[2], t[3].copy(), t[4].copy()
    for r in range(t[0]):
        s, e = rp[r], rp[r + 1]
        perm = rng.permutation(e - s)
        ci[s:e], va[s:e] = ci[s:e][perm], va[s:e][perm]
    return (t[0], t[1], rp, ci, va)


CASES = {
    "tridiag3": lambda: (tridiag3(),) * 2,                                             # class 1 (merge)
    "lap2d_40_shuffled": lambda: (shuffled(laplacian_2d(40), 1), shuffled(laplacian_2d(40), 2)),  # class 2 (thread hash)
    "rand_rect_shuffled": lambda: (shuffled(random_csr(300, 200, 0.04, 2, empty_rows=True), 3),
                                   shuffled(random_csr(200, 150, 0.03, 3, empty_rows=True), 4)),  # classes 0, 2, 3
    "lap2d_64": lambda: (laplacian_2d(64),) * 2,                                       # class 1
    "lap3d_8": lambda: (laplacian_3d_27(8),) * 2,                                      # class 2-3 (work up to 729)
    "rand_rect": lambda: (random_csr(120, 200, 0.05, 2, empty_rows=True),
                          random_csr(200, 90, 0.04, 3, empty_rows=True)),              # class 0-2
    "rand_dense_rows": lambda: (random_csr(64, 64, 0.6, 4), random_csr(64, 64, 0.6, 5)),  # w capped by cols
    "banded_10": lambda: (banded(900, 10, 16),) * 2,                                   # work 441, <= 41 a row: class 3, 128-slot tables
    "banded_20": lambda: (banded(600, 20, 6),) * 2,                                    # work 1681, <= 81 a row: class 3, 256-slot tables
    "banded_35": lambda: (banded(5000, 35, 7),) * 2,                                   # work 5041, cols 5000: class 4
    "powerlaw_hub": lambda: (powerlaw_csr(12000, 7, max_deg=6000),) * 2,               # class 5
    "powerlaw_big_hub": lambda: (powerlaw_csr(24000, 9, max_deg=2500),) * 2,           # class 6 (dense accumulator; long B rows)
}


_ORACLE_CACHE = {}


def case_with_oracle(oracle, name):
    if name not in _ORACLE_CACHE:
        A, B = CASES[name]()
        _ORACLE_CACHE[name] = (A, B) + tuple(oracle_product(oracle, A, B))
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("name", sorted(CASES))
def test_spgemm_matches_oracle(g4s, oracle, name):
    A, B, rpt, col, val, scale = case_with_oracle(oracle, name)
    Ad, Bd = as_csr(g4s, A), as_csr(g4s, B)
    C = g4s.HashSpGEMM(Ad, Bd).to_host()
    assert (C.rows, C.cols) == (A[0], B[1])
    check_against(C, rpt, col, val, scale)
    total, _ = oracle.intprod(A[2], A[3], B[2])
    assert g4s.compute_flop(Ad, Bd) == total
    # the host-pointer mkl() twin gives the same CSR and fills the phase timings
    t = g4s.Timings()
    Ch = g4s.mkl(Ad, Bd, t)
    check_against(Ch, rpt, col, val, scale)
    assert t.total > 0 and abs(t.create + t.spmm + t.export_csr + t.destroy - t.total) < 1e-3
    # reference-style comparison (CSR::operator==, EPSILON 1e-3)
    assert C == g4s.CSR(A[0], B[1], rpt, col, val)


@pytest.mark.parametrize("name", sorted(CASES))
def test_outer_spgemm_is_bit_exact(g4s, oracle, name):
    """The expand - sort - compress path (OuterSpGEMM, mm/inc/outer_mult.h:271-542) sums every output entry in the
    reference's sequential order: pattern AND values equal HashSpGEMM<false,true>'s bit for bit, for every row class."""
    A, B, rpt, col, val, _ = case_with_oracle(oracle, name)
    C = g4s.OuterSpGEMM(as_csr(g4s, A), as_csr(g4s, B)).to_host()
    assert (C.rows, C.cols) == (A[0], B[1])
    np.testing.assert_array_equal(C.rowptr, rpt)
    np.testing.assert_array_equal(C.colids, col)
    np.testing.assert_array_equal(C.values, val)


def sorted_rows(t):
    rp, ci, va = t[2], t[3].copy(), t[4].copy()
    for r in range(t[0]):
        s, e = rp[r], rp[r + 1]
        o = np.argsort(ci[s:e], kind="stable")
        ci[s:e], va[s:e] = ci[s:e][o], va[s:e][o]
    return (t[0], t[1], rp, ci, va)


@pytest.mark.parametrize("name", sorted(CASES))
def test_heap_spgemm_is_bit_exact(g4s, oracle, name):
    """HeapSpGEMM (mm/inc/heap_mult.h:47-223), the general k-way heap merge: any number of lists per row (the power-law
    cases have rows of A with hundreds of entries), ties on a column leave the heap in A's stored order, so pattern AND
    values equal HashSpGEMM<false,true>'s bit for bit.  B's rows must be sorted: the shuffled cases are sorted first (A
    keeps its shuffled order — the merge does not care) and an unsorted B is refused."""
    A, B, _, _, _, _ = case_with_oracle(oracle, name)
    Bs = sorted_rows(B)
    rpt, col, val = oracle.hash_spgemm(A, Bs)
    C = g4s.HeapSpGEMM(as_csr(g4s, A), as_csr(g4s, Bs)).to_host()
    assert (C.rows, C.cols) == (A[0], B[1])
    np.testing.assert_array_equal(C.rowptr, rpt)
    np.testing.assert_array_equal(C.colids, col)
    np.testing.assert_array_equal(C.values, val)
    if any(np.any(np.diff(B[3][B[2][r]:B[2][r + 1]]) <= 0) for r in range(B[0])):
        with pytest.raises(g4s.G4SError):
            g4s.HeapSpGEMM(as_csr(g4s, A), as_csr(g4s, B))


def test_class6_global_hash_tables_still_match(g4s, oracle, monkeypatch):
    """Class 6 has two kernels: the dense accumulator (products with at most 2^20 columns) and hash tables in global
    memory (anything wider).  The wide case is too big for a parity test, so the hash kernel is forced on a small one."""
    monkeypatch.setenv("G4S_SPGEMM_SPA", "0")
    A, B, rpt, col, val, scale = case_with_oracle(oracle, "powerlaw_big_hub")
    check_against(g4s.HashSpGEMM(as_csr(g4s, A), as_csr(g4s, B)).to_host(), rpt, col, val, scale)


def test_tiny_row_classes_are_bit_exact(g4s, oracle):
    """Classes 1 (k-way merge) and 2 (thread-per-row table) accumulate in the reference's order with separate
    multiply and add: their values equal HashSpGEMM<false,true>'s bit for bit, not just to 1e-12."""
    for A, B in ((laplacian_2d(50),) * 2, (random_csr(500, 500, 0.006, 11), random_csr(500, 500, 0.006, 12)),
                 (shuffled(laplacian_2d(30), 5), shuffled(laplacian_2d(30), 6))):
        total, work = oracle.intprod(A[2], A[3], B[2])
        assert work.max() <= 32
        rpt, col, val = oracle.hash_spgemm(A, B)
        C = g4s.HashSpGEMM(as_csr(g4s, A), as_csr(g4s, B)).to_host()
        np.testing.assert_array_equal(C.rowptr, rpt)
        np.testing.assert_array_equal(C.colids, col)
        np.testing.assert_array_equal(C.values, val)


def test_spgemm_golden_vectors(g4s):
    """Outputs of the reference's own HashSpGEMM<false,true>, captured by tests/golden/make_golden.py."""
    g = np.load(os.path.join(GOLDEN, "spgemm_golden.npz"))
    for name in sorted({k.split("__")[0] for k in g.files}):
        A = g4s.CSR(int(g[name + "__am"]), int(g[name + "__ak"]), g[name + "__arpt"], g[name + "__acol"], g[name + "__aval"])
        B = g4s.CSR(int(g[name + "__ak"]), int(g[name + "__bn"]), g[name + "__brpt"], g[name + "__bcol"], g[name + "__bval"])
        C = g4s.HashSpGEMM(A, B).to_host()
        np.testing.assert_array_equal(C.rowptr, g[name + "__crpt"])
        np.testing.assert_array_equal(C.colids, g[name + "__ccol"])
        np.testing.assert_allclose(C.values, g[name + "__cval"], rtol=1e-12, atol=1e-13 * np.abs(g[name + "__cval"]).max())


def test_spgemm_shape_errors(g4s):
    A = as_csr(g4s, random_csr(10, 20, 0.2, 1))
    with pytest.raises(ValueError):
        g4s.HashSpGEMM(A, A)
    Z = g4s.CSR(4, 4, np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    C = g4s.HashSpGEMM(Z, Z).to_host()
    assert C.nnz == 0 and np.array_equal(C.rowptr, np.zeros(5, np.int32))


def test_full_size_config4_properties(g4s):
    """BASELINE config 4: A x A, 2-D 5-point Laplacian n = 2048 (4 194 304 rows) through the closed forms of
    SURVEY.md §4: intprod = 25(n-2)^2+64(n-2)+36, nnz(A^2) = 13n^2-20n+4 = 54 484 996, sum(values) = 4n+8,
    plus sortedness of every row and the interior stencil of A^2."""
    import torch

    n = 2048
    A = g4s.CSR.laplacian2d(n)
    assert A.nnz == 5 * n * n - 4 * n
    assert g4s.compute_flop(A, A) == 25 * (n - 2) ** 2 + 64 * (n - 2) + 36 == 104783880
    C = g4s.HashSpGEMM(A, A)
    assert C.nnz == 13 * n * n - 20 * n + 4 == 54484996
    Ch = C.to_host()
    rowptr = torch.from_numpy(Ch.rowptr.astype(np.int64))
    col = torch.from_numpy(Ch.colids.astype(np.int64))
    val = torch.from_numpy(Ch.values)
    assert float(val.sum()) == 4 * n + 8
    # sorted, duplicate-free columns inside every row
    inc = col[1:] > col[:-1]
    row_start = torch.zeros(len(col), dtype=torch.bool)
    row_start[rowptr[1:-1]] = True
    assert bool(torch.all(inc | row_start[1:]))
    # an interior row of A^2 is the 13-point biharmonic stencil: 20 at the centre, -8 x4, 2 x4, 1 x4
    r = (n // 2) * n + n // 2
    s, e = int(rowptr[r]), int(rowptr[r + 1])
    assert e - s == 13
    assert sorted(val[s:e].tolist()) == sorted([20.0] + [-8.0] * 4 + [2.0] * 4 + [1.0] * 4)
    assert col[s:e].tolist() == [r - 2 * n, r - n - 1, r - n, r - n + 1, r - 2, r - 1, r, r + 1, r + 2,
                                 r + n - 1, r + n, r + n + 1, r + 2 * n]
    # the library's second SpGEMM (expand - sort - compress) gives the identical CSR, bit for bit (the merge class keeps
    # the reference's accumulation order, and so does the join)
    Eh = g4s.OuterSpGEMM(A, A).to_host()
    assert np.array_equal(Eh.rowptr, Ch.rowptr) and np.array_equal(Eh.colids, Ch.colids)
    assert np.array_equal(Eh.values, Ch.values)


def test_repeated_product_reuses_host_decisions_and_survives_a_changed_pattern(g4s, oracle):
    """The second product of the same handles launches with the first one's class counts / nnz(C) and has the device
    confirm them (spgemm.cu: SpgemmGuess): same C.  When the operand's PATTERN is then changed in place under the same
    handle and the same nnz, the confirmation fails on the device and the product must fall back and still be exact."""
    import torch

    from g4s_b200.dist import _DevArray

    A = powerlaw_csr(3000, 4, max_deg=400)
    Bt = random_csr(3000, 2500, 0.004, 5)
    Ad, Bd = as_csr(g4s, A), as_csr(g4s, Bt)
    rpt, col, val, scale = oracle_product(oracle, A, Bt)
    for _ in range(3):
        check_against(g4s.HashSpGEMM(Ad, Bd).to_host(), rpt, col, val, scale)
    # rewrite A's column ids in place: reverse every row's columns end-for-end across the matrix width
    _, ci, _ = Ad.device_arrays()
    view = torch.as_tensor(_DevArray(ci, len(A[3]), "<i4"), device="cuda")
    new_cols = (A[1] - 1 - A[3]).astype(np.int32)
    A2 = (A[0], A[1], A[2], new_cols, A[4])           # rows now unsorted as well: a different class mix
    view.copy_(torch.from_numpy(new_cols).cuda())
    torch.cuda.synchronize()
    rpt2, col2, val2, scale2 = oracle_product(oracle, A2, Bt)
    assert not np.array_equal(rpt, rpt2)
    for _ in range(2):
        check_against(g4s.HashSpGEMM(Ad, Bd).to_host(), rpt2, col2, val2, scale2)


def test_warp_per_row_class_is_bit_exact_when_b_is_sorted(g4s, oracle):
    """Class 3 (warp per row, shared-memory table): when B's rows are strictly ascending the warp walks one row of B per
    step, so the slots of an instruction are distinct, the sums run over j ascending with one product and one addition
    per term, and the values equal HashSpGEMM<false,true>'s bit for bit — no atomicAdd(double) left on that path.
    27-point A*A rows: 729 products, 125 columns."""
    A = laplacian_3d_27(9)
    rng = np.random.default_rng(21)
    A = (A[0], A[1], A[2], A[3], rng.uniform(-1, 1, len(A[3])))      # non-trivial values: rounding order matters
    rpt, col, val = oracle.hash_spgemm(A, A)
    C = g4s.HashSpGEMM(as_csr(g4s, A), as_csr(g4s, A)).to_host()
    np.testing.assert_array_equal(C.rowptr, rpt)
    np.testing.assert_array_equal(C.colids, col)
    np.testing.assert_array_equal(C.values, val)


def short_rows(rows, cols, max_len, seed):
    """Every row holds 1..max_len strictly ascending columns: with max_len <= 8 (and products <= 64) all rows of A*B are
    in the merge class, and their outputs range from 1 to 64 entries — past the 16 the kernels stage per row."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len + 1, rows)
    rowptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(cols, size=k, replace=False)) for k in lens])
    return (rows, cols, rowptr.astype(np.int32), col.astype(np.int32), rng.uniform(-1, 1, len(col)))


MERGE_CLASS_CASES = {
    "lap2d_64": lambda: (laplacian_2d(64),) * 2,                                     # 4096 rows: full warps
    "lap2d_45": lambda: (laplacian_2d(45),) * 2,                                     # 2025 rows: ragged last CTA and warp
    "tridiag3": lambda: (tridiag3(),) * 2,                                           # one partial warp
    "short_rows_long_outputs": lambda: (short_rows(3001, 700, 8, 11), short_rows(700, 5000, 8, 12)),  # up to 64 outputs a row
    "short_rows_dense_cols": lambda: (short_rows(1500, 300, 6, 13), short_rows(300, 40, 7, 14)),      # many coinciding columns
    "tridiag_times_short": lambda: (to_tuple(sp.diags([1.5, -2.0, 0.5], [-1, 0, 1], shape=(900, 900))),
                                    short_rows(900, 1200, 8, 15)),   # contiguous stretches of B of uneven rows, overlapping
    "lap2d_times_shifted": lambda: (laplacian_2d(40), to_tuple(sp.diags([1.0, 2.0, 3.0, 4.0], [-37, -1, 2, 50], shape=(1600, 1600)))),
}


@pytest.mark.parametrize("prefetch", ["2", "1", "0"])
@pytest.mark.parametrize("name", sorted(MERGE_CLASS_CASES))
def test_merge_class_product_is_bit_exact_first_and_repeated(g4s, oracle, monkeypatch, name, prefetch):
    """Every row in the merge class (one thread per row, k-way merge of sorted rows of B, output staged compactly per warp
    and stored coalesced; spgemm.cu: spgemm_merge_row_kernel).  Row pointers, columns AND values must be those of
    HashSpGEMM<false,true> bit for bit (the merge keeps the reference's accumulation order), on the first product (host
    decisions read back) and on the repeated ones (decisions guessed, binning fused into the symbolic kernel).  The cases
    cover warps whose 32 rows' output exceeds the staging stretch (direct stores) and ragged tails.  Both numeric 5-list instances
    (with and without the next-head prefetch, G4S_SPGEMM_MERGE_PF)."""
    monkeypatch.setenv("G4S_SPGEMM_MERGE_PF", prefetch)
    A, B = MERGE_CLASS_CASES[name]()
    rpt, col, val = oracle.hash_spgemm(A, B)
    Ad, Bd = as_csr(g4s, A), as_csr(g4s, B)
    for _ in range(4):
        C = g4s.HashSpGEMM(Ad, Bd).to_host()
        np.testing.assert_array_equal(C.rowptr, rpt)
        np.testing.assert_array_equal(C.colids, col)
        np.testing.assert_array_equal(C.values, val)


def test_repeated_merge_class_product_survives_changed_operands(g4s, oracle):
    """The repeated all-merge-class product allocates C from the previous product's nnz(C) and lets the symbolic kernel do the
    binning.  When B's pattern is rewritten in place under the same handle (same nnz, other columns) the device must
    notice (nnz differs) before the numeric phase writes anything, and the product must come out exact."""
    import torch

    from g4s_b200.dist import _DevArray

    A = short_rows(2100, 600, 5, 21)
    B = short_rows(600, 900, 6, 22)
    Ad, Bd = as_csr(g4s, A), as_csr(g4s, B)
    rpt, col, val = oracle.hash_spgemm(A, B)
    for _ in range(3):
        C = g4s.HashSpGEMM(Ad, Bd).to_host()
        assert np.array_equal(C.rowptr, rpt) and np.array_equal(C.colids, col) and np.array_equal(C.values, val)
    # squeeze B's columns into a fifth of the width, keeping every row strictly ascending: far more coinciding columns
    rng = np.random.default_rng(23)
    new_cols = np.concatenate([np.sort(rng.choice(180, size=k, replace=False)) for k in np.diff(B[2])]).astype(np.int32)
    _, ci, _ = Bd.device_arrays()
    torch.as_tensor(_DevArray(ci, len(new_cols), "<i4"), device="cuda").copy_(torch.from_numpy(new_cols).cuda())
    torch.cuda.synchronize()
    B2 = (B[0], B[1], B[2], new_cols, B[4])
    rpt2, col2, val2 = oracle.hash_spgemm(A, B2)
    assert rpt2[-1] < rpt[-1]
    for _ in range(3):
        C = g4s.HashSpGEMM(Ad, Bd).to_host()
        assert np.array_equal(C.rowptr, rpt2) and np.array_equal(C.colids, col2) and np.array_equal(C.values, val2)
    # and the other way round (more entries than the previous product had: the capacity guard)
    torch.as_tensor(_DevArray(ci, len(new_cols), "<i4"), device="cuda").copy_(torch.from_numpy(B[3]).cuda())
    torch.cuda.synchronize()
    for _ in range(2):
        C = g4s.HashSpGEMM(Ad, Bd).to_host()
        assert np.array_equal(C.rowptr, rpt) and np.array_equal(C.colids, col) and np.array_equal(C.values, val)
