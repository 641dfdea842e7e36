"""GPU parity tests (pytest -m gpu) for the rest of the C ABI: the reference-named dense mv entry points
(mv/mv.c:6-27), the device graph-to-CSR loader and R-MAT generator, the BSR SpMM kernels, and the multi-GPU
building blocks (column split, halo compaction, gather)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from matrices import laplacian_3d_27, powerlaw_csr, random_csr, to_scipy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4s():
    import g4s_b200

    assert g4s_b200.lib().g4s_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g4s_b200


# ---------------------------------------------------------------- dense mv (mv/mv.c:6-27) ----------------------
@pytest.mark.parametrize("dim", [1, 7, 64, 257, 1500])
def test_dense_mv_entry_points_match_oracle(g4s, oracle, dim):
    rng = np.random.default_rng(dim)
    A = rng.uniform(-1, 1, (dim, dim))
    B = rng.uniform(-1, 1, dim)
    absA, absB = np.abs(A), np.abs(B)
    tol = 1e-12 * (absA.sum(axis=0).max() + absA.sum(axis=1).max()) * max(absB.max(), 1e-300)
    for name, oname in (("dgemv", "dgemv"), ("dsymv", "dsymv"), ("dtrmv", "dtrmv"), ("sspmv", "dspmv")):
        Bo, Co = oracle.dense_mv(oname, A, B)
        Ag, Bg, Cg = A.copy().reshape(-1), B.copy(), np.full(dim, 123.0)
        getattr(g4s.mv, "matrix_multiply_" + name)(Ag, Bg, Cg, dim)
        if name == "dtrmv":  # in place on B, C untouched (mv/mv.c:12-15)
            np.testing.assert_allclose(Bg, Bo, rtol=0, atol=tol)
            np.testing.assert_array_equal(Cg, 123.0)
        else:
            np.testing.assert_allclose(Cg, Co, rtol=0, atol=tol)
            np.testing.assert_array_equal(Bg, B)
        np.testing.assert_array_equal(Ag, A.reshape(-1))


def test_dense_mv_equals_sparse_path(g4s):
    """dgemv on the row-major-filled dense buffer is y = M^T x; the CSR SpMV of M^T agrees."""
    M = random_csr(300, 300, 0.05, 9)
    dense = to_scipy(M).toarray()
    x = np.random.default_rng(2).uniform(-1, 1, 300)
    Cg = np.zeros(300)
    g4s.mv.matrix_multiply_dgemv(dense.copy().reshape(-1), x.copy(), Cg, 300)
    Mt = to_scipy(M).T.tocsr()
    Mt.sort_indices()
    y = g4s.CSR(300, 300, Mt.indptr, Mt.indices, Mt.data).spmv(x)
    np.testing.assert_allclose(Cg, y, rtol=0, atol=1e-12 * 300)


# ---------------------------------------------------------------- graph -> CSR loader, R-MAT ----------------------
def test_device_edge_list_loader_matches_oracle(g4s, oracle):
    import torch

    rng = np.random.default_rng(23)
    n, m = 1000, 40000
    start, end, w = rng.integers(0, n, m), rng.integers(0, n, m), rng.uniform(0, 1, m)
    order = np.argsort(start, kind="stable")  # the reference wants edges grouped by start vertex
    want = oracle.csr_from_graph(n, start[order], end[order], w[order])
    sd, ed, wd = (torch.from_numpy(a).cuda() for a in (start.astype(np.int64), end.astype(np.int64), w))
    h = C.c_void_p()
    g4s._lib.check(g4s.lib().g4s_csr_from_edges_device(C.c_long(m), C.c_long(n), C.c_void_p(sd.data_ptr()),
                                                      C.c_void_p(ed.data_ptr()), C.c_void_p(wd.data_ptr()),
                                                      C.byref(h), C.c_void_p(0)))
    got = g4s.CSR._from_handle(h).to_host()
    np.testing.assert_array_equal(got.rowptr, want[2])
    np.testing.assert_array_equal(got.colids, want[3])
    np.testing.assert_allclose(got.values, want[4], rtol=1e-12, atol=0)  # duplicates summed in another order
    assert got.nnz < m


def test_rmat_generator(g4s):
    import torch

    scale, ef = 12, 16
    A = g4s.CSR.rmat(scale, ef, seed=20240601).to_host()
    n = 1 << scale
    assert (A.rows, A.cols) == (n, n) and 0 < A.nnz <= ef * n
    assert np.all(np.diff(A.rowptr) >= 0) and A.rowptr[-1] == A.nnz
    for r in (0, 1, n // 2):
        seg = A.colids[A.rowptr[r]:A.rowptr[r + 1]]
        assert np.all(np.diff(seg) > 0)  # sorted, duplicates merged
    deg = np.diff(A.rowptr)
    assert deg.max() > 20 * deg.mean() and (deg == 0).sum() > n // 20  # power-law: hubs and empty rows
    # the merged weights add up to the sum of all generated weights; regeneration is reproducible
    m = ef * n
    sd, ed = (torch.empty(m, dtype=torch.int64, device="cuda") for _ in range(2))
    wd = torch.empty(m, dtype=torch.float64, device="cuda")
    g4s._lib.check(g4s.lib().g4s_rmat_edges_device(C.c_int(scale), C.c_longlong(m), C.c_ulonglong(20240601),
                                                  C.c_void_p(sd.data_ptr()), C.c_void_p(ed.data_ptr()),
                                                  C.c_void_p(wd.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    assert abs(float(wd.sum()) - A.values.sum()) <= 1e-9 * m
    assert 0.0 <= float(wd.min()) and float(wd.max()) < 1.0
    dense_count = torch.unique(sd * n + ed).numel()
    assert dense_count == A.nnz
    B = g4s.CSR.rmat(scale, ef, seed=20240601).to_host()
    assert np.array_equal(A.colids, B.colids) and np.array_equal(A.values, B.values)
    # quadrant probabilities: the top-left quadrant gets ~57 % of the edges
    frac = float(((sd < n // 2) & (ed < n // 2)).double().mean())
    assert abs(frac - 0.57) < 0.02


def test_spmv_on_rmat_matches_oracle(g4s, oracle):
    A = g4s.CSR.rmat(14, 16, seed=7).to_host()
    x = np.random.default_rng(4).uniform(-1, 1, A.cols)
    y = A.spmv(x)
    want = oracle.spmv_csr(A.rowptr, A.colids, A.values, x)
    scale = oracle.spmv_csr_abs(A.rowptr, A.colids, A.values, x)
    assert np.all(np.abs(y - want) <= 1e-12 * scale + 1e-300)


@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 131072, 1000003, (1 << 22) + 1])
def test_exclusive_scan_matches_reference_scan(g4s, n):
    """The reference's scan(in, out, N) (mm/inc/utility.h:166-209: out[0] = 0, out[k] = out[k-1] + in[k-1]; parallel above
    2^17 entries) on the device: tile boundaries (2048 entries a tile), one tile, thousands of tiles, in place, with and
    without the trailing total, twice in a row, and a smaller scan after a larger one (the scratch is kept between calls)."""
    import ctypes as C

    import torch

    L = g4s.lib()
    L.g4s_exclusive_scan_i32_device.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(n)
    for rep in range(2):
        a = rng.integers(0, 40, n).astype(np.int32)
        want = np.concatenate(([0], np.cumsum(a, dtype=np.int64)))
        d_in = torch.from_numpy(a).cuda()
        d_out = torch.full((n + 1,), -7, dtype=torch.int32, device="cuda")
        total = C.c_longlong(-1)
        g4s._lib.check(L.g4s_exclusive_scan_i32_device(d_in.data_ptr(), d_out.data_ptr(), n, 1, C.byref(total), None))
        assert total.value == int(want[-1])
        assert np.array_equal(d_out.cpu().numpy().astype(np.int64), want)
        # in place, without the total: entry n stays untouched
        buf = torch.cat([d_in, torch.full((1,), -7, dtype=torch.int32, device="cuda")])
        g4s._lib.check(L.g4s_exclusive_scan_i32_device(buf.data_ptr(), buf.data_ptr(), n, 0, None, None))
        torch.cuda.synchronize()
        got = buf.cpu().numpy().astype(np.int64)
        assert np.array_equal(got[:n], want[:n]) and got[n] == -7


# ---------------------------------------------------------------- BSR SpMM ------------------------------------------
def bsr_case(mb, density, bs, seed):
    rng = np.random.default_rng(seed)
    pat = (sp.random(mb, mb, density=density, random_state=rng, format="csr") + sp.identity(mb, format="csr")).tocsr()
    pat.sort_indices()
    blocks = rng.uniform(-1, 1, (pat.nnz, bs, bs))
    return pat.indptr.astype(np.int32), pat.indices.astype(np.int32), blocks


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_bsr3_spmm64_matches_oracle(g4s, oracle, variant):
    import torch

    mb, bs, ncol = 500, 3, 64
    rp, ci, blocks = bsr_case(mb, 0.03, bs, 5)
    Bd = np.random.default_rng(777).uniform(-1, 1, (mb * bs, ncol))
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), bs, Bd)
    absw = oracle.bsr_spmm(rp, ci, np.abs(blocks).reshape(-1), bs, np.abs(Bd))
    t = [torch.from_numpy(a).cuda() for a in (rp, ci, blocks.reshape(-1), Bd.reshape(-1))]
    out = torch.full((mb * bs * ncol,), -7.0, dtype=torch.float64, device="cuda")
    L = g4s.lib()
    g4s._lib.check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))
    g4s._lib.check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(bs), C.c_void_p(t[0].data_ptr()),
                                         C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()), C.c_int(ncol),
                                         C.c_void_p(t[3].data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    L.g4s_bsr_spmm_set_variant(C.c_int(0))
    got = out.cpu().numpy().reshape(mb * bs, ncol)
    assert np.all(np.abs(got - want) <= 1e-12 * absw + 1e-300)


@pytest.mark.parametrize("order", ["random", "pencil"])
def test_bsr3_spmm64_ordered_matches_oracle(g4s, oracle, order):
    import torch

    n0, n1, n2 = 7, 5, 6
    mb, bs, ncol = n0 * n1 * n2, 3, 64
    rp, ci, blocks = bsr_case(mb, 0.04, bs, 9)
    Bd = np.random.default_rng(778).uniform(-1, 1, (mb * bs, ncol))
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), bs, Bd)
    absw = oracle.bsr_spmm(rp, ci, np.abs(blocks).reshape(-1), bs, np.abs(Bd))
    L = g4s.lib()
    tiles, nt = None, C.c_int(0)
    if order == "random":
        perm = np.random.default_rng(3).permutation(mb).astype(np.int32)
    else:
        perm = np.empty(mb, dtype=np.int32)
        tiles = np.empty(2 * 2 + 1, dtype=np.int32)
        g4s._lib.check(L.g4s_grid_pencil_order(C.c_int(n0), C.c_int(n1), C.c_int(n2), C.c_int(4), C.c_int(4),
                                               perm.ctypes.data_as(C.c_void_p), tiles.ctypes.data_as(C.c_void_p),
                                               C.byref(nt)))
        assert np.array_equal(np.sort(perm), np.arange(mb)) and nt.value == 4 and tiles[-1] == mb
    t = [torch.from_numpy(a).cuda() for a in (rp, ci, blocks.reshape(-1), Bd.reshape(-1), perm)]
    td = torch.from_numpy(tiles).cuda() if tiles is not None else None
    out = torch.full((mb * bs * ncol,), -7.0, dtype=torch.float64, device="cuda")
    g4s._lib.check(L.g4s_bsr3_spmm64_ordered_device(C.c_int(mb), C.c_int(mb), C.c_void_p(t[0].data_ptr()),
                                                    C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()),
                                                    C.c_void_p(t[3].data_ptr()), C.c_void_p(out.data_ptr()),
                                                    C.c_void_p(t[4].data_ptr()),
                                                    C.c_void_p(td.data_ptr() if td is not None else 0), nt, C.c_void_p(0)))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(mb * bs, ncol)
    assert np.all(np.abs(got - want) <= 1e-12 * absw + 1e-300)


@pytest.mark.parametrize("bs,ncol", [(1, 5), (2, 33), (4, 64), (8, 7)])
def test_bsr_generic_shapes(g4s, oracle, bs, ncol):
    import torch

    mb = 120
    rp, ci, blocks = bsr_case(mb, 0.05, bs, bs)
    Bd = np.random.default_rng(1).uniform(-1, 1, (mb * bs, ncol))
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), bs, Bd)
    t = [torch.from_numpy(a).cuda() for a in (rp, ci, blocks.reshape(-1), Bd.reshape(-1))]
    out = torch.empty(mb * bs * ncol, dtype=torch.float64, device="cuda")
    g4s._lib.check(g4s.lib().g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(bs), C.c_void_p(t[0].data_ptr()),
                                                C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()),
                                                C.c_int(ncol), C.c_void_p(t[3].data_ptr()),
                                                C.c_void_p(out.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy().reshape(mb * bs, ncol), want, rtol=0, atol=1e-11)


# ---------------------------------------------------------------- multi-GPU building blocks on one GPU -------------
def test_split_compact_gather_reassemble(g4s, oracle):
    """y = A_diag x_own + A_off x_halo reproduces the whole product for every rank's block (emulated on one GPU)."""
    import torch

    from g4s_b200.dist import GpuOps, partition_rows

    ops = GpuOps()
    for A in (laplacian_3d_27(10), powerlaw_csr(4000, 13, max_deg=900)):
        x = np.random.default_rng(8).uniform(-1, 1, A[1])
        want = oracle.spmv_csr(A[2], A[3], A[4], x)
        scale = oracle.spmv_csr_abs(A[2], A[3], A[4], x)
        xd = torch.from_numpy(x).cuda()
        cuts = partition_rows(A[2], 3)
        for r in range(3):
            c0, c1 = cuts[r], cuts[r + 1]
            s, e = int(A[2][c0]), int(A[2][c1])
            local = g4s.CSR(c1 - c0, A[1], A[2][c0:c1 + 1] - s, A[3][s:e], A[4][s:e])
            diag, off = ops.split(local, c0, c1)
            assert diag.nnz + off.nnz == e - s and diag.cols == c1 - c0
            needed = ops.compact(off)
            nh = needed.cpu().numpy()
            assert np.all(np.diff(nh) > 0) and not np.any((nh >= c0) & (nh < c1))
            halo = torch.empty(max(len(nh), 1), dtype=torch.float64, device="cuda")
            if len(nh):
                ops.gather(halo[:len(nh)], xd, needed, None)
            y = torch.full((c1 - c0,), 9.0, dtype=torch.float64, device="cuda")
            ops.spmv(diag, xd[c0:c1].contiguous(), y, None)
            ops.spmv(off, halo, y, None, accumulate=True)
            torch.cuda.synchronize()
            err = np.abs(y.cpu().numpy() - want[c0:c1])
            assert np.all(err <= 1e-12 * scale[c0:c1] + 1e-300)


# ---------------------------------------------------------------- BASELINE configs at full size, by properties ------
def test_full_size_config3_rmat_properties(g4s):
    """BASELINE configs[2]: R-MAT scale 24, edge factor 16 (16 777 216 vertices, 2^28 generated edges, duplicates summed).
    Size-independent properties: structure invariants of CSR(graph&), A*1 = row sums (checked against a segmented sum by
    torch), linearity, and agreement of two kernel shapes."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~20 GB of device memory")
    scale = 24
    A = g4s.CSR.rmat(scale, 16, seed=20240601)
    n = 1 << scale
    assert A.rows == n and A.cols == n and 0 < A.nnz <= 1 << 28
    rp, ci, va = A.device_arrays()
    from g4s_b200.dist import _DevArray

    rowptr = torch.as_tensor(_DevArray(rp, n + 1, "<i4"), device="cuda")
    colids = torch.as_tensor(_DevArray(ci, A.nnz, "<i4"), device="cuda")
    values = torch.as_tensor(_DevArray(va, A.nnz, "<f8"), device="cuda")
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == A.nnz and bool((rowptr[1:] >= rowptr[:-1]).all())
    assert int(colids.min()) >= 0 and int(colids.max()) < n
    # strictly ascending columns inside every row (duplicates were summed): a drop is only allowed at a row start
    drop = torch.nonzero(colids[1:] <= colids[:-1]).flatten() + 1
    is_start = torch.zeros(A.nnz + 1, dtype=torch.bool, device="cuda")
    is_start[rowptr.long()] = True
    assert bool(is_start[drop].all())
    del drop, is_start
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    A.spmv_device(ones.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    want = torch.segment_reduce(values, "sum", offsets=rowptr.long())
    assert float((y - want).abs().max()) <= 1e-12 * float(want.max())          # values are U(0,1): no cancellation
    assert abs(float(y.sum()) - float(values.sum())) <= 1e-12 * float(values.sum())
    g = torch.Generator(device="cuda").manual_seed(11)
    u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    v = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    yu, yv, yw, y2 = (torch.empty_like(y) for _ in range(4))
    A.spmv_device(u.data_ptr(), yu.data_ptr())
    A.spmv_device(v.data_ptr(), yv.data_ptr())
    w = 2.0 * u + 3.0 * v
    A.spmv_device(w.data_ptr(), yw.data_ptr())
    torch.cuda.synchronize()
    bound = 2.0 * yu + 3.0 * yv                                               # = |A||w| for positive data
    assert bool(((yw - bound).abs() <= 1e-12 * bound + 1e-300).all())
    A.set_tuning(0, 6)                                                         # the long-row kernel shape on the same data
    A.spmv_device(u.data_ptr(), y2.data_ptr())
    torch.cuda.synchronize()
    A.set_tuning(0, 0)
    assert bool(((y2 - yu).abs() <= 1e-12 * yu + 1e-300).all())


def test_dev_size_config5_bsr_closed_form(g4s):
    """BASELINE configs[4] at its development size (128^3-node hex mesh, 27-block stencil, 3x3 blocks: diagonal 26 I + J,
    off-diagonal -I - 0.1 J): with B = 1 every entry of block row (i,j,k) is 29 - 1.3 (#neighbours), and every
    kernel / schedule must give the same C."""
    import torch

    from g4s_b200._lib import check
    from g4s_b200.dist import _DevArray

    free, _ = torch.cuda.mem_get_info()
    if free < 30e9:
        pytest.skip("needs ~12 GB of device memory")
    n, ncol = 128, 64
    P = g4s.CSR.laplacian3d27(n)
    rp, ci, va = P.device_arrays()
    nb, mb = P.nnz, P.rows
    vals = torch.as_tensor(_DevArray(va, nb, "<f8"), device="cuda")
    J = torch.ones(3, 3, dtype=torch.float64, device="cuda")
    I3 = torch.eye(3, dtype=torch.float64, device="cuda")
    diag = (vals > 0).double()[:, None, None]
    blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
    del diag
    B = torch.ones(mb * 3 * ncol, dtype=torch.float64, device="cuda")
    c = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    c[0] = c[-1] = 2.0
    neigh = (c[:, None, None] * c[None, :, None] * c[None, None, :]).reshape(-1) - 1.0
    want = (29.0 - 1.3 * neigh)[:, None].expand(mb, 3 * ncol).reshape(-1)
    L = g4s.lib()
    outs = []
    for variant in (1, 4, 2):
        Cd = torch.full((mb * 3 * ncol,), -7.0, dtype=torch.float64, device="cuda")
        check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))
        check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                    C.c_void_p(blocks.data_ptr()), C.c_int(ncol), C.c_void_p(B.data_ptr()),
                                    C.c_void_p(Cd.data_ptr()), C.c_void_p(0)))
        torch.cuda.synchronize()
        L.g4s_bsr_spmm_set_variant(C.c_int(0))
        assert float((Cd - want).abs().max()) <= 1e-12 * (29.0 + 1.3 * 26), variant
        outs.append(Cd)
    order = np.empty(mb, dtype=np.int32)
    check(L.g4s_grid_pencil_order(C.c_int(n), C.c_int(n), C.c_int(n), C.c_int(4), C.c_int(4),
                                  order.ctypes.data_as(C.c_void_p), None, None))
    od = torch.from_numpy(order).cuda()
    # random right-hand side: the ordered schedule is the K-packed kernel row by row, so C is bit-identical
    g = torch.Generator(device="cuda").manual_seed(777)
    B.copy_(torch.rand(mb * 3 * ncol, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    C1, C2 = torch.empty_like(B), torch.empty_like(B)
    check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                C.c_void_p(blocks.data_ptr()), C.c_int(ncol), C.c_void_p(B.data_ptr()),
                                C.c_void_p(C1.data_ptr()), C.c_void_p(0)))
    check(L.g4s_bsr3_spmm64_ordered_device(C.c_int(mb), C.c_int(mb), C.c_void_p(rp), C.c_void_p(ci),
                                           C.c_void_p(blocks.data_ptr()), C.c_void_p(B.data_ptr()),
                                           C.c_void_p(C2.data_ptr()), C.c_void_p(od.data_ptr()), C.c_void_p(0), C.c_int(0),
                                           C.c_void_p(0)))
    torch.cuda.synchronize()
    assert torch.equal(C1, C2)
    # the sliding-window sweep (inspector / executor) on the same matrix: grid lines along the third axis, 4 x 4 patches
    from g4s_b200 import bsr

    plan = bsr.BsrPlan(mb, mb, rp, ci, bsr.grid_pencil_strips(n, n, 0, n)).set_values(blocks.data_ptr())
    info = plan.info()
    assert info["slot_fill"] > 0.95, info
    plan.spmm(B.data_ptr(), C2.data_ptr())
    torch.cuda.synchronize()
    # same products in another order: |C| <= 52.3 * 3 here, each of the 81 terms rounds once
    assert float((C1 - C2).abs().max()) <= 1e-12 * 81 * 3
    B.fill_(1.0)
    plan.spmm(B.data_ptr(), C2.data_ptr())
    torch.cuda.synchronize()
    assert float((C2 - want).abs().max()) <= 1e-12 * (29.0 + 1.3 * 26)
    plan.destroy()


# ---------------------------------------------------------------- the library's own radix sort -----------------------------
@pytest.mark.parametrize("n,bits", [(1, 8), (31, 5), (4096, 12), (4097, 40), (300000, 57), (1 << 20, 64)])
def test_radix_sort_pairs_is_a_stable_sort(g4s, n, bits):
    """g4s_radix_sort_pairs_device against numpy's stable argsort: keys with many duplicates, values = input position, so any
    pair of equal keys that changed order shows (mm/inc/radix_sort.h:701-705 is the reference's counterpart; the join
    SpGEMM sums the values of equal keys left to right and needs the stability)."""
    import torch

    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, min(hi, 1 << 62), n, dtype=np.uint64, endpoint=True)
    keys[rng.integers(0, n, n // 2)] = keys[0]                      # lots of duplicates
    keys &= np.uint64(hi)
    vals = np.arange(n, dtype=np.uint64)
    k = torch.from_numpy(keys.view(np.int64)).cuda()
    v = torch.from_numpy(vals.view(np.int64)).cuda()
    kt, vt = torch.empty_like(k), torch.empty_like(v)
    g4s._lib.check(g4s.lib().g4s_radix_sort_pairs_device(C.c_void_p(k.data_ptr()), C.c_void_p(kt.data_ptr()), C.c_void_p(v.data_ptr()),
                                                        C.c_void_p(vt.data_ptr()), C.c_longlong(n), C.c_int(bits), C.c_void_p(0)))
    torch.cuda.synchronize()
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k.cpu().numpy().view(np.uint64), keys[order])
    np.testing.assert_array_equal(v.cpu().numpy().view(np.uint64), vals[order])
