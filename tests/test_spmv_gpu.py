"""GPU parity tests for the CSR SpMV path (run on the B200 box: pytest -m gpu).
The CUDA result (through the C ABI) is compared with the CPU oracle on the same inputs.
Tolerance (BASELINE.json north_star: fp64 within 1e-12 relative, differences attributable to summation order):
|y_gpu - y_oracle| <= 1e-12 * sum_j |a_ij||x_j| per row."""
import numpy as np
import pytest

from matrices import laplacian_2d, laplacian_3d_27, powerlaw_csr, random_csr

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def g4s():
    import g4s_b200
    from g4s_b200 import lib

    assert lib().g4s_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g4s_b200


def check_spmv(g4s, oracle, A, x, lanes=0, variant=0):
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4]).set_tuning(lanes, variant)
    y = M.spmv(x)
    want = oracle.spmv_csr(A[2], A[3], A[4], x)
    scale = oracle.spmv_csr_abs(A[2], A[3], A[4], x)
    err = np.abs(y - want)
    bad = np.nonzero(err > RTOL * scale + 1e-300)[0]
    assert bad.size == 0, "rows %s differ: got %s want %s" % (bad[:5], y[bad[:5]], want[bad[:5]])
    return y


def test_generators_match_scipy(g4s):
    for n in (1, 2, 3, 17):
        want = laplacian_2d(n) if n > 1 else (1, 1, np.array([0, 1]), np.array([0]), np.array([4.0]))
        got = g4s.CSR.laplacian2d(n).to_host()
        assert (got.rows, got.cols) == (n * n, n * n)
        np.testing.assert_array_equal(got.rowptr, want[2])
        np.testing.assert_array_equal(got.colids, want[3])
        np.testing.assert_array_equal(got.values, want[4])
    for n in (2, 3, 9):
        want = laplacian_3d_27(n)
        got = g4s.CSR.laplacian3d27(n).to_host()
        np.testing.assert_array_equal(got.rowptr, want[2])
        np.testing.assert_array_equal(got.colids, want[3])
        np.testing.assert_array_equal(got.values, want[4])
        assert got.nnz == (3 * n - 2) ** 3 == g4s.lib().g4s_laplacian3d27_nnz(n, 0, -1)
    # a row block keeps global column ids and rebased row pointers
    n, r0, r1 = 9, 100, 517
    want = laplacian_3d_27(n)
    got = g4s.CSR.laplacian3d27(n, r0, r1).to_host()
    np.testing.assert_array_equal(got.rowptr, want[2][r0:r1 + 1] - want[2][r0])
    np.testing.assert_array_equal(got.colids, want[3][want[2][r0]:want[2][r1]])


CASES = {
    "lap2d_30": lambda: laplacian_2d(30),
    "lap2d_100": lambda: laplacian_2d(100),
    "lap3d_7": lambda: laplacian_3d_27(7),
    "lap3d_24": lambda: laplacian_3d_27(24),
    "random_empty_rows": lambda: random_csr(3000, 2500, 0.004, 8, empty_rows=True),
    "random_wide": lambda: random_csr(700, 9000, 0.02, 9),
    "powerlaw_hub": lambda: powerlaw_csr(30000, 5),            # one 20000-long row: spans ~10 tiles (carry chain)
    "mostly_empty": lambda: random_csr(50000, 300, 0.0002, 10),  # thousands of empty rows per tile
    "one_row": lambda: random_csr(1, 5000, 0.9, 11),
    "one_col": lambda: random_csr(4000, 1, 0.5, 12),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_spmv_matches_oracle(g4s, oracle, name):
    A = CASES[name]()
    rng = np.random.default_rng(12345)
    x = rng.uniform(-1, 1, A[1])
    check_spmv(g4s, oracle, A, x)
    check_spmv(g4s, oracle, A, np.ones(A[1]))  # x = 1.0 as mv/mv.c:65-67


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_spmv_every_lane_width(g4s, oracle, lanes):
    rng = np.random.default_rng(lanes)
    for A in (laplacian_3d_27(12), powerlaw_csr(20000, 3), random_csr(2000, 2000, 0.01, 4, empty_rows=True)):
        check_spmv(g4s, oracle, A, rng.uniform(-1, 1, A[1]), lanes=lanes)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15])
def test_spmv_every_kernel_variant(g4s, oracle, variant):
    rng = np.random.default_rng(variant)
    for A in (laplacian_3d_27(16), powerlaw_csr(20000, 6), laplacian_2d(70)):  # 2-D: several 32-row passes per chunk
        check_spmv(g4s, oracle, A, rng.uniform(-1, 1, A[1]), variant=variant)


def test_spmv_lanes1_is_bit_exact_on_short_rows(g4s, oracle):
    """One lane per row sums left to right like the oracle; with FMA contraction the only difference allowed
    is the fused rounding, which vanishes for integer-valued data."""
    A = laplacian_3d_27(10)
    x = np.arange(A[1], dtype=np.float64) % 7
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4]).set_tuning(1, 0)
    np.testing.assert_array_equal(M.spmv(x), oracle.spmv_csr(A[2], A[3], A[4], x))


def test_empty_and_degenerate(g4s):
    # all-zero matrix with rows
    M = g4s.CSR(5, 4, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0))
    np.testing.assert_array_equal(M.spmv(np.ones(4)), np.zeros(5))
    # one-shot entry point
    A = laplacian_2d(12)
    x = np.linspace(-1, 1, A[1])
    y = g4s.spmv_csr_f64(A[0], A[1], A[2], A[3], A[4], x)
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    np.testing.assert_array_equal(y, M.spmv(x))
    with pytest.raises(ValueError):
        M.spmv(np.ones(3))
    with pytest.raises(g4s.G4SError):
        M.set_tuning(3, 0)


def test_spmv_device_pointers_and_row_map(g4s, oracle):
    import torch

    A = random_csr(500, 400, 0.05, 31)
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, A[1])
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((A[0] * 2,), 7.0, dtype=torch.float64, device="cuda")
    rmap = torch.arange(0, 2 * A[0], 2, dtype=torch.int32, device="cuda")  # scatter rows to even slots
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        M.spmv_device(xd.data_ptr(), yd.data_ptr(), stream=s, row_map_ptr=rmap.data_ptr(), accumulate=True)
    s.synchronize()
    want = oracle.spmv_csr(A[2], A[3], A[4], x)
    got = yd.cpu().numpy()
    np.testing.assert_allclose(got[0::2], 7.0 + want, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(got[1::2], 7.0)


def test_full_size_config2_properties(g4s):
    """BASELINE config 2 (3-D 27-point, n=400: 64 000 000 rows, 1 719 374 392 nnz) through size-independent
    properties: A*1 = 27 - (#stencil points inside the grid) exactly, and linearity."""
    import torch

    n = 400
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~25 GB of device memory")
    A = g4s.CSR.laplacian3d27(n)
    assert A.rows == n ** 3 and A.nnz == 1719374392
    ones = torch.ones(A.cols, dtype=torch.float64, device="cuda")
    y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
    A.spmv_device(ones.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    c = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    c[0] = c[-1] = 2.0
    want = 27.0 - (c[:, None, None] * c[None, :, None] * c[None, None, :]).reshape(-1)
    assert torch.equal(y, want)
    assert float(y.sum()) == 27.0 * n ** 3 - A.nnz
    # linearity on random vectors (rounding only)
    g = torch.Generator(device="cuda").manual_seed(5)
    u = torch.rand(A.cols, dtype=torch.float64, device="cuda", generator=g) - 0.5
    v = torch.rand(A.cols, dtype=torch.float64, device="cuda", generator=g) - 0.5
    yu, yv, yw = (torch.empty_like(y) for _ in range(3))
    A.spmv_device(u.data_ptr(), yu.data_ptr())
    A.spmv_device(v.data_ptr(), yv.data_ptr())
    w = 2.0 * u - 3.0 * v
    A.spmv_device(w.data_ptr(), yw.data_ptr())
    torch.cuda.synchronize()
    err = (yw - (2.0 * yu - 3.0 * yv)).abs().max().item()
    assert err <= 1e-12 * 52 * 5
