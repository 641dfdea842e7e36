"""Small test matrices built with scipy, independent of the product's own generators.
A matrix is the tuple (rows, cols, rowptr[int32], colids[int32], values[float64])."""
import numpy as np
import scipy.sparse as sp


def to_tuple(m):
    m = sp.csr_matrix(m)
    m.eliminate_zeros()  # sp.kron keeps explicit zeros of small dense factors
    m.sort_indices()
    return (m.shape[0], m.shape[1], m.indptr.astype(np.int32), m.indices.astype(np.int32),
            m.data.astype(np.float64))


def to_scipy(t):
    return sp.csr_matrix((t[4], t[3], t[2]), shape=(t[0], t[1]))


def laplacian_2d(n):
    """5-point Laplacian on an n x n grid, diag 4, off-diag -1, natural ordering."""
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    I = sp.identity(n)
    return to_tuple(sp.kron(I, T) + sp.kron(T, I))


def laplacian_3d_27(n):
    """27-point Laplacian on an n^3 grid, diag 26, off-diag -1, natural ordering (config 2)."""
    T = sp.diags([1.0, 1.0, 1.0], [-1, 0, 1], shape=(n, n))
    K = sp.kron(sp.kron(T, T), T)
    return to_tuple(27.0 * sp.identity(n ** 3) - K)


def random_csr(rows, cols, density, seed, empty_rows=False):
    rng = np.random.default_rng(seed)
    m = sp.random(rows, cols, density=density, random_state=rng, format="csr",
                  data_rvs=lambda k: rng.uniform(-1.0, 1.0, k))
    if empty_rows:
        m = m.tolil()
        for r in rng.choice(rows, size=max(1, rows // 5), replace=False):
            m.rows[r] = []
            m.data[r] = []
        m = m.tocsr()
    return to_tuple(m)


def powerlaw_csr(rows, seed, max_deg=None):
    """Rows with Zipf-distributed lengths (a few very long rows, many empty ones): merge-path territory."""
    rng = np.random.default_rng(seed)
    deg = np.minimum(rng.zipf(1.6, rows) - 1, max_deg or rows)
    deg[rng.integers(0, rows)] = min(rows, max_deg or 20000)  # one hub row
    r = np.repeat(np.arange(rows, dtype=np.int64), deg)
    c = rng.integers(0, rows, len(r))
    key = np.unique(r * rows + c)  # sorted (row, col), duplicates dropped
    r, c = key // rows, key % rows
    rowptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=rows), out=rowptr[1:])
    vals = rng.uniform(-1.0, 1.0, len(c))
    return (rows, rows, rowptr.astype(np.int32), c.astype(np.int32), vals)


def tridiag3():
    """[2 -1 0; -1 2 -1; 0 -1 2]: the 3x3 known-answer case of SURVEY.md §4."""
    return to_tuple(sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(3, 3)))
