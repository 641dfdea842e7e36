"""Property tests (hypothesis) of the host-side pieces of the path: the row partitioner against the oracle's restatement of
BIN::set_rows_offset (mm/inc/BIN.h:100-122), the binary CSR cache, the sub-matrix constructor, the CitcomS node-format
converter on arbitrary small meshes, and the tile-major order helper."""
import ctypes as C

import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings
from hypothesis import strategies as st

import g4s_b200

L = g4s_b200.lib()
FAST = settings(max_examples=40, deadline=None)


@st.composite
def csr_matrices(draw, max_rows=40, max_cols=40):
    rows = draw(st.integers(1, max_rows))
    cols = draw(st.integers(1, max_cols))
    density = draw(st.floats(0.0, 0.5))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    m = sp.random(rows, cols, density=density, random_state=rng, format="csr", data_rvs=lambda n: rng.uniform(-1, 1, n))
    m.sort_indices()
    return g4s_b200.CSR(rows, cols, m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.float64))


@FAST
@given(st.lists(st.integers(0, 50), min_size=1, max_size=200), st.integers(1, 9))
def test_partition_rows_cuts_are_monotone_cover_and_balanced(work, parts):
    prefix = np.zeros(len(work) + 1, dtype=np.int64)
    np.cumsum(work, out=prefix[1:])
    cuts = np.zeros(parts + 1, dtype=np.int32)
    assert L.g4s_partition_rows_i64(prefix.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(len(work)), C.c_int(parts),
                                    cuts.ctypes.data_as(C.POINTER(C.c_int))) == 0
    assert cuts[0] == 0 and cuts[-1] == len(work) and np.all(np.diff(cuts) >= 0)
    # the rule of BIN::set_rows_offset: part p ends at the first row whose prefix reaches (p+1) * ceil(total / parts)
    avg = -(-int(prefix[-1]) // parts)
    for p in range(parts - 1):
        want = min(int(np.searchsorted(prefix, avg * (p + 1), side="left")), len(work))
        assert cuts[p + 1] == want
        if 0 < cuts[p + 1] < len(work):  # no part carries more than the average plus one row's work
            assert prefix[cuts[p + 1]] - prefix[cuts[p]] <= avg + max(work)


@FAST
@given(csr_matrices())
def test_binary_cache_round_trip(tmp_path_factory, A):
    path = tmp_path_factory.mktemp("cache") / "m.g4scsr"
    A.save_binary(path)
    B = g4s_b200.CSR.load_binary(path)
    assert (B.rows, B.cols) == (A.rows, A.cols)
    assert np.array_equal(B.rowptr, A.rowptr) and np.array_equal(B.colids, A.colids)
    assert np.array_equal(B.values.view(np.int64), A.values.view(np.int64))


@FAST
@given(csr_matrices(), st.data())
def test_submatrix_is_the_leading_block(A, data):
    m = data.draw(st.integers(1, A.rows))
    n = data.draw(st.integers(1, A.cols))
    S = A.submatrix(m, n)
    dense = sp.csr_matrix((A.values, A.colids, A.rowptr), shape=(A.rows, A.cols)).toarray()[:m, :n]
    got = sp.csr_matrix((S.values, S.colids, S.rowptr), shape=(m, n)).toarray()
    assert (S.rows, S.cols) == (m, n) and np.array_equal(got, dense)


@settings(max_examples=15, deadline=None)
@given(st.integers(2, 5), st.integers(2, 5), st.integers(2, 5), st.integers(0, 10 ** 6))
def test_citcoms_converter_on_arbitrary_meshes(oracle, nox, noy, noz, seed):
    rng = np.random.default_rng(seed)
    nel = (nox - 1) * (noy - 1) * (noz - 1)
    half = rng.uniform(-1, 1, (nel, 24, 24))
    _, node_map, k1, k2, k3 = oracle.citcoms_mesh(nox, noy, noz, half + half.transpose(0, 2, 1))
    nno = nox * noy * noz
    rp, ci, blocks = g4s_b200.bsr_from_citcoms_nodes(node_map, k1, k2, k3)
    K = sp.bsr_matrix((blocks, ci, rp), shape=(3 * nno, 3 * nno)).tocsr()
    assert abs(K - K.T).max() == 0.0
    u = rng.uniform(-1, 1, 3 * nno)
    want = oracle.citcoms_n_assemble_del2_u(node_map, k1, k2, k3, u)
    scale = abs(K) @ np.abs(u)
    assert np.all(np.abs(K @ u - want) <= 1e-12 * scale + 1e-300)
    for b in range(nno):  # rows sorted, diagonal present
        cols = ci[rp[b]:rp[b + 1]]
        assert np.all(np.diff(cols) > 0) and b in cols


@FAST
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 9), st.integers(1, 5), st.integers(1, 5))
def test_grid_pencil_order_is_a_permutation_with_consistent_tiles(n0, n1, n2, p0, p1):
    n = n0 * n1 * n2
    order = np.empty(n, dtype=np.int32)
    nt_max = -(-n0 // p0) * -(-n1 // p1)
    tiles = np.empty(nt_max + 1, dtype=np.int32)
    nt = C.c_int()
    assert L.g4s_grid_pencil_order(C.c_int(n0), C.c_int(n1), C.c_int(n2), C.c_int(p0), C.c_int(p1),
                                   order.ctypes.data_as(C.c_void_p), tiles.ctypes.data_as(C.c_void_p), C.byref(nt)) == 0
    assert nt.value == nt_max and tiles[0] == 0 and tiles[nt_max] == n and np.all(np.diff(tiles) > 0)
    assert np.array_equal(np.sort(order), np.arange(n))
    for t in range(nt_max):  # a tile is one patch of the first two axes, swept along the third
        nodes = order[tiles[t]:tiles[t + 1]]
        i, j, k = nodes % n0, (nodes // n0) % n1, nodes // (n0 * n1)
        assert i.max() - i.min() < p0 and j.max() - j.min() < p1 and np.all(np.diff(k) >= 0)


def test_matrix_market_reader_token_stream_semantics(oracle, tmp_path):
    """The reference reads the body with `in >> i >> j >> v`: a token stream in which line breaks mean nothing.  Randomly
    formatted files (entries spanning lines, several per line, tabs / CRLF, '+' signs, exponent notations, duplicates, every
    banner kind) must give the oracle's CSR bit for bit through the parallel parser."""
    rng = np.random.default_rng(0)
    kinds = ["real general", "real symmetric", "pattern general", "integer general", "complex general",
             "real skew-symmetric", "pattern symmetric"]
    seps = [" ", "\n", "\t", "  \n", "\r\n", " \t "]
    checked = 0
    for trial in range(84):
        kind = kinds[trial % len(kinds)]
        sym = "general" not in kind
        R = int(rng.integers(1, 30))
        Cc = R if sym else int(rng.integers(1, 30))
        entries = []
        for _ in range(int(rng.integers(1, 60))):
            i, j = int(rng.integers(1, R + 1)), int(rng.integers(1, Cc + 1))
            if sym and j > i:
                i, j = j, i
            if "skew" in kind and i == j:
                continue
            ent = [("+%d" % i) if rng.random() < 0.1 else str(i), str(j)]
            if "pattern" not in kind:
                v = rng.uniform(-5, 5)
                fmt = ["%r", "%.3e", "%.17g", "%+.5f"][int(rng.integers(0, 4))]
                ent.append(str(int(v * 10)) if "integer" in kind else fmt % v)
            if "complex" in kind:
                ent.append("%.2f" % rng.uniform(-1, 1))
            entries.append(ent)
        if not entries:
            continue
        body = "".join(tok + seps[int(rng.integers(0, len(seps)))] for ent in entries for tok in ent)
        p = str(tmp_path / "fuzz.mtx")
        with open(p, "w", newline="") as fh:
            fh.write("%%MatrixMarket matrix coordinate " + kind + "\n%c\n" + "%d %d %d\n" % (R, Cc, len(entries)) + body)
        want = oracle.mm_construct(p)
        got = g4s_b200.CSR.construct(p)
        assert (got.rows, got.cols) == want[:2]
        assert np.array_equal(got.rowptr, want[2]) and np.array_equal(got.colids, want[3])
        assert np.array_equal(got.values.view(np.int64), np.asarray(want[4]).view(np.int64))
        checked += 1
    assert checked > 70
    # one token short: the count error of the reference, not a crash
    with open(p, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real general\n3 3 2\n1 1 1.0\n2 2\n")
    try:
        g4s_b200.CSR.construct(p)
        raise AssertionError("accepted a truncated file")
    except g4s_b200.G4SError as e:
        assert e.status == -5 and "read nnz not equal to declared nnz 1" in str(e)


@FAST
@given(st.integers(1, 7), st.integers(1, 7), st.integers(1, 9), st.data())
def test_grid_pencil_order_local_covers_a_rank_slab(n0, n1, n2, data):
    """The schedule of one rank of a row-partitioned mesh matrix: a permutation of its LOCAL rows, whatever the cut."""
    from g4s_b200.dist import grid_pencil_order_local

    n = n0 * n1 * n2
    row0 = data.draw(st.integers(0, n - 1))
    row1 = data.draw(st.integers(row0 + 1, n))
    order = grid_pencil_order_local(n0, n1, n2, row0, row1, p0=3, p1=2)
    assert order.dtype == np.int32 and np.array_equal(np.sort(order), np.arange(row1 - row0))


@FAST
@given(st.integers(1, 40), st.integers(0, 400), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_edge_list_constructor_equals_oracle(oracle, n, m, seed, grouped):
    """CSR(graph&) (mm/inc/CSR.h:255-329): runs of equal start vertex are sorted by (end, weight) and duplicates summed left
    to right; a start vertex that comes back later opens a new, unmerged run.  The parallel constructor must give the oracle's
    arrays bit for bit, grouped input or not."""
    rng = np.random.default_rng(seed)
    start = rng.integers(0, n, m).astype(np.int64)
    if grouped:
        start = np.sort(start)
    end = rng.integers(0, n, m).astype(np.int64)
    w = rng.integers(-4, 5, m).astype(np.float64) * 0.25 + rng.uniform(0, 1e-3, m)
    got = g4s_b200.CSR.from_graph(n, start, end, w)
    want = oracle.csr_from_graph(n, start, end, w)[-3:]
    assert np.array_equal(got.rowptr, want[0]) and np.array_equal(got.colids, want[1])
    assert np.array_equal(got.values.view(np.int64), np.asarray(want[2]).view(np.int64))
