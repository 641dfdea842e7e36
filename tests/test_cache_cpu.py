"""CPU tests of the binary CSR cache: write/read round trip, rejection of damaged files, and
g4s_csr_read_cached as a drop-in for construct() — same CSR as the MatrixMarket parse that tests/test_oracle.py pins
against the reference's CSR::construct."""
import os
import shutil
import time

import numpy as np
import pytest

from matrices import laplacian_2d, random_csr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def g4s():
    import g4s_b200

    return g4s_b200


def same(a, b):
    return (a.rows, a.cols) == (b.rows, b.cols) and np.array_equal(a.rowptr, b.rowptr) and \
        np.array_equal(a.colids, b.colids) and np.array_equal(a.values.view(np.int64), b.values.view(np.int64))


@pytest.mark.parametrize("case", ["laplacian", "ragged", "empty"])
def test_binary_round_trip_is_bit_exact(g4s, tmp_path, case):
    if case == "laplacian":
        A = laplacian_2d(17)
    elif case == "ragged":
        A = random_csr(311, 97, 0.05, 3, empty_rows=True)
    else:
        A = (5, 7, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0))
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    path = tmp_path / "m.g4scsr"
    M.save_binary(path)
    assert os.path.getsize(path) == 64 + 4 * (M.rows + 1) + 12 * len(M.colids)
    assert same(g4s.CSR.load_binary(path), M)


def test_damaged_files_are_rejected(g4s, tmp_path):
    A = laplacian_2d(9)
    M = g4s.CSR(A[0], A[1], A[2], A[3], A[4])
    good = tmp_path / "good.g4scsr"
    M.save_binary(good)
    blob = bytearray(open(good, "rb").read())

    def expect(data, status):
        bad = tmp_path / "bad.g4scsr"
        open(bad, "wb").write(data)
        with pytest.raises(g4s.G4SError) as e:
            g4s.CSR.load_binary(bad)
        assert e.value.status == status

    expect(blob[:40], -5)                       # shorter than the header
    expect(blob[:-8], -5)                       # truncated arrays
    expect(blob + b"\0" * 4, -5)                # trailing bytes
    expect(b"NOTACSR\0" + blob[8:], -5)         # wrong magic
    flipped = bytearray(blob)
    flipped[-3] ^= 0x10                         # one bit of the last value: checksum
    expect(flipped, -5)
    rp = bytearray(blob)
    rp[64 + 4] ^= 0x01                          # rowptr[1] changed: still monotone, checksum catches it
    expect(rp, -5)
    with pytest.raises(g4s.G4SError) as e:
        g4s.CSR.load_binary(tmp_path / "missing.g4scsr")
    assert e.value.status == -4


def test_read_cached_is_construct(g4s, tmp_path):
    mtx = tmp_path / "sym_pattern.mtx"
    shutil.copy(os.path.join(GOLDEN, "sym_pattern.mtx"), mtx)
    want = g4s.CSR.construct(mtx)
    first = g4s.CSR.construct(mtx, cache=True)
    assert not first.cache_hit and same(first, want) and os.path.exists(str(mtx) + ".g4scsr")
    second = g4s.CSR.construct(mtx, cache=True)
    assert second.cache_hit and same(second, want)
    # a text file newer than its cache wins, and the cache is rewritten
    lines = open(mtx).read().splitlines()
    head = [i for i, l in enumerate(lines) if not l.startswith("%")][0]
    body = lines[head + 1:]
    r, c, n = lines[head].split()
    open(mtx, "w").write("\n".join(lines[:head] + ["%s %s %d" % (r, c, int(n) - 1)] + body[:-1]) + "\n")
    future = time.time() + 5
    os.utime(mtx, (future, future))
    third = g4s.CSR.construct(mtx, cache=True)
    assert not third.cache_hit and same(third, g4s.CSR.construct(mtx)) and not same(third, want)
    os.utime(str(mtx) + ".g4scsr", (future + 5, future + 5))
    assert g4s.CSR.construct(mtx, cache=True).cache_hit
    # a damaged cache is ignored and replaced
    open(str(mtx) + ".g4scsr", "wb").write(b"garbage")
    os.utime(str(mtx) + ".g4scsr", (future + 9, future + 9))
    again = g4s.CSR.construct(mtx, cache=True)
    assert not again.cache_hit and same(again, third)
    assert g4s.CSR.load_binary(str(mtx) + ".g4scsr").rows == third.rows
