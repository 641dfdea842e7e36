"""CPU tests of the BSR sweep inspector (g4s_bsr3_plan_inspect_host, csrc/bsr_sweep.cu): the schedule is built from host
arrays and REPLAYED here step by step in numpy exactly as the kernel walks it (stages, staged rows of B, position lists,
sub-step-major records, relative window slots, slide-and-store), and the result must be A B of the oracle's BSR product.
No device is touched: this is host logic."""
import numpy as np
import pytest
import scipy.sparse as sp

from g4s_b200 import bsr

HDR, ROWS, RUNS, REC, STRIPS = 64, 4, 20, 30, 16


def replay(mb, rp, ci, blocks, Bd, plan, world=1, cuts=None):
    stream = np.zeros(plan["stream_bytes"] // 8, dtype=np.float64)
    ints = stream.view(np.int32)
    for q in range(plan["nstages"]):
        m0, m1 = plan["meta_off"][q], plan["meta_off"][q + 1]
        o = plan["stage_table"][q, 0] // 4
        ints[o:o + (m1 - m0)] = plan["meta"][m0:m1]
    base = plan["base"]
    assert len(np.unique(base)) == len(base)
    for p in range(len(ci)):                      # sweep_set_values_kernel
        for r in range(3):
            for s in range(3):
                stream[base[p] + 10 * s + r] = blocks[p, r, s]
    C = np.full((mb * 3, 64), np.nan)
    assert plan["cta_ptr"][0] == 0 and plan["cta_ptr"][-1] == plan["nstages"]
    for t in range(plan["grid"]):
        acc = np.zeros((STRIPS, 3, 3, 64))
        for q in range(plan["cta_ptr"][t], plan["cta_ptr"][t + 1]):
            off, packed = plan["stage_table"][q]
            chunk, tx = int(packed & 0xffffffff), int(packed >> 32)
            h = ints[off // 4: off // 4 + HDR]
            nruns, S, rotate, b_off = int(h[0]), int(h[1]), int(h[2]), int(h[3])
            spad = (S + 3) & ~3
            assert chunk == HDR * 4 + STRIPS * 4 * spad + STRIPS * S * REC * 8 and b_off == (chunk + 127) & ~127
            staged = []
            pd = plan["prod"][q]                   # what warp 0 reads to start the stage's copies
            assert (int(pd[60]) & 0xffffffff) | (int(pd[61]) << 32) == off and pd[62] == chunk and pd[63] == tx
            assert not pd[2 * nruns:60].any() and nruns <= 30
            for k in range(nruns):
                j0, ln = int(h[RUNS + 2 * k]), int(h[RUNS + 2 * k + 1])
                assert pd[2 * k] == j0 and pd[2 * k + 1] == ln | (len(staged) << 8) and 0 < ln < 256
                if world > 1:  # a run never straddles two owners of B
                    assert np.searchsorted(cuts, j0, side="right") == np.searchsorted(cuts, j0 + ln - 1, side="right")
                staged += list(range(j0, j0 + ln))
            assert tx == chunk + len(staged) * 1536 and b_off + len(staged) * 1536 <= plan["stage_smem_bytes"]
            ul = ints[off // 4 + HDR: off // 4 + HDR + STRIPS * spad].reshape(STRIPS, spad) if S else None
            rec0 = (off + HDR * 4 + STRIPS * 4 * spad) // 8
            for strip in range(STRIPS):
                for st in range(S):
                    J = staged[ul[strip, st]]
                    r0 = rec0 + (strip * S + st) * REC
                    for s in range(3):
                        a9 = stream[r0 + 10 * s: r0 + 10 * s + 9].reshape(3, 3)     # [slot][block row]
                        acc[strip] += a9[:, :, None] * Bd[J * 3 + s][None, None, :]
                if rotate:
                    row = int(h[ROWS + strip])
                    if row >= 0:
                        assert np.all(np.isnan(C[row * 3:row * 3 + 3]))               # every row is stored once
                        C[row * 3:row * 3 + 3] = acc[strip, 0]
                    else:
                        assert not acc[strip, 0].any()
                    acc[strip, 0], acc[strip, 1] = acc[strip, 1].copy(), acc[strip, 2].copy()
                    acc[strip, 2] = 0.0
        assert not acc.any()
    return C


def stencil_bsr(n0, n1, n2, rng):
    idx = np.arange(n0 * n1 * n2).reshape(n2, n1, n0)
    rows, cols = [], []
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                src = idx[max(0, -dk):n2 - max(0, dk), max(0, -dj):n1 - max(0, dj), max(0, -di):n0 - max(0, di)]
                dst = idx[max(0, dk):n2 - max(0, -dk), max(0, dj):n1 - max(0, -dj), max(0, di):n0 - max(0, -di)]
                rows.append(src.ravel())
                cols.append(dst.ravel())
    pat = sp.csr_matrix((np.ones(sum(len(r) for r in rows)), (np.concatenate(rows), np.concatenate(cols))),
                        shape=(idx.size, idx.size))
    pat.sort_indices()
    return pat.indptr.astype(np.int32), pat.indices.astype(np.int32), rng.uniform(-1, 1, (pat.nnz, 3, 3))


def check(oracle, mb, rp, ci, blocks, strips, world=1, cuts=None, min_fill=0.0, grid=5, ctas=2):
    Bd = np.random.default_rng(3).uniform(-1, 1, (mb * 3, 64))
    plan = bsr.inspect_host(mb, mb, rp, ci, strips, world, cuts, grid, ctas)
    got = replay(mb, rp, ci, blocks, Bd, plan, world, cuts)
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), 3, Bd)
    absw = oracle.bsr_spmm(rp, ci, np.abs(blocks).reshape(-1), 3, np.abs(Bd))
    assert not np.isnan(got).any()
    assert np.all(np.abs(got - want) <= 1e-12 * absw + 1e-300)
    assert plan["slot_fill"] >= min_fill
    return plan


def test_mesh_operator_swept_along_grid_lines(oracle):
    rng = np.random.default_rng(1)
    n0, n1, n2 = 9, 6, 7
    rp, ci, blocks = stencil_bsr(n0, n1, n2, rng)
    strips = bsr.grid_pencil_strips(n0, n1, 0, n2)
    assert np.array_equal(np.sort(strips[1]), np.arange(n0 * n1 * n2)) and len(strips[0]) == n0 * n1 + 1
    # interior lines load every row of B once for three blocks; the mesh boundary and the partly filled last tile cost the rest
    plan = check(oracle, n0 * n1 * n2, rp, ci, blocks, strips, min_fill=0.6, grid=3, ctas=2)
    assert plan["stage_smem_bytes"] <= (227 * 1024 // 2 - 1024) // 2
    plan = check(oracle, n0 * n1 * n2, rp, ci, blocks, strips, min_fill=0.6, grid=2, ctas=1)  # whole phases per stage
    assert plan["stage_smem_bytes"] <= (227 * 1024 - 1024) // 2


def test_default_strips_on_a_banded_matrix(oracle):
    rng = np.random.default_rng(2)
    rp, ci, blocks = stencil_bsr(12, 5, 4, rng)
    check(oracle, 12 * 5 * 4, rp, ci, blocks, None)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_unstructured_pattern_with_empty_rows_and_ragged_strips(oracle, seed):
    rng = np.random.default_rng(seed)
    mb = 157
    pat = sp.random(mb, mb, density=0.04, random_state=rng, format="csr").tocsr()
    pat.sort_indices()
    blocks = rng.uniform(-1, 1, (pat.nnz, 3, 3))
    perm = rng.permutation(mb).astype(np.int32)
    cutpoints = np.unique(np.concatenate(([0, mb], rng.integers(0, mb, 11))))  # strips of length 1 .. ~40, and empty tiles avoided
    check(oracle, mb, pat.indptr.astype(np.int32), pat.indices.astype(np.int32), blocks, (cutpoints.astype(np.int32), perm))


def test_duplicate_blocks_in_a_row_are_both_applied(oracle):
    rp = np.array([0, 3, 4, 6], dtype=np.int32)
    ci = np.array([1, 1, 2, 0, 0, 0], dtype=np.int32)
    blocks = np.random.default_rng(5).uniform(-1, 1, (6, 3, 3))
    check(oracle, 3, rp, ci, blocks, None)


def test_partitioned_plan_cuts_runs_at_owner_boundaries(oracle):
    rng = np.random.default_rng(4)
    n0, n1, n2 = 5, 5, 9
    rp, ci, blocks = stencil_bsr(n0, n1, n2, rng)
    mb = n0 * n1 * n2
    # rank 1 of 3 owns planes 3..5: its rows keep GLOBAL column ids; B is cut at plane boundaries
    cuts = [0, 3 * 25, 6 * 25, mb]
    r0, r1 = cuts[1], cuts[2]
    lrp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    lci, lbl = ci[rp[r0]:rp[r1]], blocks[rp[r0]:rp[r1]]
    strips = bsr.grid_pencil_strips(n0, n1, 3, 6)
    Bd = rng.uniform(-1, 1, (mb * 3, 64))
    plan = bsr.inspect_host(r1 - r0, mb, lrp, lci, strips, 3, cuts, grid=4)
    got = replay(r1 - r0, lrp, lci, lbl, Bd, plan, 3, np.array(cuts))
    want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), 3, Bd)[r0 * 3:r1 * 3]
    assert np.allclose(got, want, rtol=0, atol=1e-12 * 27 * 3)


def test_bad_strips_are_rejected():
    from g4s_b200 import G4SError

    rp = np.array([0, 1, 2], dtype=np.int32)
    ci = np.array([0, 1], dtype=np.int32)
    with pytest.raises(G4SError):
        bsr.inspect_host(2, 2, rp, ci, (np.array([0, 2], dtype=np.int32), np.array([0, 0], dtype=np.int32)))
    with pytest.raises(G4SError):
        bsr.inspect_host(2, 2, rp, ci, (np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32)))
    with pytest.raises(G4SError):
        bsr.inspect_host(2, 1, rp, ci, None)


def test_random_patterns_and_strips_replay_exactly(oracle):
    """Property test of the inspector: random block patterns (band + noise + duplicates + empty rows), random strip
    decompositions (random order, ragged lengths), random grids and both CTAs-per-SM settings — the replayed schedule must be
    A B every time."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 70), st.integers(0, 2 ** 31 - 1), st.integers(1, 7), st.sampled_from([1, 2]), st.booleans())
    def run(mb, seed, grid, ctas, banded):
        rng = np.random.default_rng(seed)
        dens = rng.uniform(0.0, 0.2)
        pat = sp.random(mb, mb, density=dens, random_state=rng, format="csr")
        if banded:
            pat = pat + sp.diags([1.0, 1.0, 1.0], [-1, 0, 1], (mb, mb))
        pat = pat.tocsr()
        pat.sort_indices()
        rp, ci = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
        if len(ci) and rng.uniform() < 0.3:  # a duplicated block in some row
            r = int(rng.integers(0, mb))
            if rp[r + 1] > rp[r]:
                ci = np.insert(ci, rp[r], ci[rp[r]])
                rp = rp.copy()
                rp[r + 1:] += 1
        blocks = rng.uniform(-1, 1, (len(ci), 3, 3))
        perm = rng.permutation(mb).astype(np.int32)
        ncut = int(rng.integers(0, max(1, mb // 2) + 1))
        cuts = np.unique(np.concatenate(([0, mb], rng.integers(0, mb + 1, ncut)))).astype(np.int32)
        check(oracle, mb, rp, ci, blocks, (cuts, perm), grid=grid, ctas=ctas)

    run()
