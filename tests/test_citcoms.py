"""CitcomS node-format operator -> BSR 3x3.  The oracle restates construct_ien / construct_node_maps
/ construct_node_ks / n_assemble_del2_u (citcoms/lib/Construct_arrays.c:45-81, :264-310, :330-456;
Element_calculations.c:516-565).  PARITY UNPINNED (CitcomS is not buildable here): the restatement is cross-checked
against an independently assembled sparse matrix, then the converter and the GPU BSR SpMM are checked against it."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp


def mesh_case(oracle, nox, noy, noz, seed):
    rng = np.random.default_rng(seed)
    nel = (nox - 1) * (noy - 1) * (noz - 1)
    half = rng.uniform(-1, 1, (nel, 24, 24))
    elt_k = half + half.transpose(0, 2, 1)  # element matrices are symmetric, as get_elt_k's are
    ien, node_map, k1, k2, k3 = oracle.citcoms_mesh(nox, noy, noz, elt_k)
    return elt_k, ien, node_map, k1, k2, k3


def assembled(ien, elt_k, nno):
    dofs = (3 * (ien[:, :, None] - 1) + np.arange(3)[None, None, :]).reshape(len(ien), 24)
    rows = np.repeat(dofs, 24, axis=1).reshape(-1)
    cols = np.tile(dofs, (1, 24)).reshape(-1)
    return sp.coo_matrix((elt_k.reshape(-1), (rows, cols)), shape=(3 * nno, 3 * nno)).tocsr()


def test_node_map_layout_matches_reference_description(oracle):
    nox, noy, noz = 4, 3, 5
    _, ien, node_map, _, _, _ = mesh_case(oracle, nox, noy, noz, 0)
    nno = nox * noy * noz
    m = node_map.reshape(nno, 14, 3)
    assert ien.min() == 1 and ien.max() == nno
    # element 1: node 1 and its z / x / y neighbours (z fastest, then x, then y)
    assert list(ien[0]) == [1, 1 + noz, 1 + noz + noz * nox, 1 + noz * nox, 2, 2 + noz, 2 + noz + noz * nox, 2 + noz * nox]
    assert np.array_equal(m[:, 0, :], 3 * np.arange(nno)[:, None] + np.arange(3)[None, :])
    used = (m[:, 1:, 0] != 3 * nno).sum(axis=1)
    assert used.max() == 13 and used[0] == 0          # an interior node has 13 lower neighbours, node 1 none
    interior = (1 * nox + 1) * noz + 1                # y = 1, x = 1, z = 1 (0-based)
    assert used[interior] == 13
    low = m[interior, 1:, 0] // 3
    assert np.all(np.diff(low) > 0) and low.max() < interior


def test_oracle_node_format_product_equals_assembled_operator(oracle):
    nox, noy, noz = 5, 4, 6
    elt_k, ien, node_map, k1, k2, k3 = mesh_case(oracle, nox, noy, noz, 1)
    nno = nox * noy * noz
    u = np.random.default_rng(2).uniform(-1, 1, 3 * nno)
    K = assembled(ien, elt_k, nno)
    got = oracle.citcoms_n_assemble_del2_u(node_map, k1, k2, k3, u)
    # coefficients are accumulated and stored as float (higher_precision): agreement to float rounding only
    scale = abs(K) @ np.abs(u)
    assert np.all(np.abs(got - K @ u) <= 4e-6 * scale)


@pytest.mark.parametrize("shape", [(5, 4, 6), (2, 2, 2), (3, 7, 2)])
def test_converter_reproduces_the_operator(oracle, shape):
    import g4s_b200

    nox, noy, noz = shape
    elt_k, ien, node_map, k1, k2, k3 = mesh_case(oracle, nox, noy, noz, 3)
    nno = nox * noy * noz
    rp, ci, blocks = g4s_b200.bsr_from_citcoms_nodes(node_map, k1, k2, k3)
    # structure: exactly the block pattern of the assembled matrix, rows sorted, both triangles
    pat = assembled(ien, np.ones_like(elt_k), nno)
    bpat = sp.bsr_matrix(pat, blocksize=(3, 3))
    bpat.sort_indices()
    assert np.array_equal(rp, bpat.indptr) and np.array_equal(ci, bpat.indices)
    K = sp.bsr_matrix((blocks, ci, rp), shape=(3 * nno, 3 * nno))
    assert abs(K - K.T).max() == 0.0
    # values: every coefficient is the float the reference stores, widened exactly
    assert np.array_equal(blocks, blocks.astype(np.float32).astype(np.float64))
    rng = np.random.default_rng(4)
    Bd = rng.uniform(-1, 1, (3 * nno, 5))
    got = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), 3, Bd)
    absK = sp.bsr_matrix((np.abs(blocks), ci, rp), shape=(3 * nno, 3 * nno))
    for c in range(Bd.shape[1]):
        want = oracle.citcoms_n_assemble_del2_u(node_map, k1, k2, k3, Bd[:, c])
        assert np.all(np.abs(got[:, c] - want) <= 1e-12 * (absK @ np.abs(Bd[:, c])) + 1e-300)
    # double-precision coefficient arrays are accepted too
    rp8, ci8, blocks8 = g4s_b200.bsr_from_citcoms_nodes(node_map, k1.astype(np.float64), k2.astype(np.float64),
                                                       k3.astype(np.float64))
    assert np.array_equal(rp8, rp) and np.array_equal(ci8, ci) and np.array_equal(blocks8, blocks)


def test_converter_rejects_foreign_maps(oracle):
    import g4s_b200

    _, _, node_map, k1, k2, k3 = mesh_case(oracle, 3, 3, 3, 5)
    for edit in ("own", "upper", "half", "twice"):
        bad = node_map.copy().reshape(-1, 14, 3)
        if edit == "own":
            bad[4, 0, 1] += 1
        elif edit == "upper":
            bad[4, 1, :] = 3 * 9 + np.arange(3)          # a higher-numbered node in a lower slot
        elif edit == "half":
            bad[0, 1, 0] = 0                              # node 1 has no neighbours: slot 1 is unused
        else:
            bad[13, 2, :] = bad[13, 1, :]
        with pytest.raises(g4s_b200.G4SError) as e:
            g4s_b200.bsr_from_citcoms_nodes(bad.reshape(-1), k1, k2, k3)
        assert e.value.status == -5, edit


@pytest.mark.gpu
def test_bsr_spmm_on_a_citcoms_operator_matches_n_assemble_del2_u(oracle):
    import torch

    import g4s_b200
    from g4s_b200._lib import check

    nox, noy, noz = 9, 6, 7
    elt_k, ien, node_map, k1, k2, k3 = mesh_case(oracle, nox, noy, noz, 6)
    nno = nox * noy * noz
    rp, ci, blocks = g4s_b200.bsr_from_citcoms_nodes(node_map, k1, k2, k3)
    Bd = np.random.default_rng(777).uniform(-1, 1, (3 * nno, 64))
    t = [torch.from_numpy(a).cuda() for a in (rp, ci, blocks.reshape(-1), Bd.reshape(-1))]
    absK = sp.bsr_matrix((np.abs(blocks), ci, rp), shape=(3 * nno, 3 * nno))
    L = g4s_b200.lib()
    for variant in (0, 1, 2, 4):
        out = torch.full((3 * nno * 64,), -7.0, dtype=torch.float64, device="cuda")
        check(L.g4s_bsr_spmm_set_variant(C.c_int(variant)))
        check(L.g4s_bsr_spmm_device(C.c_int(nno), C.c_int(nno), C.c_int(3), C.c_void_p(t[0].data_ptr()),
                                    C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()), C.c_int(64),
                                    C.c_void_p(t[3].data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(0)))
        torch.cuda.synchronize()
        L.g4s_bsr_spmm_set_variant(C.c_int(0))
        got = out.cpu().numpy().reshape(3 * nno, 64)
        for c in (0, 17, 63):
            want = oracle.citcoms_n_assemble_del2_u(node_map, k1, k2, k3, Bd[:, c])
            assert np.all(np.abs(got[:, c] - want) <= 1e-12 * (absK @ np.abs(Bd[:, c])) + 1e-300), variant
