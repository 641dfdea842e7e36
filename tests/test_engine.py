"""The G4S engine ABI: spmm_dense's body on the host (callbacks) and the device form of its CitcomS
instance, the element-by-element operator, against the oracle's restatement of the reference's gather callback
(citcoms/lib/Element_calculations.c:453-471) and against an assembled sparse matrix."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp


def hex_mesh(nx, ny, nz, seed):
    """Structured hex mesh: 8 nodes per element, 3 dof per node, equation number = 3*node + direction."""
    rng = np.random.default_rng(seed)
    node = lambda i, j, k: (k * ny + j) * nx + i  # noqa: E731
    elems = []
    for k in range(nz - 1):
        for j in range(ny - 1):
            for i in range(nx - 1):
                elems.append([node(i, j, k), node(i + 1, j, k), node(i + 1, j + 1, k), node(i, j + 1, k),
                              node(i, j, k + 1), node(i + 1, j, k + 1), node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1)])
    ien = np.array(elems, dtype=np.int32)
    dofs = (3 * ien[:, :, None] + np.arange(3, dtype=np.int32)[None, None, :]).reshape(len(ien), 24)
    elt_k = rng.uniform(-1, 1, (len(ien), 24, 24))
    neq = 3 * nx * ny * nz
    u = rng.uniform(-1, 1, neq)
    return ien, dofs.astype(np.int32), elt_k, u, neq


def assembled(dofs, elt_k, neq):
    rows = np.repeat(dofs, 24, axis=1).reshape(-1)
    cols = np.tile(dofs, (1, 24)).reshape(-1)
    return sp.coo_matrix((elt_k.reshape(-1), (rows, cols)), shape=(neq, neq)).tocsr()


def test_oracle_ebe_equals_assembled_operator(oracle):
    ien, dofs, elt_k, u, neq = hex_mesh(5, 4, 6, 1)
    want = assembled(dofs, elt_k, neq) @ u
    np.testing.assert_allclose(oracle.ebe_matvec(elt_k, dofs, u, neq), want, rtol=0, atol=1e-12)


def test_spmm_dense_engine_with_citcoms_style_callbacks(oracle):
    """E->spmm_dense(nel, ends, elt_k, u, Au, Au, gather, apply, &time, 1) (Element_calculations.c:500) with a gather
    written like the reference's: vertex = element, neighbour = local node a, scatter-add of a 3 x 24 slice."""
    import g4s_b200

    L = g4s_b200.lib()
    ien, dofs, elt_k, u, neq = hex_mesh(4, 4, 4, 2)
    nel = len(ien)
    rows = [np.ascontiguousarray(elt_k[e].reshape(-1)) for e in range(nel)]
    ptrs = (C.POINTER(C.c_double) * nel)(*[r.ctypes.data_as(C.POINTER(C.c_double)) for r in rows])
    GATHER = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_double), C.POINTER(C.c_double))
    APPLY = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_double), C.POINTER(C.c_double))
    calls = {"gather": 0, "apply": 0}

    def gather(e, a, k, uu, Au):  # 0-based e, a as the engine passes them; the reference adds 1 to both inside
        calls["gather"] += 1
        n, dims, ends = 24, 3, 8
        a1 = a + 1
        for i in range(1, 4):
            aa = dofs[e][dims * a + (i - 1)]
            for b in range(1, ends + 1):
                ii = (a1 * n + b) * dims - (dims * n + dims) + (i - 1) * n
                db = dofs[e][dims * (b - 1):dims * b]
                Au[aa] += k[e][ii] * uu[db[0]] + k[e][ii + 1] * uu[db[1]] + k[e][ii + 2] * uu[db[2]]

    def apply(e, k, uu, Au):
        calls["apply"] += 1

    Au = np.zeros(neq)
    t = C.c_double(0.0)
    L.spmm_dense.restype = None
    L.spmm_dense(C.c_uint32(nel), C.c_uint32(8), ptrs, u.ctypes.data_as(C.POINTER(C.c_double)),
                 Au.ctypes.data_as(C.POINTER(C.c_double)), Au.ctypes.data_as(C.POINTER(C.c_double)), GATHER(gather),
                 APPLY(apply), C.byref(t), C.c_int(1))
    assert calls == {"gather": nel * 8, "apply": nel} and t.value > 0.0
    np.testing.assert_allclose(Au, oracle.ebe_matvec(elt_k, dofs, u, neq), rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 3, 3), (9, 7, 5), (20, 17, 12)])
def test_ebe_device_matches_oracle(oracle, shape):
    import torch

    import g4s_b200
    from g4s_b200._lib import check

    ien, dofs, elt_k, u, neq = hex_mesh(*shape, 3)
    want = oracle.ebe_matvec(elt_k, dofs, u, neq)
    scale = oracle.ebe_matvec(np.abs(elt_k), dofs, np.abs(u), neq)
    kd, dd, ud = torch.from_numpy(elt_k.reshape(-1)).cuda(), torch.from_numpy(dofs.reshape(-1)).cuda(), torch.from_numpy(u).cuda()
    Aud = torch.zeros(neq, dtype=torch.float64, device="cuda")
    check(g4s_b200.lib().g4s_ebe_matvec_device(C.c_int(len(ien)), C.c_int(24), C.c_void_p(kd.data_ptr()),
                                               C.c_void_p(dd.data_ptr()), C.c_void_p(ud.data_ptr()),
                                               C.c_void_p(Aud.data_ptr()), C.c_void_p(0)))
    torch.cuda.synchronize()
    assert np.all(np.abs(Aud.cpu().numpy() - want) <= 1e-12 * scale + 1e-300)
    with pytest.raises(g4s_b200.G4SError):
        check(g4s_b200.lib().g4s_ebe_matvec_device(C.c_int(1), C.c_int(7), C.c_void_p(kd.data_ptr()), C.c_void_p(dd.data_ptr()),
                                                   C.c_void_p(ud.data_ptr()), C.c_void_p(Aud.data_ptr()), C.c_void_p(0)))
