"""Worker for tests/test_dist_gpu.py (launched with torchrun, one rank per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from matrices import laplacian_2d, laplacian_3d_27, powerlaw_csr
    from oracle.binding import Oracle

    import g4s_b200
    from g4s_b200.dist import DistSpGEMM, DistSpMV, partition_rows

    oracle = Oracle()
    ok = True
    # ---- SpMV, both exchange modes ---------------------------------------------------------------------------
    for name, A in (("lap3d", laplacian_3d_27(14)), ("powerlaw", powerlaw_csr(6000, 3, max_deg=900))):
        x = np.random.default_rng(7).uniform(-1, 1, A[1])
        want = oracle.spmv_csr(A[2], A[3], A[4], x)
        scale = oracle.spmv_csr_abs(A[2], A[3], A[4], x)
        for mode in ("halo", "allgather", "auto", "peer"):
            op = DistSpMV.from_global(g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4]), mode=mode)
            xl = torch.from_numpy(x[op.c0:op.c1].copy()).cuda()
            yl = torch.zeros(op.local_rows, dtype=torch.float64, device="cuda")
            for it in range(3):  # repeated application reuses the plan, the streams and both peer buffers
                if mode == "peer" and it == 2:
                    op.next_x().copy_(xl)  # fill the shared buffer in place
                    op.apply(op.next_x(), yl)
                else:
                    op.apply(xl, yl)
            torch.cuda.synchronize()
            if mode == "peer":
                # the partitioned product works on a private copy of the column ids: the handle must still serve the
                # plain entry points with its global ids (the first version rewrote them in place)
                y2 = op.A.spmv(x)
                good2 = bool(np.all(np.abs(y2 - want[op.c0:op.c1]) <= 1e-12 * scale[op.c0:op.c1] + 1e-300))
                if not good2:
                    print("rank %d: %s: handle unusable after a partitioned product" % (rank, name), flush=True)
                ok = ok and good2
                op.close()
            err = np.abs(yl.cpu().numpy() - want[op.c0:op.c1])
            good = bool(np.all(err <= 1e-12 * scale[op.c0:op.c1] + 1e-300))
            if not good:
                print("rank %d: spmv %s/%s max err %g" % (rank, name, mode, err.max()), flush=True)
            ok = ok and good
    # device-generated row blocks (the bench path)
    n = 24
    op = DistSpMV.laplacian3d27(n)
    A = oracle.gen_laplacian3d27(n)
    x = np.random.default_rng(9).uniform(-1, 1, A[1])
    want = oracle.spmv_csr(A[2], A[3], A[4], x)
    xl = torch.from_numpy(x[op.c0:op.c1].copy()).cuda()
    yl = torch.zeros(op.local_rows, dtype=torch.float64, device="cuda")
    op.apply(xl, yl)
    torch.cuda.synchronize()
    good = bool(np.all(np.abs(yl.cpu().numpy() - want[op.c0:op.c1]) <= 1e-12 * 52 * 2))
    ok = ok and good and op.mode == "halo"
    op = DistSpMV.laplacian3d27(n, mode="peer")
    yl.zero_()
    op.apply(xl, yl)
    torch.cuda.synchronize()
    ok = ok and bool(np.all(np.abs(yl.cpu().numpy() - want[op.c0:op.c1]) <= 1e-12 * 52 * 2))
    op.close()
    # ---- SpGEMM: A cut by work, B broadcast -----------------------------------------------------------------------
    A = laplacian_2d(40)
    total, work = oracle.intprod(A[2], A[3], A[2])
    prefix = np.zeros(A[0] + 1, dtype=np.int64)
    np.cumsum(work, out=prefix[1:])
    cuts = partition_rows(prefix, world)
    c0, c1 = cuts[rank], cuts[rank + 1]
    s, e = int(A[2][c0]), int(A[2][c1])
    A_local = g4s_b200.CSR(c1 - c0, A[1], A[2][c0:c1 + 1] - s, A[3][s:e], A[4][s:e])
    B = g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4]) if rank == 0 else None
    mm = DistSpGEMM(A_local, B, cuts)
    C_local, offset = mm.multiply()
    C_local = C_local.to_host()
    rpt, col, val = oracle.hash_spgemm(A, A)
    good = (offset == int(rpt[c0]) and mm.global_nnz == len(col)
            and np.array_equal(C_local.rowptr + offset, rpt[c0:c1 + 1])
            and np.array_equal(C_local.colids, col[rpt[c0]:rpt[c1]])
            and np.allclose(C_local.values, val[rpt[c0]:rpt[c1]], rtol=1e-12, atol=1e-12))
    if not good:
        print("rank %d: spgemm mismatch" % rank, flush=True)
    ok = ok and good
    # ---- BSR SpMM: block rows and B partitioned, remote rows of B read over NVLink ----------------------------------
    import scipy.sparse as sp

    from g4s_b200.dist import DistBsrSpMM

    rng = np.random.default_rng(5)
    mb_all = 400
    pat = (sp.random(mb_all, mb_all, density=0.03, random_state=rng, format="csr") + sp.identity(mb_all, format="csr")).tocsr()
    pat.sort_indices()
    blocks = rng.uniform(-1, 1, (pat.nnz, 3, 3))
    Bd = rng.uniform(-1, 1, (mb_all * 3, 64))
    want = oracle.bsr_spmm(pat.indptr, pat.indices, blocks.reshape(-1), 3, Bd)
    cuts = partition_rows(pat.indptr, world)
    c0, c1 = cuts[rank], cuts[rank + 1]
    s, e = int(pat.indptr[c0]), int(pat.indptr[c1])
    op = DistBsrSpMM(torch.from_numpy((pat.indptr[c0:c1 + 1] - s).astype(np.int32)).cuda(),
                     torch.from_numpy(pat.indices[s:e].astype(np.int32)).cuda(),
                     torch.from_numpy(blocks[s:e].reshape(-1)).cuda(), cuts)
    op.B_local.copy_(torch.from_numpy(Bd[c0 * 3:c1 * 3]).cuda())
    Cl = torch.empty((c1 - c0) * 3, 64, dtype=torch.float64, device="cuda")
    import ctypes as C

    for mode in ("dfma", "kpack", "ordered"):
        Cl.fill_(-7.0)
        g4s_b200.lib().g4s_bsr_spmm_set_variant(C.c_int(4 if mode == "kpack" else 0))
        op.row_order = torch.from_numpy(np.random.default_rng(rank).permutation(c1 - c0).astype(np.int32)).cuda() \
            if mode == "ordered" else None
        op.apply(Cl)
        torch.cuda.synchronize()
        good = bool(np.allclose(Cl.cpu().numpy(), want[c0 * 3:c1 * 3], rtol=0, atol=1e-11))
        if not good:
            print("rank %d: dist bsr mismatch (%s)" % (rank, mode), flush=True)
        ok = ok and good
    g4s_b200.lib().g4s_bsr_spmm_set_variant(C.c_int(0))
    op.close()
    # ---- the same through the sliding-window sweep plan: mesh operator cut into slabs of planes, halo planes of B read
    # over NVLink by the TMA copies of the stage loader; B changes between two products (begin_update orders the write)
    from test_bsr_plan_cpu import stencil_bsr

    from g4s_b200 import bsr

    n0, n1, n2 = 6, 5, 4 * world
    rp, ci, blocks = stencil_bsr(n0, n1, n2, np.random.default_rng(11))
    plane = n0 * n1
    cuts = [4 * plane * q for q in range(world + 1)]
    c0, c1 = cuts[rank], cuts[rank + 1]
    s, e = int(rp[c0]), int(rp[c1])
    op = DistBsrSpMM(torch.from_numpy((rp[c0:c1 + 1] - s).astype(np.int32)).cuda(), torch.from_numpy(ci[s:e]).cuda(),
                     torch.from_numpy(blocks[s:e].reshape(-1)).cuda(), cuts,
                     strips=bsr.grid_pencil_strips(n0, n1, 4 * rank, 4 * rank + 4))
    Cl = torch.empty((c1 - c0) * 3, 64, dtype=torch.float64, device="cuda")
    for trial in range(3):
        Bd = np.random.default_rng(100 + trial).uniform(-1, 1, (n0 * n1 * n2 * 3, 64))
        op.begin_update().copy_(torch.from_numpy(Bd[c0 * 3:c1 * 3]).cuda())
        op.apply(Cl)
        torch.cuda.synchronize()
        want = oracle.bsr_spmm(rp, ci, blocks.reshape(-1), 3, Bd)
        good = bool(np.allclose(Cl.cpu().numpy(), want[c0 * 3:c1 * 3], rtol=0, atol=1e-11))
        if not good:
            print("rank %d: dist bsr sweep mismatch (trial %d)" % (rank, trial), flush=True)
        ok = ok and good
    op.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
