"""Worker for tests/test_dist_cpu.py: runs the multi-GPU exchange plan of g4s_b200.dist under gloo on CPU.
The local kernels are replaced by an oracle-backed stand-in (TEST ONLY) so that the partitioning, the
diagonal / off-diagonal split bookkeeping, the request exchange and both exchange modes are exercised
without a GPU."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class CpuOps:
    device_type = "cpu"
    device = torch.device("cpu")

    def __init__(self, oracle):
        self.oracle = oracle

    def split(self, A, c0, c1):
        rp, ci, va = np.asarray(A.rowptr), np.asarray(A.colids), np.asarray(A.values)
        rows = len(rp) - 1
        inside = (ci >= c0) & (ci < c1)
        rowid = np.repeat(np.arange(rows), np.diff(rp))
        d_rp = np.zeros(rows + 1, dtype=np.int32)
        np.cumsum(np.bincount(rowid[inside], minlength=rows), out=d_rp[1:])
        diag = types.SimpleNamespace(rows=rows, cols=c1 - c0, rowptr=d_rp, colids=(ci[inside] - c0).astype(np.int32),
                                     values=va[inside].copy(), row_map=None)
        out_cnt = np.bincount(rowid[~inside], minlength=rows)
        keep = np.nonzero(out_cnt)[0]
        o_rp = np.zeros(len(keep) + 1, dtype=np.int32)
        np.cumsum(out_cnt[keep], out=o_rp[1:])
        off = types.SimpleNamespace(rows=len(keep), cols=A.cols, rowptr=o_rp, colids=ci[~inside].astype(np.int32).copy(),
                                    values=va[~inside].copy(), row_map=keep.astype(np.int32))
        off.nnz = len(off.colids)
        return diag, off

    def compact(self, off):
        needed, inv = np.unique(off.colids, return_inverse=True)
        off.colids[:] = inv.astype(np.int32)
        off.cols = len(needed)
        return torch.from_numpy(needed.astype(np.int32))

    def colids_view(self, A):
        return torch.from_numpy(A.colids)

    def gather(self, dst, src, idx, stream):
        dst.copy_(src[idx.long()])

    def spmv(self, A, x, y, stream, accumulate=False):
        if A.rows == 0:
            return
        part = self.oracle.spmv_csr(A.rowptr, A.colids, A.values, x.numpy())
        yv = y.numpy()
        if A.row_map is None:
            yv[:] = yv + part if accumulate else part
        else:
            yv[A.row_map] = (yv[A.row_map] if accumulate else 0.0) + part


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from matrices import laplacian_3d_27, powerlaw_csr, random_csr
    from oracle.binding import Oracle

    import g4s_b200
    from g4s_b200.dist import DistSpMV, partition_rows

    oracle = Oracle()
    ops = CpuOps(oracle)
    ok = True
    for name, A in (("lap3d", laplacian_3d_27(9)), ("powerlaw", powerlaw_csr(1500, 3, max_deg=400)),
                    ("random", random_csr(400, 400, 0.02, 5, empty_rows=True))):
        x = np.random.default_rng(7).uniform(-1, 1, A[1])
        want = oracle.spmv_csr(A[2], A[3], A[4], x)
        scale = oracle.spmv_csr_abs(A[2], A[3], A[4], x)
        for mode in ("halo", "allgather", "auto"):
            op = DistSpMV.from_global(g4s_b200.CSR(A[0], A[1], A[2], A[3], A[4]), mode=mode, ops=ops)
            assert op.cuts == partition_rows(A[2], world)
            xl = torch.from_numpy(x[op.c0:op.c1].copy())
            yl = torch.zeros(op.local_rows, dtype=torch.float64)
            op.apply(xl, yl)
            err = np.abs(yl.numpy() - want[op.c0:op.c1])
            good = bool(np.all(err <= 1e-12 * scale[op.c0:op.c1] + 1e-300))
            if name == "lap3d" and mode == "auto":
                good = good and op.mode == "halo" and 0 < op.n_halo <= 9 * 9 + 9 + 1  # about one plane of x from the neighbour
            if not good:
                print("rank %d: %s/%s mismatch, max err %g" % (rank, name, mode, err.max()), flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
