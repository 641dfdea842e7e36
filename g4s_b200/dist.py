"""Row-partitioned multi-GPU SpMV / SpGEMM: one process per GPU, torch.distributed (NCCL over NVLink) for the
exchange, libg4s_b200.so for every kernel.

SpMV: rows are cut by nnz balance (BIN::set_rows_offset's cut, mm/inc/BIN.h:100-122, applied to the
nnz prefix = rowptr); rank r owns rows/x/y entries [cuts[r], cuts[r+1]).  Its row block is split once into the
DIAGONAL block (columns it owns) and the row-compressed OFF-DIAGONAL block.  Every product then runs

    comm stream : pack the x entries other ranks need  ->  all_to_all (NCCL)  ->  halo buffer
    main stream : y  = A_diag x_local                      (overlaps the exchange)
                  y += A_off  halo                         (after the exchange event)

"peer" mode (one NVSwitch box, <= 8 GPUs) removes the exchange altogether: every rank keeps its x slice in a
CUDA-IPC-shared buffer, the row block is NOT split, and ONE SpMV kernel loads the entries owned by other GPUs
straight from their memory over NVLink while it streams its own matrix (g4s_spmv_partitioned_device).  There is no
collective in the product: after writing its slice a rank stores the product's epoch into every peer's flag array
(g4s_peer_signal, a release store over NVLink), and only the chunks of the SpMV kernel that touch another GPU's
slice wait for the owners' flags; the rest of the matrix streams meanwhile.  x is double-buffered and every kernel
observes all flags before it ends, so a rank can never overwrite a slice that a slower peer is still reading.

"halo" mode moves only the x entries that are referenced (for a stencil: two planes per neighbour instead of the
whole vector); "allgather" mode is the plain NCCL all-gather of x named in BASELINE.json's north_star and is the
fallback for matrices whose off-diagonal block references most of x (R-MAT).

SpGEMM: A is cut by intermediate-product balance, B is replicated with NCCL broadcast, each rank multiplies its
row block; C stays distributed (global row pointers = local ones + an exclusive scan of per-rank nnz).

The exchange plan is backend-agnostic: `ops` supplies the local kernels (GpuOps = the CUDA library; the CPU
tests plug in an oracle-backed stand-in so the plan runs under gloo with world_size 2)."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, f64p, i32p, lib
from .csr import CSR, HashSpGEMM, _stream_ptr


# ------------------------------------------------------------------------------------------------ partitioning
def partition_by_prefix(prefix_fn, rows, parts):
    """cuts[p+1] = lower_bound(prefix, ceil(total/parts)*(p+1)) over prefix[0..rows], last cut = rows — the rule of
    BIN::set_rows_offset (and of g4s_partition_rows_*), for a prefix given as a function (closed-form generators)."""
    total = prefix_fn(rows)
    avg = (total + parts - 1) // parts
    cuts = [0]
    for p in range(parts):
        target = avg * (p + 1)
        lo, hi = 0, rows + 1
        while lo < hi:
            mid = (lo + hi) // 2
            if prefix_fn(mid) < target:
                lo = mid + 1
            else:
                hi = mid
        cuts.append(min(lo, rows))
    cuts[-1] = rows
    return cuts


def partition_rows(prefix, parts):
    """Same cut on an explicit prefix array (rowptr for nnz balance); calls the C ABI partitioner."""
    prefix = np.ascontiguousarray(prefix, dtype=np.int64)
    cuts = np.zeros(parts + 1, dtype=np.int32)
    check(lib().g4s_partition_rows_i64(prefix.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(len(prefix) - 1),
                                       C.c_int(parts), cuts.ctypes.data_as(i32p)))
    return [int(c) for c in cuts]


# ------------------------------------------------------------------------------------------------ local kernels
class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuOps:
    """Local operations on the GPU through the C ABI."""
    device_type = "cuda"

    def __init__(self):
        self.device = torch.device("cuda", torch.cuda.current_device())

    def split(self, A, c0, c1):
        d, o = C.c_void_p(), C.c_void_p()
        check(lib().g4s_csr_split_columns(A.handle, C.c_int(c0), C.c_int(c1), C.byref(d), C.byref(o), C.c_void_p(0)))
        return CSR._from_handle(d), CSR._from_handle(o)

    def compact(self, off):
        ptr, n = i32p(), C.c_int()
        check(lib().g4s_csr_compact_columns(off.handle, C.byref(ptr), C.byref(n), C.c_void_p(0)))
        addr = C.cast(ptr, C.c_void_p).value
        if n.value:
            t = torch.as_tensor(_DevArray(addr, n.value, "<i4"), device=self.device).clone()
        else:
            t = torch.empty(0, dtype=torch.int32, device=self.device)
        check(lib().g4s_device_free(C.c_void_p(addr)))
        off.cols = n.value
        return t

    def colids_view(self, A):
        """Zero-copy torch view of a handle's column ids (used to remap them for the all-gather layout)."""
        _, ci, _ = A.device_arrays()
        return torch.as_tensor(_DevArray(ci, A.nnz, "<i4"), device=self.device) if A.nnz else \
            torch.empty(0, dtype=torch.int32, device=self.device)

    def gather(self, dst, src, idx, stream):
        check(lib().g4s_gather_f64(C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), C.c_void_p(idx.data_ptr()),
                                   C.c_longlong(idx.numel()), _stream_ptr(stream)))

    def spmv(self, A, x, y, stream, accumulate=False):
        if A.rows == 0:
            return
        if accumulate:
            A.spmv_device(x.data_ptr(), y.data_ptr(), stream=stream, accumulate=True)
        else:
            A.spmv_device(x.data_ptr(), y.data_ptr(), stream=stream)


# ------------------------------------------------------------------------------------------------ SpMV
class DistSpMV:
    """y = A x with A row-partitioned over the ranks of `group`.

    A_local : CSR holding this rank's rows [cuts[rank], cuts[rank+1]) with GLOBAL column ids (cols = global n).
    cuts    : world+1 row cuts (also the ownership of x and y entries; A must be square)."""

    def __init__(self, A_local, cuts, group=None, mode="auto", ops=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.ops = ops if ops is not None else GpuOps()
        self.cuts = [int(c) for c in cuts]
        assert len(self.cuts) == self.world + 1
        self.c0, self.c1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        self.local_rows = self.c1 - self.c0
        assert A_local.rows == self.local_rows, "A_local must hold exactly this rank's rows"
        dev = self.ops.device
        self.comm_stream = None
        if mode == "peer":
            self._init_peer(A_local, dev)
            return
        self.diag, self.off = self.ops.split(A_local, self.c0, self.c1)
        # how much of x the off-diagonal block references decides the exchange
        needed = self.ops.compact(self.off)
        frac = torch.tensor([needed.numel() / max(1, self.cuts[-1] - self.local_rows)], dtype=torch.float64, device=dev)
        dist.all_reduce(frac, op=dist.ReduceOp.MAX, group=self.group)
        self.mode = mode if mode != "auto" else ("allgather" if float(frac.item()) > 0.5 else "halo")
        bounds = torch.tensor(self.cuts, dtype=needed.dtype, device=dev)
        # needed is sorted, hence grouped by owner
        edges = torch.searchsorted(needed, bounds)
        self.recv_counts = [int(v) for v in (edges[1:] - edges[:-1]).tolist()]
        self.n_halo = int(needed.numel())
        if self.mode == "halo":
            mine = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
            theirs = torch.empty_like(mine)
            dist.all_to_all_single(theirs, mine, group=self.group)
            self.send_counts = [int(v) for v in theirs.tolist()]
            req_in = torch.empty(sum(self.send_counts), dtype=needed.dtype, device=dev)
            dist.all_to_all_single(req_in, needed, self.send_counts, self.recv_counts, group=self.group)
            self.send_idx = (req_in - self.c0).to(torch.int32).contiguous()
            self.send_buf = torch.empty(self.send_idx.numel(), dtype=torch.float64, device=dev)
            self.halo = torch.empty(max(self.n_halo, 1), dtype=torch.float64, device=dev)
            self.exchange_bytes = 8 * self.n_halo
        else:
            # all-gather layout: slice q sits at q*slot; off-diagonal column ids (halo positions after compact)
            # are mapped to that layout once
            self.slot = max(self.cuts[q + 1] - self.cuts[q] for q in range(self.world))
            owner = torch.searchsorted(bounds, needed, right=True) - 1
            padded = (owner * self.slot + (needed - bounds[owner])).to(torch.int32)
            col = self.ops.colids_view(self.off)
            if col.numel():
                col.copy_(padded[col.long()])
            self.off.cols = self.slot * self.world
            self.x_pad = torch.zeros(self.slot, dtype=torch.float64, device=dev)
            self.halo = torch.empty(self.slot * self.world, dtype=torch.float64, device=dev)
            self.exchange_bytes = 8 * self.slot * (self.world - 1)
        self.comm_stream = torch.cuda.Stream() if self.ops.device_type == "cuda" else None
        self._k = 0
        self.launches_per_step = 1 + (1 if self.off.rows else 0) + (1 if self.mode == "halo" and self.send_idx.numel() else 0)

    def _init_peer(self, A_local, dev):
        if self.world > 8:
            raise ValueError("peer mode supports up to 8 GPUs (one NVSwitch box)")
        L = lib()
        self.mode = "peer"
        self.A = A_local
        self._own, mine = [], []
        for b in range(3):  # two x buffers and the flag array
            ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
            nbytes = 8 * max(self.local_rows, 1) if b < 2 else 8 * 8
            check(L.g4s_peer_alloc(C.c_size_t(nbytes), C.byref(ptr), h))
            self._own.append(ptr.value)
            mine.append(bytes(h))
        self._flags = torch.as_tensor(_DevArray(self._own[2], 8, "<i8"), device=dev)
        self._flags.zero_()
        torch.cuda.synchronize()
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=self.group)
        self._opened, self._parts = [], []
        for b in range(3):
            arr = (C.c_void_p * self.world)()
            for q in range(self.world):
                if q == self.rank:
                    arr[q] = self._own[b]
                else:
                    p = C.c_void_p()
                    check(L.g4s_peer_open((C.c_ubyte * 64).from_buffer_copy(every[q][b]), C.byref(p)))
                    arr[q] = p.value
                    self._opened.append(p.value)
            self._parts.append(arr)
        self.x_buffers = [torch.as_tensor(_DevArray(self._own[b], max(self.local_rows, 1), "<f8"), device=dev)[:self.local_rows]
                          for b in range(2)]
        self._cuts_c = (C.c_int * (self.world + 1))(*self.cuts)
        self._flag_arrays = self._parts.pop()
        self._k = 0
        self.n_halo, self.exchange_bytes, self.launches_per_step = 0, 0, 1
        dist.barrier(group=self.group)
        # set-up product on zeros: numbers this rank's remote columns (the library rewrites their ids once) and tells
        # which ranks own them.  A product then waits for those ranks only -- widened here to the ranks that read THIS
        # rank's slice, so that the double-buffering argument also holds for patterns that are not symmetric.
        for xb in self.x_buffers:
            xb.zero_()
        scratch = torch.empty(max(self.local_rows, 1), dtype=torch.float64, device=dev)
        self.apply(self.x_buffers[0], scratch)
        torch.cuda.synchronize()
        n_remote, mask = C.c_int(), C.c_uint()
        check(L.g4s_spmv_partition_info(self.A.handle, C.byref(n_remote), C.byref(mask)))
        self.n_halo, self.exchange_bytes = max(n_remote.value, 0), 8 * max(n_remote.value, 0)
        every_mask = [None] * self.world
        dist.all_gather_object(every_mask, int(mask.value), group=self.group)
        wait = int(mask.value)
        for q in range(self.world):
            if (every_mask[q] >> self.rank) & 1:
                wait |= 1 << q
        check(L.g4s_spmv_partition_set_wait_mask(self.A.handle, C.c_uint(wait)))
        dist.barrier(group=self.group)

    def next_x(self):
        """peer mode: the shared buffer the next apply() will read; fill it in place to skip apply()'s copy."""
        return self.x_buffers[self._k % 2]

    def close(self):
        if getattr(self, "mode", None) == "peer" and self._own:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for p in self._opened:
                lib().g4s_peer_close(C.c_void_p(p))
            dist.barrier(group=self.group)
            for p in self._own:
                lib().g4s_peer_free(C.c_void_p(p))
            self._own, self._opened = [], []

    # -- convenience constructors ----------------------------------------------------------------------------
    @classmethod
    def laplacian3d27(cls, n, group=None, mode="auto"):
        """BASELINE configs[1]: each rank generates its own nnz-balanced row block on its GPU."""
        group = group if group is not None else dist.group.WORLD
        L = lib()
        rows = n ** 3
        cuts = partition_by_prefix(lambda r: int(L.g4s_laplacian3d27_nnz(C.c_int(n), C.c_longlong(0), C.c_longlong(r))),
                                   rows, dist.get_world_size(group))
        r = dist.get_rank(group)
        A_local = CSR.laplacian3d27(n, cuts[r], cuts[r + 1])
        op = cls(A_local, cuts, group, mode)
        if op.mode != "peer":
            A_local.make_empty()  # the split blocks are what the product uses
        return op

    @classmethod
    def from_global(cls, A, group=None, mode="auto", ops=None):
        """Every rank holds the same host CSR `A` (square); it keeps the rows of its nnz-balanced cut."""
        group = group if group is not None else dist.group.WORLD
        A = A.to_host()
        world, r = dist.get_world_size(group), dist.get_rank(group)
        cuts = partition_rows(A.rowptr, world)
        s, e = int(A.rowptr[cuts[r]]), int(A.rowptr[cuts[r + 1]])
        local = CSR(cuts[r + 1] - cuts[r], A.cols, A.rowptr[cuts[r]:cuts[r + 1] + 1] - s, A.colids[s:e], A.values[s:e])
        return cls(local, cuts, group, mode, ops)

    # -- the product ---------------------------------------------------------------------------------------------
    def apply(self, x_local, y_local):
        """y_local = (A x)[owned rows]; x_local / y_local are this rank's slices (device tensors, float64)."""
        ops = self.ops
        if self.mode == "peer":
            b = self._k % 2
            self._k += 1
            xb = self.x_buffers[b]
            if x_local.data_ptr() != xb.data_ptr():
                xb.copy_(x_local)
            st = _stream_ptr(torch.cuda.current_stream())
            epoch = C.c_ulonglong(self._k)  # products are numbered from 1; flags start at 0
            # one launch: the kernel publishes this rank's epoch (release store into every peer's flag array) when it
            # starts and waits for the owners' flags only in the chunks that read their slices
            check(lib().g4s_spmv_partitioned_device(self.A.handle, C.c_int(self.world), C.c_int(self.rank),
                                                    self._parts[b], self._cuts_c, C.c_void_p(y_local.data_ptr()),
                                                    C.c_void_p(self._own[2]), epoch, self._flag_arrays, st))
            return y_local
        if ops.device_type != "cuda":
            return self._apply_sync(x_local, y_local)
        main = torch.cuda.current_stream()
        comm = self.comm_stream
        comm.wait_stream(main)  # x_local is ready on the main stream
        with torch.cuda.stream(comm):
            self._exchange(x_local, comm)
            done = comm.record_event()
        ops.spmv(self.diag, x_local, y_local, main)            # overlaps the exchange
        main.wait_event(done)
        ops.spmv(self.off, self.halo, y_local, main, accumulate=True)
        return y_local

    def apply_host(self, x_host, y_host):
        """Host-pointer product of the partitioned operator (peer mode), the multi-GPU twin of g4s_spmv_host: x_host / y_host
        are this rank's slices in PINNED host memory.  x is copied into the shared buffer the product reads; the kernel
        stores y straight into the pinned buffer (device-mapped under unified addressing: one coalesced 256-byte store per
        warp and 32 rows), so y travels over the host link WHILE the product runs instead of after it.  Returns when y is
        complete in host memory."""
        if self.mode != "peer":
            raise ValueError("apply_host needs peer mode")
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise ValueError("apply_host needs pinned host tensors")
        xb = self.next_x()
        xb.copy_(x_host, non_blocking=True)
        self.apply(xb, y_host)  # apply() only takes y's address
        torch.cuda.current_stream().synchronize()
        return y_host

    def _exchange(self, x_local, stream):
        if self.mode == "halo":
            if self.send_idx.numel():
                self.ops.gather(self.send_buf, x_local, self.send_idx, stream)
            dist.all_to_all_single(self.halo[:self.n_halo], self.send_buf, self.recv_counts, self.send_counts,
                                   group=self.group)
        else:
            self.x_pad[:self.local_rows].copy_(x_local)
            dist.all_gather_into_tensor(self.halo, self.x_pad, group=self.group)

    def _apply_sync(self, x_local, y_local):
        self._exchange(x_local, None)
        self.ops.spmv(self.diag, x_local, y_local, None)
        self.ops.spmv(self.off, self.halo, y_local, None, accumulate=True)
        return y_local


# ------------------------------------------------------------------------------------------------ SpGEMM
class DistSpGEMM:
    """C = A B with A's rows cut by intermediate-product balance and B replicated by NCCL broadcast from rank 0.

    `A_local` holds this rank's rows of A; `B` is the full matrix on rank 0 (None elsewhere)."""

    def __init__(self, A_local, B, cuts, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.cuts = [int(c) for c in cuts]
        self.A_local = A_local
        dev = torch.device("cuda", torch.cuda.current_device())
        meta = torch.zeros(3, dtype=torch.int64, device=dev)
        if self.rank == 0:
            meta[0], meta[1], meta[2] = B.rows, B.cols, B.nnz
        dist.broadcast(meta, 0, group=self.group)
        rows, cols, nnz = (int(v) for v in meta.tolist())
        if self.rank == 0:
            rp, ci, va = B.device_arrays()
            self._b = (torch.as_tensor(_DevArray(rp, rows + 1, "<i4"), device=dev),
                       torch.as_tensor(_DevArray(ci, nnz, "<i4"), device=dev),
                       torch.as_tensor(_DevArray(va, nnz, "<f8"), device=dev))
            self.B = B
        else:
            self._b = (torch.empty(rows + 1, dtype=torch.int32, device=dev),
                       torch.empty(nnz, dtype=torch.int32, device=dev),
                       torch.empty(nnz, dtype=torch.float64, device=dev))
        for t in self._b:
            dist.broadcast(t, 0, group=self.group)
        self.broadcast_bytes = 4 * (rows + 1) + 12 * nnz
        if self.rank != 0:
            h = C.c_void_p()
            check(lib().g4s_csr_create_device(C.byref(h), C.c_int(rows), C.c_int(cols), C.c_void_p(self._b[0].data_ptr()),
                                              C.c_void_p(self._b[1].data_ptr()), C.c_void_p(self._b[2].data_ptr()),
                                              C.c_void_p(0)))
            self.B = CSR._from_handle(h)

    def multiply(self, offsets=True):
        """Returns (C_local, global_nnz_offset): this rank's rows of C and where they start in the global CSR.
        offsets=False skips the all-gather of the per-rank nnz (a collective plus a host read per product) and returns
        (C_local, None): C stays distributed, and a loop that only needs the local rows never leaves the GPU."""
        C_local = HashSpGEMM(self.A_local, self.B)
        if not offsets:
            return C_local, None
        dev = torch.device("cuda", torch.cuda.current_device())
        mine = torch.full((1,), C_local.nnz, dtype=torch.int64, device=dev)  # a fill kernel: no pageable copy, no sync
        every = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        counts = every.tolist()                                              # the step's one host read
        self.global_nnz = int(sum(counts))
        return C_local, int(sum(counts[:self.rank]))


# ------------------------------------------------------------------------------------------------ BSR SpMM
def grid_pencil_order_local(n0, n1, n2, row0, row1, p0=4, p1=4):
    """Tile-major order (g4s_grid_pencil_order) of the rows [row0, row1) of an n0 x n1 x n2 grid, as LOCAL row numbers:
    the schedule for one rank's slab of a row-partitioned mesh matrix."""
    import numpy as np

    k0, k1 = row0 // (n0 * n1), -(-row1 // (n0 * n1))
    order = np.empty(n0 * n1 * (k1 - k0), dtype=np.int32)
    check(lib().g4s_grid_pencil_order(C.c_int(n0), C.c_int(n1), C.c_int(k1 - k0), C.c_int(p0), C.c_int(p1),
                                      order.ctypes.data_as(C.c_void_p), None, None))
    local = order.astype(np.int64) + k0 * n0 * n1 - row0
    return local[(local >= 0) & (local < row1 - row0)].astype(np.int32)


class DistBsrSpMM:
    """C = A_bsr B (3x3 blocks, 64 dense columns) with block rows cut over the ranks of one NVSwitch box.

    Every rank holds its block rows (GLOBAL block-column ids) and the matching rows of B in a CUDA-IPC-shared
    buffer (`self.B_local`, shape [local_block_rows * 3, 64]); the kernel reads the rows of B owned by other GPUs
    straight over NVLink (g4s_bsr3_spmm64_partitioned_device).  No halo exchange and no replicated B: at the
    256^3-node size of BASELINE config 5, B is 25.8 GB and a replica per GPU would cost more than the matrix."""

    def __init__(self, browptr, bcolids, bvalues, cuts, group=None, row_order=None, strips=None, kb=None):
        """strips = (strip_ptr, strip_rows) of the LOCAL block rows (g4s_b200.bsr.grid_pencil_strips) switches the product
        to the sliding-window sweep (g4s_bsr3_plan_*): the plan is built here, once, and packs the block values; call
        repack() after changing bvalues."""
        self.row_order = row_order  # optional device int32 permutation of the local block rows (tile-major schedule)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("peer-memory BSR SpMM supports up to 8 GPUs (one NVSwitch box)")
        self.cuts = [int(c) for c in cuts]
        self.mb = self.cuts[self.rank + 1] - self.cuts[self.rank]
        self.browptr, self.bcolids, self.bvalues = browptr, bcolids, bvalues  # device tensors (int32, int32, float64)
        dev = torch.device("cuda", torch.cuda.current_device())
        L = lib()
        ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
        nbytes = 8 * 3 * 64 * max(self.mb, 1)
        check(L.g4s_peer_alloc(C.c_size_t(nbytes), C.byref(ptr), h))
        self._own = ptr.value
        every = [None] * self.world
        dist.all_gather_object(every, bytes(h), group=self.group)
        self._opened = []
        self._parts = (C.c_void_p * self.world)()
        for q in range(self.world):
            if q == self.rank:
                self._parts[q] = self._own
            else:
                p = C.c_void_p()
                check(L.g4s_peer_open((C.c_ubyte * 64).from_buffer_copy(every[q]), C.byref(p)))
                self._parts[q] = p.value
                self._opened.append(p.value)
        self.B_local = torch.as_tensor(_DevArray(self._own, max(self.mb, 1) * 3 * 64, "<f8"), device=dev)[:self.mb * 3 * 64] \
            .view(self.mb * 3, 64)
        self._cuts_c = (C.c_int * (self.world + 1))(*self.cuts)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.plan = None
        if strips is not None:
            from .bsr import BsrPlan

            self.plan = BsrPlan(self.mb, self.cuts[-1] if kb is None else kb, browptr.data_ptr(), bcolids.data_ptr(), strips,
                                self.world, self.cuts)
            self.plan.set_values(bvalues.data_ptr())
            torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def repack(self):
        """New block values (same pattern): pack them into the sweep plan again."""
        if self.plan is not None:
            self.plan.set_values(self.bvalues.data_ptr(), torch.cuda.current_stream())

    def begin_update(self):
        """Call BEFORE writing new contents into B_local: a 4-byte all-reduce on the current stream orders the write after
        every peer's previous product, which reads this rank's slice over NVLink.  (apply(sync=True) only orders a product
        after the peers' WRITES; without this second point a fast rank would overwrite its slice while a slower peer's
        kernel is still reading it.)  Returns B_local."""
        dist.all_reduce(self._flag, group=self.group)
        return self.B_local

    def apply(self, C_local, sync=True):
        """C_local[local_block_rows * 3, 64] = (A B)[owned rows].  sync=True orders the product after every rank's
        writes to its slice of B with a 4-byte all-reduce on the current stream.  That is ONE of the two ordering points
        an iteration needs: before B_local is modified again call begin_update() (or order the write after all ranks'
        products by other means)."""
        if sync:
            dist.all_reduce(self._flag, group=self.group)
        if self.plan is not None:
            self.plan.spmm_partitioned(self._parts, C_local.data_ptr(), torch.cuda.current_stream())
            return C_local
        if self.row_order is not None:
            check(lib().g4s_bsr3_spmm64_partitioned_ordered_device(
                C.c_int(self.mb), C.c_void_p(self.browptr.data_ptr()), C.c_void_p(self.bcolids.data_ptr()),
                C.c_void_p(self.bvalues.data_ptr()), C.c_int(self.world), self._parts, self._cuts_c,
                C.c_void_p(C_local.data_ptr()), C.c_void_p(self.row_order.data_ptr()),
                _stream_ptr(torch.cuda.current_stream())))
            return C_local
        check(lib().g4s_bsr3_spmm64_partitioned_device(
            C.c_int(self.mb), C.c_void_p(self.browptr.data_ptr()), C.c_void_p(self.bcolids.data_ptr()),
            C.c_void_p(self.bvalues.data_ptr()), C.c_int(self.world), self._parts, self._cuts_c,
            C.c_void_p(C_local.data_ptr()), _stream_ptr(torch.cuda.current_stream())))
        return C_local

    def close(self):
        if self.plan is not None:
            self.plan.destroy()
            self.plan = None
        if self._own:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for p in self._opened:
                lib().g4s_peer_close(C.c_void_p(p))
            dist.barrier(group=self.group)
            lib().g4s_peer_free(C.c_void_p(self._own))
            self._own, self._opened = None, []
