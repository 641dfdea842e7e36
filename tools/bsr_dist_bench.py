"""BASELINE config 5 on N GPUs (torchrun): BSR 3x3 SpMM x 64 columns on an n^3-node hex mesh (27-block stencil),
block rows and B partitioned, remote rows of B read over NVLink.  usage: torchrun ... tools/bsr_dist_bench.py [n=256]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402

import g4s_b200  # noqa: E402
from g4s_b200.dist import DistBsrSpMM, _DevArray, grid_pencil_order_local, partition_by_prefix  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L = g4s_b200.lib()
rows = n ** 3
cuts = partition_by_prefix(lambda r: int(L.g4s_laplacian3d27_nnz(C.c_int(n), C.c_longlong(0), C.c_longlong(r))), rows, world)
P = g4s_b200.CSR.laplacian3d27(n, cuts[rank], cuts[rank + 1])
rp, ci, va = P.device_arrays()
nb, mb = P.nnz, P.rows
dev = torch.device("cuda", local)
browptr = torch.as_tensor(_DevArray(rp, mb + 1, "<i4"), device=dev)
bcol = torch.as_tensor(_DevArray(ci, nb, "<i4"), device=dev)
vals = torch.as_tensor(_DevArray(va, nb, "<f8"), device=dev)
J = torch.ones(3, 3, dtype=torch.float64, device=dev)
I3 = torch.eye(3, dtype=torch.float64, device=dev)
blocks = torch.empty(nb * 9, dtype=torch.float64, device=dev)
step = 1 << 24
for s in range(0, nb, step):  # chunked: the broadcast temporaries stay small
    d = (vals[s:s + step] > 0).double()[:, None, None]
    blocks[s * 9:(s + d.shape[0]) * 9] = (d * (26 * I3 + J) + (1 - d) * (-I3 - 0.1 * J)).reshape(-1)
op = DistBsrSpMM(browptr, bcol, blocks, cuts)
g = torch.Generator(device=dev).manual_seed(777 + rank)
op.B_local.copy_(torch.rand(mb * 3, 64, dtype=torch.float64, device=dev, generator=g) * 2 - 1)
Cl = torch.empty(mb * 3, 64, dtype=torch.float64, device=dev)
order = torch.from_numpy(grid_pencil_order_local(n, n, n, cuts[rank], cuts[rank + 1])).to(dev)
tot = torch.tensor([float(nb)], dtype=torch.float64, device=dev)
dist.all_reduce(tot)
results = {}
for mode in ("dfma", "kpack", "ordered"):  # plain DFMA kernel, K-packed kernel, K-packed kernel in tile-major row order
    L.g4s_bsr_spmm_set_variant(C.c_int(4 if mode == "kpack" else 1))
    op.row_order = order if mode == "ordered" else None
    for _ in range(2):
        op.apply(Cl)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op.apply(Cl)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    results[mode] = (float(t.item()), float(Cl.double().sum().item()))
L.g4s_bsr_spmm_set_variant(C.c_int(0))
if rank == 0:
    nbt = float(tot.item())
    nbytes = 76.0 * nbt + 4.0 * (rows + 1) + 2 * 8.0 * 3 * rows * 64
    for mode, (ms, chk) in results.items():
        print(json.dumps({"workload": "BSR 3x3 SpMM x 64 cols, %d^3 nodes (BASELINE configs[4])" % n, "kernel": mode,
                          "n_gpus": world, "blocks": nbt, "ms": ms, "algorithmic_gbs": nbytes / ms / 1e6,
                          "tflops": 2 * 9 * nbt * 64 / ms / 1e9,
                          "pct_of_hbm_roofline": nbytes / ms / 1e6 / (6528.4 * world) * 100, "checksum_rank0": chk}), flush=True)
op.close()
dist.destroy_process_group()
