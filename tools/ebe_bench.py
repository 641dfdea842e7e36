"""Developer tool: element-by-element operator on an n^3-node hex mesh (device-timed)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402
from g4s_b200._lib import check  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
ne = n - 1
nel = ne ** 3
idx = torch.arange(nel, device="cuda")
i, j, k = idx % ne, (idx // ne) % ne, idx // (ne * ne)
node = lambda a, b, c: (c * n + b) * n + a  # noqa: E731
ien = torch.stack([node(i, j, k), node(i + 1, j, k), node(i + 1, j + 1, k), node(i, j + 1, k), node(i, j, k + 1),
                   node(i + 1, j, k + 1), node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1)], dim=1)
dofs = (3 * ien[:, :, None] + torch.arange(3, device="cuda")[None, None, :]).reshape(nel, 24).to(torch.int32).contiguous()
elt_k = torch.rand(nel * 576, dtype=torch.float64, device="cuda") - 0.5
neq = 3 * n ** 3
u = torch.rand(neq, dtype=torch.float64, device="cuda") - 0.5
Au = torch.zeros(neq, dtype=torch.float64, device="cuda")
L = g4s_b200.lib()


def run():
    check(L.g4s_ebe_matvec_device(C.c_int(nel), C.c_int(24), C.c_void_p(elt_k.data_ptr()), C.c_void_p(dofs.data_ptr()),
                                  C.c_void_p(u.data_ptr()), C.c_void_p(Au.data_ptr()), C.c_void_p(0)))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = nel * (576 * 8 + 24 * 4) + 2 * 8 * neq
print("ebe n=%d elements=%d: %.3f ms  %.0f GB/s algorithmic (%.1f%% of 6528)  %.1f GFLOP/s" %
      (n, nel, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.284, 2.0 * 576 * nel / ms / 1e6))
