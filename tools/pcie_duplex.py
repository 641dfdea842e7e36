"""Developer tool: what the host link gives — 512 MB up alone, down alone, and both at once (pinned memory, two streams)."""
import time

import torch

n = 64_000_000
hx = torch.empty(n, dtype=torch.float64).pin_memory()
hy = torch.empty(n, dtype=torch.float64).pin_memory()
dx = torch.empty(n, dtype=torch.float64, device="cuda")
dy = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for _ in range(2):
    run(True, True, 2)
print("up alone   %.2f ms  (%.1f GB/s)" % (run(True, False), 0.512 / run(True, False) * 1e3))
print("down alone %.2f ms  (%.1f GB/s)" % (run(False, True), 0.512 / run(False, True) * 1e3))
both = run(True, True)
print("both       %.2f ms  (%.1f GB/s each way)" % (both, 0.512 / both * 1e3))


def pieces(nb, reps=10, chain=True):
    """nb pieces each way; with chain, piece b goes down only after piece b has come up (the host pipeline's shape)."""
    step = n // nb
    evs = [torch.cuda.Event() for _ in range(nb)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for b in range(nb):
            sl = slice(b * step, (b + 1) * step)
            with torch.cuda.stream(s1):
                dx[sl].copy_(hx[sl], non_blocking=True)
                evs[b].record(s1)
            with torch.cuda.stream(s2):
                if chain:
                    s2.wait_event(evs[b])
                hy[sl].copy_(dy[sl], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for nb in (4, 8, 16, 32, 64):
    print("pieces %3d: chained %.2f ms, free %.2f ms" % (nb, pieces(nb), pieces(nb, chain=False)))
