#!/usr/bin/env python
"""Benchmark of the G4S mv/mm hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, libg4s_b200.so)
  python bench.py --impl reference [--gpus N] ...                the CPU arm (reference / oracle port)

Headline workload = BASELINE.json configs[1]: SpMV y = A x on the 3-D 27-point Laplacian, n = 400
(64 000 000 rows, 1 719 374 392 nnz, fp64 values, int32 indices), row-partitioned by nnz over N GPUs.
A step is one SpMV.  metric = achieved GB/s on the ALGORITHMIC bytes 12 nnz + 4(rows+1) + 8 cols + 8 rows.
The SpGEMM GFLOP/s of configs[3] (A x A, 2-D 5-point Laplacian, n = 2048) rides along in "spgemm" at N = 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner when the box sets
# NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the whole run and the line is written to the
# original stdout by emit().
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 400                       # configs[1]
SPGEMM_GRID = 2048                 # configs[3]
CPU_SAMPLE_ROWS = 6_400_000        # 10 % of the rows: the bounded sample the CPU legs run on
METRIC = "spmv_achieved_gbs"


def spmv_bytes(rows, cols, nnz):
    return 12.0 * nnz + 4.0 * (rows + 1) + 8.0 * cols + 8.0 * rows


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def wait_first_sample(self, timeout=8.0):
        """nvidia-smi takes up to a second to attach to every GPU of the box; the timed region must not start while
        it does (its start-up was inside the 10-20 ms timed region of the multi-GPU runs of round 1)."""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_rows_for_this_box():
    """The CPU arm runs the WHOLE configs[1] matrix (64 M rows: 20.6 GB of CSR + 1 GB of vectors on the host) when the box
    has the memory for it, else the first CPU_SAMPLE_ROWS rows.  Both arms evaluate this the same way, so their `config`
    blocks agree."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    return N_GRID ** 3 if avail >= 48e9 else CPU_SAMPLE_ROWS


def workload_config(n=N_GRID):
    """`config` of both arms: the workload, nothing about how an arm runs it."""
    rows, nnz = n ** 3, (3 * n - 2) ** 3
    cpu_rows = cpu_rows_for_this_box() if n == N_GRID else None
    cfg = {"workload": "SpMV y=Ax, 3-D 27-point Laplacian n=%d (BASELINE configs[1])" % n, "rows": rows, "nnz": nnz,
           "index_dtype": "int32", "bytes_per_step": spmv_bytes(rows, rows, nnz),
           "l2": "inputs (21.9 GB over all GPUs) exceed the 126 MB L2; no flush"}
    if cpu_rows is not None and cpu_rows != rows:
        cfg["workload"] += "; the CPU arm runs a bounded sample per step (first %d rows: host memory)" % cpu_rows
        cfg["cpu_arm_rows"] = cpu_rows
    return cfg


def cpu_spmv_sample(oracle, min_seconds, steps=None, warmup=1, rows=None):
    """The oracle's OpenMP CSR SpMV (port of "y = A x", mv/mv.c:23-27, on mm/inc/CSR.h's container — the
    reference's own mv is dense MKL and cannot hold this matrix) on the first `rows` rows of the same 27-point matrix
    (all of them when the host has the memory), all host threads.  Returns (GB/s, seconds per pass, description)."""
    rows = cpu_rows_for_this_box() if rows is None else rows
    A = oracle.gen_laplacian3d27(N_GRID, 0, rows)
    ncols_touched = min(N_GRID ** 3, rows + N_GRID * N_GRID + N_GRID + 1)  # x entries these rows read
    x = np.random.default_rng(12345).uniform(-1.0, 1.0, A[1])
    nnz = len(A[3])
    nbytes = 12.0 * nnz + 4.0 * (rows + 1) + 8.0 * ncols_touched + 8.0 * rows
    for _ in range(warmup):
        oracle.spmv_csr(A[2], A[3], A[4], x, omp=True)
    n, t0 = 0, time.perf_counter()
    while True:
        oracle.spmv_csr(A[2], A[3], A[4], x, omp=True)
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= min_seconds):
            break
    desc = ("%s of the n=%d 27-point Laplacian (%d nnz, %.2f GB algorithmic), %d passes of the OpenMP CSR row loop"
            % ("all %d rows" % rows if rows == N_GRID ** 3 else "rows [0,%d)" % rows, N_GRID, nnz, nbytes / 1e9, n))
    return nbytes * n / el / 1e9, el / n, desc


def cpu_spgemm(ref, oracle, grid):
    """Reference HashSpGEMM<false,true> (oracle/_ref) on configs[3]; falls back to the port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from matrices import laplacian_2d

    A = laplacian_2d(grid)
    total, _ = oracle.intprod(A[2], A[3], A[2])
    if ref is not None and ref.available:
        _, _, cnnz, secs = ref.hash_spgemm(A, A, variant=0, want_output=False)
        kind, cores = "reference", ref.omp_max_threads()
    else:
        out = oracle.hash_spgemm(A, A, threads=oracle.omp_max_threads())
        secs, kind, cores = out[3], "port", oracle.omp_max_threads()
    return {"value": 2.0 * total / secs / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": "full configs[3]: A x A, 2-D 5-point n=%d, one multiply (%.3f s)" % (grid, secs)}


def run_reference(args, rank):
    if rank != 0:
        return
    # rank 0 runs alone (the other ranks have exited) and uses every host core it may: torchrun exports
    # OMP_NUM_THREADS=1 to its children, which would turn the CPU arm into a single-thread number
    if "LOCAL_RANK" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    from oracle.binding import Oracle

    oracle = Oracle()
    cores = oracle.omp_max_threads()
    gbs, sec, desc = cpu_spmv_sample(oracle, 0.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- parity inside the bench
def det_x_numpy(idx):
    """x as a function of the GLOBAL index (exact in fp64): any rank and the CPU checker can regenerate any entry."""
    h = (idx.astype(np.uint64) * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    return h.astype(np.float64) / 4294967296.0 - 0.5


def det_x_torch(torch, lo, hi, device):
    idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
    h = (idx * 2654435761 + 12345) & 0xFFFFFFFF
    return h.to(torch.float64) / 4294967296.0 - 0.5


def spmv_parity_check(torch, dist, n, c0, c1, product, world, rank):
    """Correctness of the very operator that was timed, on the box it was timed on:
      (1) x = 1  =>  y[r] = 27 - (#stencil points of row r), exactly, for EVERY owned row (closed form of the matrix);
      (2) x = det_x(global index): stretches of owned rows — always including both ends of the rank's range, where
          the entries of x live on the neighbouring GPUs — against the oracle's CSR row loop on rows regenerated on
          the CPU, tolerance 1e-12 * sum_j |a_ij x_j|.
    `product(x_local_tensor) -> y_local_tensor` runs one product.  Returns the "parity_check" block (all ranks)."""
    from oracle.binding import Oracle

    dev = torch.device("cuda", torch.cuda.current_device())
    rows = c1 - c0
    y = product(torch.ones(rows, dtype=torch.float64, device=dev))
    r = torch.arange(c0, c1, dtype=torch.int64, device=dev)
    cnt = torch.ones(rows, dtype=torch.int64, device=dev)
    for q in (r % n, (r // n) % n, r // (n * n)):
        cnt *= 3 - ((q == 0) | (q == n - 1)).to(torch.int64)
    del r, q
    bad_ones = int((y != (27 - cnt).to(torch.float64)).sum().item())
    del cnt
    y = product(det_x_torch(torch, c0, c1, dev))
    oracle = Oracle()
    stretch = 4096
    nstretch = max(2, -(-26 // world))
    starts = sorted(set(int(v) for v in np.linspace(c0, max(c0, c1 - stretch), nstretch)))
    checked, worst = 0, 0.0
    for r0 in starts:
        r1 = min(r0 + stretch, c1)
        A = oracle.gen_laplacian3d27(n, r0, r1)
        lo, hi = int(A[3].min()), int(A[3].max()) + 1
        xw = det_x_numpy(np.arange(lo, hi, dtype=np.int64))
        col = (A[3] - lo).astype(np.int32)
        want = oracle.spmv_csr(A[2], col, A[4], xw)
        scale = oracle.spmv_csr_abs(A[2], col, A[4], xw)
        got = y[r0 - c0:r1 - c0].cpu().numpy()
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(scale, 1e-300))))
        checked += r1 - r0
    t = torch.tensor([float(bad_ones), float(checked), float(rows)], dtype=torch.float64, device=dev)
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    bad_ones, checked, rows_all, worst = int(t[0].item()), int(t[1].item()), int(t[2].item()), float(w.item())
    return {"ok": bad_ones == 0 and worst <= 1e-12,
            "ones_closed_form": {"rows": rows_all, "mismatches": bad_ones, "rule": "y = 27 - #stencil points, exact"},
            "sampled_vs_oracle": {"rows": checked, "max_err_over_abs_row_sum": worst, "tolerance": 1e-12,
                                  "x": "det_x(global index)", "stretches_per_rank": len(starts)}}


def per_step_times(torch, dist, step, steps, barrier):
    """Diagnostic pass AFTER the timed region: one CUDA event per step, per-rank median / p90 / max / first."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    barrier()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
    mine = torch.tensor([float(np.median(ms)), float(np.percentile(ms, 90)), float(ms.max()), float(ms[0]), float(ms.sum())],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(every, mine)
    else:
        every = [mine]
    keys = ("median", "p90", "max", "first", "sum")
    return {"note": "separate pass right after the timed region, one event per step; per rank",
            **{k: [round(float(e[i].item()), 5) for e in every] for i, k in enumerate(keys)}}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world):
    if "LOCAL_RANK" in os.environ:  # torchrun exports OMP_NUM_THREADS=1: give the host-side inspectors their share of the cores
        os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // max(world, 1)))
    import torch

    import g4s_b200
    from g4s_b200 import lib

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L = lib()
    if L.g4s_device_count() < 1:
        raise RuntimeError("no CUDA device; g4s_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    rows_total = n ** 3
    nnz_total = (3 * n - 2) ** 3
    total_bytes = spmv_bytes(rows_total, rows_total, nnz_total)

    if world == 1:
        A = g4s_b200.CSR.laplacian3d27(n)
        x_host = torch.empty(rows_total, dtype=torch.float64).pin_memory()
        x_host.numpy()[:] = np.random.default_rng(12345).uniform(-1.0, 1.0, rows_total)
        y_host = torch.empty(rows_total, dtype=torch.float64).pin_memory()
        x = x_host.cuda(non_blocking=True)
        y = torch.empty(rows_total, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()

        def step():
            A.spmv_device(x.data_ptr(), y.data_ptr())

        def e2e_step():
            A.spmv(x_host.numpy(), y_host.numpy())

        h2d, d2h = 8 * rows_total, 8 * rows_total
        launches_per_step = 1
        parallelism = "1 GPU"
    else:
        from g4s_b200.dist import DistSpMV

        op = DistSpMV.laplacian3d27(n, dist.group.WORLD, mode=args.mode)
        x_host = torch.empty(op.local_rows, dtype=torch.float64).pin_memory()
        x_host.numpy()[:] = np.random.default_rng(12345 + rank).uniform(-1.0, 1.0, op.local_rows)
        y_host = torch.empty(op.local_rows, dtype=torch.float64).pin_memory()
        x = x_host.cuda(non_blocking=True)
        y = torch.empty(op.local_rows, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        if op.mode == "peer":  # x lives in the rank's two CUDA-IPC-shared buffers; products alternate between them
            for xb in op.x_buffers:
                xb.copy_(x)

            def step():
                op.apply(op.next_x(), y)

            if os.environ.get("G4S_BENCH_E2E_STAGED"):  # upload, product, download one after the other
                def e2e_step():
                    xb = op.next_x()
                    xb.copy_(x_host, non_blocking=True)
                    op.apply(xb, y)
                    y_host.copy_(y, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            else:                                       # the kernel stores y into the pinned host buffer as it goes
                def e2e_step():
                    op.apply_host(x_host, y_host)
        else:
            def step():
                op.apply(x, y)

            def e2e_step():
                x.copy_(x_host, non_blocking=True)
                op.apply(x, y)
                y_host.copy_(y, non_blocking=True)
                torch.cuda.current_stream().synchronize()

        h2d, d2h = 8 * rows_total, 8 * rows_total  # summed over ranks
        launches_per_step = op.launches_per_step
        parallelism = {
            "peer": "rows cut by nnz over %d GPUs; ONE fused SpMV kernel per GPU loads remote x entries over NVLink from "
                    "CUDA-IPC peer memory, ordered by per-rank readiness flags (release/acquire over NVLink); no collective in the step",
            "halo": "rows cut by nnz over %d GPUs; halo of x exchanged with NCCL all_to_all, overlapped with the "
                    "diagonal-block product",
            "allgather": "rows cut by nnz over %d GPUs; x assembled with NCCL all-gather, overlapped with the "
                         "diagonal-block product"}[op.mode] % world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:  # running (attached to every GPU) well before the timed region starts
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = L.g4s_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.g4s_kernel_launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tl)
        launches = int(tl.item())
    ms_per_step = ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # the same loop once more with one event per step (diagnostic; same conditions as the timed region: the clock sampler
    # is still attached — detaching it perturbs the next few milliseconds on every GPU of the box)
    step_ms = per_step_times(torch, dist, step, args.steps, barrier)

    # end to end through the host-pointer API: matrix resident (created once, like mkl_sparse_d_create_csr in
    # the reference's mkl()), x from pinned host memory and y back to the host every step
    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = total_bytes / (e2e_s / args.steps) / 1e9

    # correctness of the operator that was just timed + a per-step look at the same loop (both outside the timed regions)
    if world == 1:
        c0, c1 = 0, rows_total

        def product(xt):
            A.spmv_device(xt.data_ptr(), y.data_ptr())
            return y
    else:
        c0, c1 = op.c0, op.c1

        def product(xt):
            if op.mode == "peer":
                xb = op.next_x()  # no barrier: the kernel's own flag protocol orders the peers' reads after this copy
                xb.copy_(xt)
                return op.apply(xb, y)
            return op.apply(xt, y)
    parity = None
    if not args.no_parity:
        # what the end-to-end call left in host memory == the device product of the same x (same kernel: bit for bit)
        e2e_step()
        x_dev = x_host.cuda()
        y_ref = product(x_dev).clone()
        torch.cuda.synchronize()
        e2e_same = torch.tensor([1.0 if torch.equal(y_ref.cpu(), y_host) else 0.0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(e2e_same, op=dist.ReduceOp.MIN)
        parity = spmv_parity_check(torch, dist, n, c0, c1, product, world, rank)
        parity["e2e_output_equals_device_product"] = bool(e2e_same.item() == 1.0)
        parity["ok"] = bool(parity["ok"] and parity["e2e_output_equals_device_product"])
        if world > 1 and op.mode == "peer":  # leave the shared buffers as the timed loop had them
            for xb in op.x_buffers:
                xb.copy_(x)
            torch.cuda.synchronize()
            dist.barrier()

    spgemm_multi = None
    if dist is not None and not args.no_spgemm:
        spgemm_multi = bench_spgemm_dist(g4s_b200, torch, dist, args, rank, world)
    bsr_line = None
    if not args.no_bsr:
        if world > 1:  # free the SpMV operator first: the BSR rider needs ~20 GB per GPU
            op.close()
            if hasattr(op, "A"):
                op.A.make_empty()
        else:
            A.make_empty()
        del x, y
        torch.cuda.empty_cache()
        try:
            bsr_line = bench_bsr(g4s_b200, torch, dist, rank, world, measured_peak()[0])
        except Exception as e:  # the headline line must not depend on a rider (collectives inside: all ranks fail alike)
            bsr_line = {"error": str(e)[:300]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (spmv_chunk_kernel): algorithmic bytes of one launch over its mean
    # duration inside the timed region (the fix-up / halo kernels are inside that time, so this is conservative)
    per_gpu_bytes = total_bytes / world
    achieved = per_gpu_bytes / (ms_per_step * 1e-3) / 1e9
    # DRAM traffic of the kernel comes from an ncu capture (it cannot be measured live); the capture records the SHA-256 of
    # the kernel's source file, and a capture of another build is not reported
    traffic, traffic_src = None, None
    try:
        import hashlib

        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            tj = json.load(f)
        with open(os.path.join(ROOT, "g4s_b200", "csrc", "spmv.cu"), "rb") as f:
            same_build = hashlib.sha256(f.read()).hexdigest() == tj.get("spmv_cu_sha256")
        if world == 1 and n == N_GRID and same_build:
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("kernel", "") + " — " + tj.get("capture", "")
        elif world == 1 and n == N_GRID:
            traffic_src = "stale: profiles/spmv_traffic.json was captured from another build of spmv.cu"
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "parallelism": parallelism,
        "gflops": 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9,
        "pct_of_8TBs": value / world / 8000.0 * 100.0,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "note": ("g4s_spmv_host: matrix resident on the GPU (g4s_csr handle created once, as MKL's create_csr "
                         "in the reference's mkl()); every step uploads x from pinned host memory and downloads y; "
                         "the call pipelines upload / row-block products / download on three streams") if dist is None else
                        ("DistSpMV.apply_host on every rank: the rank's slice of x is uploaded from pinned host memory into "
                         "its shared buffer, the fused kernel stores its slice of y straight into pinned host memory "
                         "(device-mapped), so the download overlaps the product" if args.mode == "peer" and
                         not os.environ.get("G4S_BENCH_E2E_STAGED") else
                         "per rank: upload of the x slice, product, download of the y slice, one after the other")},
        "gpu_launches": int(launches),
        "step_ms": step_ms,
        "parity_check": parity,
        "roofline": {"bound": "hbm", "kernel": "spmv_chunk_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "traffic_source": traffic_src},
    }
    if spgemm_multi is not None:
        line["spgemm"] = spgemm_multi
    if bsr_line is not None:
        line["bsr_spmm"] = bsr_line
    if world == 1 and not args.no_cpu:
        from oracle.binding import Oracle, Ref

        oracle = Oracle()
        gbs, sec, desc = cpu_spmv_sample(oracle, args.cpu_seconds)
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": oracle.omp_max_threads(), "kind": "port",
                                "sample": desc}
        if hasattr(L, "g4s_spgemm_device") and not args.no_spgemm:
            line["spgemm"] = bench_spgemm(g4s_b200, torch, args)
            try:
                ref = Ref()
            except Exception:
                ref = None
            line["spgemm"]["cpu_baseline"] = cpu_spgemm(ref, oracle, SPGEMM_GRID)
            try:
                line["spgemm"]["other_inputs"] = bench_spgemm_other(g4s_b200, torch)
            except Exception as e:
                line["spgemm"]["other_inputs"] = {"error": str(e)[:200]}
        if not args.no_other:
            try:
                line["other_configs"] = bench_other_configs(g4s_b200, torch, peak)
            except Exception as e:  # the headline line must not depend on the riders
                line["other_configs"] = {"error": str(e)[:200]}
            try:
                line["dense_mv"] = bench_dense_mv(g4s_b200, torch, peak)
            except Exception as e:
                line["dense_mv"] = {"error": str(e)[:200]}
            try:
                line["opt_matmul"] = bench_opt_matmul(g4s_b200, torch)
            except Exception as e:
                line["opt_matmul"] = {"error": str(e)[:200]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("parity check of the timed operator FAILED: %s" % json.dumps(parity))


def bench_other_configs(g4s_b200, torch, peak):
    """BASELINE configs[0] (2-D 5-point SpMV, L2-resident) and configs[2] (R-MAT scale 24 SpMV) on one GPU, device-timed
    (CUDA events, 3 warm-ups, 20 launches).  `frac` = algorithmic bytes / time over the measured HBM peak.  configs[4] has
    its own rider, bench_bsr()."""
    import ctypes as C

    import numpy as np

    from g4s_b200._lib import check
    from g4s_b200.dist import _DevArray

    L = g4s_b200.lib()

    def timeit(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {}
    for key, make, note in (
            ("configs[0] spmv 2-D 5-point n=1000", lambda: g4s_b200.CSR.laplacian2d(1000), "80 MB: L2-resident"),
            ("configs[2] spmv R-MAT scale 24 ef 16", lambda: g4s_b200.CSR.rmat(24, 16, seed=20240601), "duplicates summed")):
        A = make()
        nbytes, flops = A.spmv_cost()
        x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
        y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: A.spmv_device(x.data_ptr(), y.data_ptr()), 20)
        out[key] = {"rows": A.rows, "nnz": A.nnz, "ms": ms, "algorithmic_gbs": nbytes / ms / 1e6, "gflops": flops / ms / 1e6,
                    "frac_of_hbm_peak": nbytes / ms / 1e6 / peak, "note": note}
        A.make_empty()
        del x, y
    return out


BSR_GRID = {1: 128, 2: 161, 4: 203, 8: 256}   # configs[4]: 256^3 nodes on 8 GPUs; n^3 ~ 128^3 * N below (weak scaling)


def det_b_torch(torch, row0, row1, device):
    """Rows [row0, row1) of the dense operand, B[i, c] = det_x(64 i + c): regenerable anywhere."""
    return det_x_torch(torch, row0 * 64, row1 * 64, device)


def bench_bsr(g4s_b200, torch, dist, rank, world, peak, steps=10):
    """BASELINE configs[4]: C = A B, A = 3x3-block 27-point mesh operator (diagonal block 26 I + J, off-diagonal -I - 0.1 J;
    SURVEY.md §8d), B 64 dense columns, through the sliding-window sweep plan (g4s_bsr3_plan_*).  N GPUs: the n^3-node mesh
    (n = BSR_GRID[N]; 256^3 at N = 8) is cut into slabs of planes, every rank keeps its rows of B in CUDA-IPC memory and the
    stage loader's TMA copies read the two halo planes straight from the neighbours over NVLink; one 4-byte all-reduce per
    step orders the product after the peers' writes to B.  Device-timed, max over ranks; parity inside: B = 1 gives
    29 - 1.3 (#neighbours) in every entry (closed form, all rows), and stretches of rows at both ends and inside the slab
    are checked against the oracle's BSR product on rows regenerated on the CPU (tolerance 1e-12 |A||B|)."""
    import ctypes as C

    from g4s_b200 import bsr
    from g4s_b200._lib import check
    from g4s_b200.dist import DistBsrSpMM, _DevArray
    from oracle.binding import Oracle

    L = g4s_b200.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    n = BSR_GRID.get(world, int(round((128 ** 3 * world) ** (1.0 / 3.0))))
    plane = n * n
    kcut = [(n * q) // world for q in range(world + 1)]
    cuts = [k * plane for k in kcut]
    r0, r1 = cuts[rank], cuts[rank + 1]
    P = g4s_b200.CSR.laplacian3d27(n, r0, r1)
    rp, ci, va = P.device_arrays()
    nb, mb = P.nnz, P.rows
    vals = torch.as_tensor(_DevArray(va, nb, "<f8"), device=dev)
    J = torch.ones(3, 3, dtype=torch.float64, device=dev)
    I3 = torch.eye(3, dtype=torch.float64, device=dev)
    diag = (vals > 0).double()[:, None, None]
    blocks = (diag * (26 * I3 + J) + (1 - diag) * (-I3 - 0.1 * J)).contiguous().reshape(-1)
    del diag, vals
    strips = bsr.grid_pencil_strips(n, n, kcut[rank], kcut[rank + 1])
    Cd = torch.empty(mb * 192, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    if world == 1:
        plan = bsr.BsrPlan(mb, mb, rp, ci, strips).set_values(blocks.data_ptr())
        Bd = torch.empty(mb * 192, dtype=torch.float64, device=dev)
        info = plan.info()

        def step():
            plan.spmm(Bd.data_ptr(), Cd.data_ptr())
    else:
        op = DistBsrSpMM(torch.as_tensor(_DevArray(rp, mb + 1, "<i4"), device=dev), torch.as_tensor(_DevArray(ci, nb, "<i4"), device=dev),
                         blocks, cuts, strips=strips)
        Bd = op.B_local.view(-1)
        info = op.plan.info()

        def step():
            op.apply(Cd.view(mb * 3, 64))
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def set_b(t):
        if world > 1:
            op.begin_update()
        Bd.copy_(t)

    set_b(det_b_torch(torch, r0 * 3, r1 * 3, dev))
    for _ in range(3):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    # ---- parity ----------------------------------------------------------------------------------------------------------
    oracle = Oracle()
    got_all = Cd.view(mb * 3, 64)
    stretch, worst, checked = 512, 0.0, 0
    for s0 in sorted(set(int(v) for v in np.linspace(r0, max(r0, r1 - stretch), 5))):
        s1 = min(s0 + stretch, r1)
        A = oracle.gen_laplacian3d27(n, s0, s1)
        lo, hi = int(A[3].min()), int(A[3].max()) + 1
        hb = np.where(A[4] > 0, 1.0, 0.0)[:, None, None]
        hblocks = hb * (26 * np.eye(3) + np.ones((3, 3))) + (1 - hb) * (-np.eye(3) - 0.1 * np.ones((3, 3)))
        Bw = det_x_numpy(np.arange(lo * 192, hi * 192, dtype=np.int64)).reshape((hi - lo) * 3, 64)
        col = (A[3] - lo).astype(np.int32)
        want = oracle.bsr_spmm(A[2], col, hblocks.reshape(-1), 3, Bw)
        absw = oracle.bsr_spmm(A[2], col, np.abs(hblocks).reshape(-1), 3, np.abs(Bw))
        got = got_all[(s0 - r0) * 3:(s1 - r0) * 3].cpu().numpy()
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(absw, 1e-300))))
        checked += s1 - s0
    set_b(torch.ones(mb * 192, dtype=torch.float64, device=dev))
    step()
    torch.cuda.synchronize()
    r = torch.arange(r0, r1, dtype=torch.int64, device=dev)
    cnt = torch.ones(mb, dtype=torch.float64, device=dev)
    for q in (r % n, (r // n) % n, r // plane):
        cnt *= 3.0 - ((q == 0) | (q == n - 1)).double()
    want1 = 29.0 - 1.3 * (cnt - 1.0)
    bad = int(((got_all - want1[:, None].repeat_interleave(3, dim=0)).abs().amax(dim=1) > 1e-12 * (29.0 + 1.3 * 26)).sum().item())
    del r, cnt, want1
    bytes_local = 76.0 * nb + 4.0 * (mb + 1) + 2 * 8.0 * 192 * mb
    t = torch.tensor([ms, bytes_local, 18.0 * nb * 64, float(bad), float(checked), float(mb)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    ms_max, tot_bytes, tot_flops = float(tmax[0].item()), float(t[1].item()), float(t[2].item())
    bad, checked, rows_all, worst = int(t[3].item()), int(t[4].item()), int(t[5].item()), float(w.item())
    out = {"metric": "bsr_spmm_algorithmic_gbs", "value": tot_bytes / ms_max / 1e6, "unit": "GB/s", "ms": ms_max, "n_gpus": world,
           "tflops_fp64": tot_flops / ms_max / 1e9, "steps": steps,
           "config": {"workload": "BSR 3x3 SpMM x 64 columns, 27-block mesh operator on %d^3 nodes (BASELINE configs[4]%s)"
                                  % (n, "" if n == 256 else "; 256^3 is the 8-GPU size, n^3 ~ 128^3 N below"),
                      "block_rows": rows_all, "bytes": tot_bytes, "flops": tot_flops,
                      "partition": "slabs of planes; rows of B in CUDA-IPC memory, halo planes read over NVLink by the stage "
                                   "loader's TMA copies; one 4-byte all-reduce per step" if world > 1 else "1 GPU"},
           "plan": dict(info, setup_seconds=setup_s),
           "roofline": {"bound": "hbm", "kernel": "bsr3_sweep_kernel", "achieved": bytes_local / ms / 1e6 if world == 1
                        else tot_bytes / world / ms_max / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": (tot_bytes / world / ms_max / 1e6) / peak,
                        "note": "fp64 FMA-bound at this size as much as HBM-bound: 64.2 GFLOP per 10.7 GB; a warp issues one "
                                "DFMA per 2.5-2.6 cycles per scheduler at best on this part (tools/micro/dfma_rate.cu)"},
           "parity_check": {"ok": bad == 0 and worst <= 1e-12,
                            "ones_closed_form": {"block_rows": rows_all, "mismatches": bad},
                            "sampled_vs_oracle": {"block_rows": checked, "max_err_over_abs": worst, "tolerance": 1e-12}}}
    if world == 1:  # the row-wise kernel of round 1 beside it
        for _ in range(2):
            check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                        C.c_void_p(blocks.data_ptr()), C.c_int(64), C.c_void_p(Bd.data_ptr()),
                                        C.c_void_p(Cd.data_ptr()), C.c_void_p(0)))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            check(L.g4s_bsr_spmm_device(C.c_int(mb), C.c_int(mb), C.c_int(3), C.c_void_p(rp), C.c_void_p(ci),
                                        C.c_void_p(blocks.data_ptr()), C.c_int(64), C.c_void_p(Bd.data_ptr()),
                                        C.c_void_p(Cd.data_ptr()), C.c_void_p(0)))
        e1.record()
        torch.cuda.synchronize()
        out["rowwise_kpack_kernel_ms"] = e0.elapsed_time(e1) / 5
        plan.destroy()
    else:
        op.close()
    P.make_empty()
    return out


def bench_dense_mv(g4s_b200, torch, peak, dim=16384):
    """the literal mv/mv.c entry points (mv/mv.c:6-27) at a dense-feasible size.  GPU: the four
    operations on device-resident buffers (g4s_dense_mv_device; CUDA events, 3 warm-ups, 10 launches), bytes = the stored
    elements each one must read (8 dim^2 for dgemv, 4 dim (dim+1) for the triangular / symmetric ones) + the vectors.
    CPU: the reference's own matrix_multiply_* (oracle/_ref/libmv_ref.so = unmodified mv/mv.c on OpenBLAS, the MKL stand-in;
    the oracle's restatement when that library is absent), one call after one warm-up, all host threads OpenBLAS uses."""
    import ctypes as C

    from g4s_b200._lib import check
    from oracle.binding import Oracle, Ref

    L = g4s_b200.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand(dim * dim, dtype=torch.float64, device="cuda", generator=g)
    B = torch.ones(dim, dtype=torch.float64, device="cuda")
    Cv = torch.empty(dim, dtype=torch.float64, device="cuda")
    try:
        ref = Ref()
        impl, kind = (ref, "reference") if ref.mv_available else (Oracle(), "port")
    except Exception:
        impl, kind = Oracle(), "port"
    Ah = A.cpu().numpy()
    out = {"dim": dim, "cpu_kind": kind, "note": "dense dim x dim buffer as mv/mv.c builds it; dtrmv works in place on B"}
    for op, name, cpu_name, nbytes in ((0, "dgemv", "dgemv", 8.0 * dim * dim + 16.0 * dim),
                                       (1, "dsymv", "dsymv", 4.0 * dim * (dim + 1) + 16.0 * dim),
                                       (2, "dtrmv", "dtrmv", 4.0 * dim * (dim + 1) + 16.0 * dim),
                                       (3, "sspmv", "sspmv" if kind == "reference" else "dspmv", 4.0 * dim * (dim + 1) + 16.0 * dim)):
        def run():
            if op == 2:
                B.fill_(1.0 / dim)
            check(L.g4s_dense_mv_device(C.c_int(op), C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()),
                                        C.c_void_p(Cv.data_ptr()), C.c_int(dim), C.c_void_p(0)))
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
        Bh = np.full(dim, 1.0 / dim)
        impl.dense_mv(cpu_name, Ah, Bh)
        t0 = time.perf_counter()
        impl.dense_mv(cpu_name, Ah, Bh)
        cpu_s = time.perf_counter() - t0
        out["matrix_multiply_" + name] = {"gpu_ms": ms, "gpu_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                                          "cpu_ms": cpu_s * 1e3, "cpu_gbs": nbytes / cpu_s / 1e9}
    return out


def bench_opt_matmul(g4s_b200, torch):
    """OptMatmul res = xx w (deepmd/source/op/opt_matmul.cc:24-62) at the layer shapes of a DeePMD
    model (fitting net 240 x 240, embedding net 25 -> 50 -> 100) for M = atoms x frames rows.  GPU: device-resident, CUDA
    events, 3 warm-ups, 10 launches.  CPU: the reference's own engine loop (GraphProcess of deepmd/source/op/graph.h with the
    op's gather; 8 OpenMP threads, as the reference pins them) on the first 8192 rows."""
    from g4s_b200.opt_matmul import opt_matmul_device
    from oracle.binding import Oracle, Ref

    try:
        ref = Ref()
        impl, kind = (ref, "reference") if ref.available and hasattr(ref.lib, "ref_opt_matmul") else (Oracle(), "port")
    except Exception:
        impl, kind = Oracle(), "port"
    out = {"cpu_kind": kind, "cpu_rows": 8192}
    g = torch.Generator(device="cuda").manual_seed(9)
    for M, N, K in ((131072, 240, 240), (1048576, 25, 50), (1048576, 50, 100)):
        xx = torch.rand(M, N, dtype=torch.float64, device="cuda", generator=g) - 0.5
        w = torch.rand(N, K, dtype=torch.float64, device="cuda", generator=g) - 0.5
        res = torch.empty(M, K, dtype=torch.float64, device="cuda")
        for _ in range(3):
            opt_matmul_device(M, N, K, xx.data_ptr(), w.data_ptr(), res.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            opt_matmul_device(M, N, K, xx.data_ptr(), w.data_ptr(), res.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        xh, wh = xx[:8192].cpu().numpy(), w.cpu().numpy()
        impl.opt_matmul(xh, wh)
        t0 = time.perf_counter()
        want = impl.opt_matmul(xh, wh)
        cpu_s = time.perf_counter() - t0
        err = float(np.abs(res[:8192].cpu().numpy() - want).max())
        out["M=%d N=%d K=%d" % (M, N, K)] = {"gpu_ms": ms, "gpu_tflops": 2.0 * M * N * K / ms / 1e9,
                                           "gpu_gbs": 8.0 * (M * N + N * K + M * K) / ms / 1e6,
                                           "cpu_gflops": 2.0 * 8192 * N * K / cpu_s / 1e9, "max_abs_err_vs_cpu": err}
        del xx, w, res
    return out


def bench_spgemm_other(g4s_b200, torch):
    """The hash / dense-accumulator size classes of g4s_spgemm_device on inputs where they carry the work (VERDICT r01 item 4;
    configs[3] runs entirely in the merge class): A*A on the 3-D 27-point Laplacian n = 100 (705 M products, rows of 125
    distinct columns: warp-per-row tables) and on R-MAT scale 16 (power-law rows: CTA tables and the dense accumulator).
    Device-timed (3 warm-ups, 5 products), phases from one more product with the phase events on."""
    import ctypes as C

    L = g4s_b200.lib()
    out = {}
    for key, make in (("A*A 3-D 27-point n=100", lambda: g4s_b200.CSR.laplacian3d27(100)),
                      ("A*A R-MAT scale 16 ef 16", lambda: g4s_b200.CSR.rmat(16, 16, seed=20240601))):
        A = make()
        flop = 2.0 * g4s_b200.compute_flop(A, A)
        for _ in range(3):
            g4s_b200.HashSpGEMM(A, A).make_empty()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            Cm = g4s_b200.HashSpGEMM(A, A)
            nnzc = Cm.nnz
            Cm.make_empty()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ph = (C.c_double * 4)()
        L.g4s_spgemm_set_phase_timing(C.c_int(1))
        g4s_b200.HashSpGEMM(A, A).make_empty()
        L.g4s_spgemm_last_phase_ms(ph)
        L.g4s_spgemm_set_phase_timing(C.c_int(0))
        nbytes = 2 * (12.0 * A.nnz + 4.0 * (A.rows + 1)) + 12.0 * nnzc + 4.0 * (A.rows + 1)
        out[key] = {"rows": A.rows, "nnzA": A.nnz, "nnzC": nnzc, "ms": ms, "gflops": flop / ms / 1e6,
                    "algorithmic_gbs": nbytes / ms / 1e6,
                    "phase_ms": {"binning": ph[0], "symbolic": ph[1], "scan_alloc": ph[2], "numeric": ph[3]}}
        A.make_empty()
    return out


def bench_spgemm_dist(g4s_b200, torch, dist, args, rank, world):
    """configs[3] on N GPUs: A's rows cut over the ranks (rows of the 2-D Laplacian carry near-equal work, so the nnz
    cut is the work cut), B generated on rank 0 and replicated with NCCL broadcast, every rank multiplies its block."""
    import ctypes as C

    from g4s_b200.dist import DistSpGEMM, partition_by_prefix

    L = g4s_b200.lib()
    n = SPGEMM_GRID
    cuts = partition_by_prefix(lambda r: int(L.g4s_laplacian2d_nnz(C.c_int(n), C.c_longlong(0), C.c_longlong(r))),
                               n * n, world)
    A_local = g4s_b200.CSR.laplacian2d(n, cuts[rank], cuts[rank + 1])
    B = g4s_b200.CSR.laplacian2d(n) if rank == 0 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mm = DistSpGEMM(A_local, B, cuts)
    torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t0
    flop_local = 2.0 * g4s_b200.compute_flop(A_local, mm.B)
    mm.multiply()[0].make_empty()  # with the global offsets once (sets mm.global_nnz)
    for _ in range(3):
        mm.multiply(offsets=False)[0].make_empty()
    dist.barrier()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        Cl, _ = mm.multiply(offsets=False)  # C stays distributed: no collective, one host wait per product
        Cl.make_empty()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps, flop_local], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t)
    ms, flop = float(tmax[0].item()), float(t[1].item())
    # parity of the distributed product through closed forms of A*A on the 5-point Laplacian:
    # nnz = 13 n^2 - 20 n + 4, sum of all entries = 4 n + 8, every local row's columns strictly ascending
    from g4s_b200.dist import _DevArray
    Cl, _ = mm.multiply()
    rp, ci, va = Cl.device_arrays()
    vals = torch.as_tensor(_DevArray(va, Cl.nnz, "<f8"), device="cuda")
    cols = torch.as_tensor(_DevArray(ci, Cl.nnz, "<i4"), device="cuda")
    rpt = torch.as_tensor(_DevArray(rp, Cl.rows + 1, "<i4"), device="cuda").long()
    inner = torch.ones(Cl.nnz, dtype=torch.bool, device="cuda")
    inner[rpt[1:-1][rpt[1:-1] < Cl.nnz]] = False          # first entry of a row: no left neighbour in the same row
    unsorted = int(((cols[1:] <= cols[:-1]) & inner[1:]).sum().item())
    chk = torch.tensor([float(Cl.nnz), float(vals.sum().item()), float(unsorted)], dtype=torch.float64, device="cuda")
    dist.all_reduce(chk)
    Cl.make_empty()
    want_nnz, want_sum = 13 * n * n - 20 * n + 4, 4.0 * n + 8.0
    parity = {"ok": int(chk[0].item()) == want_nnz and abs(float(chk[1].item()) - want_sum) <= 1e-9 * want_sum
              and int(chk[2].item()) == 0,
              "global_nnz": int(chk[0].item()), "want_nnz": want_nnz, "sum_of_values": float(chk[1].item()),
              "want_sum": want_sum, "rows_with_unsorted_columns": int(chk[2].item())}
    return {"metric": "spgemm_gflops", "value": flop / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms": ms, "n_gpus": world,
            "parity_check": parity,
            "config": {"workload": "SpGEMM C=A*A, 2-D 5-point Laplacian n=%d (BASELINE configs[3]); A row-partitioned, "
                                   "B replicated by NCCL broadcast (%.1f ms, once), C left distributed" % (n, bcast_s * 1e3),
                       "nnzC": mm.global_nnz, "flop": flop}}


def bench_spgemm(g4s_b200, torch, args):
    """configs[3]: C = A A on the 2-D 5-point Laplacian n = 2048 (4 194 304 rows), device-timed, plus the
    host-pointer mkl() entry end to end."""
    import ctypes as C

    A = g4s_b200.CSR.laplacian2d(SPGEMM_GRID)
    flop = 2.0 * g4s_b200.compute_flop(A, A)
    for _ in range(3):
        g4s_b200.HashSpGEMM(A, A).make_empty()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        Cm = g4s_b200.HashSpGEMM(A, A)
        nnzc = Cm.nnz
        Cm.make_empty()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    phases = (C.c_double * 4)()
    g4s_b200.lib().g4s_spgemm_set_phase_timing(C.c_int(1))  # one more product with the per-phase events switched on
    g4s_b200.HashSpGEMM(A, A).make_empty()
    g4s_b200.lib().g4s_spgemm_last_phase_ms(phases)
    g4s_b200.lib().g4s_spgemm_set_phase_timing(C.c_int(0))
    nnza = A.nnz
    rows = A.rows
    nbytes = 2 * (12.0 * nnza + 4.0 * (rows + 1)) + 12.0 * nnzc + 4.0 * (rows + 1)
    Ah = A.to_host()
    t = g4s_b200.Timings()
    g4s_b200.mkl(Ah, Ah, t)
    g4s_b200.mkl(Ah, Ah, t)
    e2e_s = t.total  # seconds inside g4s_mkl: host CSR in -> host CSR out (upload, multiply, download, frees)
    return {"metric": "spgemm_gflops", "value": flop / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms": ms,
            "config": {"workload": "SpGEMM C=A*A, 2-D 5-point Laplacian n=%d (BASELINE configs[3])" % SPGEMM_GRID,
                       "rows": rows, "nnzA": nnza, "nnzC": nnzc, "flop": flop},
            "phase_ms": {"binning": phases[0], "symbolic": phases[1], "scan_alloc": phases[2], "numeric": phases[3]},
            "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "spgemm_merge_row_kernel<5,1> (numeric; every row of this product is in the "
                         "merge class)", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": measured_peak()[0], "unit": "GB/s",
                         "frac": nbytes / (ms * 1e-3) / 1e9 / measured_peak()[0],
                         "note": "algorithmic bytes of the whole product (A + B + C, SURVEY.md 8d) over the whole product's "
                                 "time (binning + symbolic + scan + numeric): conservative for the numeric kernel alone"},
            "e2e": {"value": flop / e2e_s / 1e9, "unit": "GFLOP/s", "seconds": e2e_s,
                    "h2d_bytes_per_step": 2 * (12 * nnza + 4 * (rows + 1)), "d2h_bytes_per_step": 12 * nnzc + 4 * (rows + 1),
                    "timings": {k: getattr(t, k) for k in ("create", "spmm", "export_csr", "destroy", "total")}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=N_GRID, help="grid edge (default 400 = BASELINE configs[1])")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline budget at N=1")
    ap.add_argument("--mode", default="peer", choices=["peer", "halo", "allgather", "auto"],
                    help="multi-GPU x assembly (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check of the timed operator")
    ap.add_argument("--no-spgemm", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the single-GPU riders for configs[0], [2]")
    ap.add_argument("--no-bsr", action="store_true", help="skip the configs[4] rider (BSR SpMM)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world != args.gpus and world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
