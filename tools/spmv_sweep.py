"""Developer tool: time the SpMV kernel variants / lane widths on a device-generated Laplacian (CUDA events).
usage: python tools/spmv_sweep.py [n=400] [iters=10]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g4s_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    variants = [int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0,1,2,3,4,9".split(","))]
    lanes_list = [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else "0,1,2,4,8,16".split(","))]
    A = g4s_b200.CSR.laplacian3d27(n)
    torch.cuda.synchronize()
    nbytes, flops = A.spmv_cost()
    print("n=%d rows=%d nnz=%d bytes=%.3f GB" % (n, A.rows, A.nnz, nbytes / 1e9), flush=True)
    x = torch.rand(A.cols, dtype=torch.float64, device="cuda") - 0.5
    y = torch.empty(A.rows, dtype=torch.float64, device="cuda")
    for variant in variants:
        for lanes in lanes_list:
            if variant == 9 and lanes != lanes_list[0]:
                continue
            A.set_tuning(lanes, variant)
            for _ in range(3):
                A.spmv_device(x.data_ptr(), y.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                A.spmv_device(x.data_ptr(), y.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print("variant %d lanes %2d : %.3f ms  %.1f GB/s  (%.1f%% of 6528)" %
                  (variant, lanes, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.28), flush=True)


if __name__ == "__main__":
    main()
