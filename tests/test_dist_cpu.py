"""world_size-2 gloo test of the multi-GPU exchange plan (g4s_b200/dist.py) on CPU, plus the closed-form
partitioner used by the device generators."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_dist_spmv_plan_world2_gloo():
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_cpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_partition_by_prefix_matches_c_partitioner(oracle):
    from g4s_b200.dist import partition_by_prefix, partition_rows

    A = oracle.gen_laplacian3d27(11)
    rp = A[2].astype(np.int64)
    for parts in (1, 2, 3, 8):
        assert partition_by_prefix(lambda r: int(rp[r]), A[0], parts) == partition_rows(rp, parts)
