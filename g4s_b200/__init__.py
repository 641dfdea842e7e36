"""g4s_b200 — B200-native (sm_100a) implementation of G4S's mv/ (matrix-vector) and mm/ (matrix-matrix)
sparse linear-algebra hot path.  The product is the CUDA library g4s_b200/libg4s_b200.so behind the C ABI of
include/g4s_b200.h; this package is the thin host-side mirror of the reference's interface used by the tests,
the benchmark and the multi-GPU (torch.distributed / NCCL) plumbing."""
from ._lib import G4SError, Timings, lib  # noqa: F401
from .csr import CSR, HashSpGEMM, HeapSpGEMM, OuterSpGEMM, bsr_from_citcoms_nodes, compute_flop, mkl, spmv_csr_f64  # noqa: F401
from . import mv  # noqa: F401
from .opt_matmul import opt_matmul  # noqa: F401
