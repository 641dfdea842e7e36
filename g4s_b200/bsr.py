"""Host-side mirror of the BSR 3x3 SpMM x 64 inspector / executor (include/g4s_b200.h, csrc/bsr_sweep.cu): the role of
MKL's mkl_sparse_set_mm_hint + mkl_sparse_optimize for the CitcomS-style node operator of BASELINE config 5
(citcoms/lib/Element_calculations.c:516-571)."""
import ctypes as C

import numpy as np

from ._lib import check, i32p, lib


def _stream_ptr(stream):
    return C.c_void_p(stream.cuda_stream if stream is not None else 0)


def grid_pencil_strips(n0, n1, k_begin, k_end, p0=4, p1=4):
    """(strip_ptr, strip_rows) for the nodes k_begin <= k < k_end of an n0 x n1 x * grid: one strip per grid line along the
    third axis, LOCAL row numbers, the lines of a p0 x p1 patch consecutive (g4s_grid_pencil_strips)."""
    sp = np.empty(n0 * n1 + 1, dtype=np.int32)
    sr = np.empty(n0 * n1 * (k_end - k_begin), dtype=np.int32)
    check(lib().g4s_grid_pencil_strips(C.c_int(n0), C.c_int(n1), C.c_int(k_begin), C.c_int(k_end), C.c_int(p0), C.c_int(p1),
                                       sp.ctypes.data_as(i32p), sr.ctypes.data_as(i32p)))
    return sp, sr


class BsrPlan:
    """Sliding-window sweep plan for C = A_bsr B (3x3 blocks, 64 columns).  browptr / bcolids / bvalues / B / C are raw
    DEVICE pointers (ints); strips = (strip_ptr, strip_rows) numpy int32 arrays or None (runs of 64 consecutive rows)."""

    def __init__(self, mb, kb, browptr_ptr, bcolids_ptr, strips=None, world=1, cuts=None):
        self.handle = C.c_void_p()
        self.mb, self.kb, self.world = mb, kb, world
        self._cuts = (C.c_int * (world + 1))(*[int(c) for c in cuts]) if cuts is not None else None
        if strips is None:
            n, sp, sr = 0, None, None
        else:
            sp = np.ascontiguousarray(strips[0], dtype=np.int32)
            sr = np.ascontiguousarray(strips[1], dtype=np.int32)
            n = len(sp) - 1
        check(lib().g4s_bsr3_plan_create(C.byref(self.handle), C.c_int(mb), C.c_int(kb), C.c_void_p(browptr_ptr),
                                         C.c_void_p(bcolids_ptr), C.c_int(n),
                                         sp.ctypes.data_as(i32p) if sp is not None else None,
                                         sr.ctypes.data_as(i32p) if sr is not None else None, C.c_int(world), self._cuts))

    def set_values(self, bvalues_ptr, stream=None):
        check(lib().g4s_bsr3_plan_set_values(self.handle, C.c_void_p(bvalues_ptr), _stream_ptr(stream)))
        return self

    def spmm(self, B_ptr, C_ptr, stream=None):
        check(lib().g4s_bsr3_plan_spmm64_device(self.handle, C.c_void_p(B_ptr), C.c_void_p(C_ptr), _stream_ptr(stream)))

    def spmm_partitioned(self, parts, C_ptr, stream=None):
        """parts: ctypes array of world device pointers (rank q's rows of B; own or CUDA-IPC peer memory)."""
        check(lib().g4s_bsr3_plan_spmm64_partitioned_device(self.handle, C.c_int(self.world), parts, self._cuts,
                                                            C.c_void_p(C_ptr), _stream_ptr(stream)))

    def info(self):
        fill, nbytes, ns, nt, sm = C.c_double(), C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
        check(lib().g4s_bsr3_plan_info(self.handle, C.byref(fill), C.byref(nbytes), C.byref(ns), C.byref(nt), C.byref(sm)))
        return {"slot_fill": fill.value, "stream_bytes": nbytes.value, "stages": ns.value, "tiles": nt.value,
                "stage_smem_bytes": sm.value}

    def destroy(self):
        if self.handle:
            lib().g4s_bsr3_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def inspect_host(mb, kb, browptr, bcolids, strips=None, world=1, cuts=None, grid=296, ctas_per_sm=2):
    """The schedule alone from host arrays (g4s_bsr3_plan_inspect_host): no device is touched."""
    L = lib()
    rp = np.ascontiguousarray(browptr, dtype=np.int32)
    ci = np.ascontiguousarray(bcolids, dtype=np.int32)
    if strips is None:
        n, sp, sr = 0, None, None
    else:
        sp = np.ascontiguousarray(strips[0], dtype=np.int32)
        sr = np.ascontiguousarray(strips[1], dtype=np.int32)
        n = len(sp) - 1
    cc = (C.c_int * (world + 1))(*[int(c) for c in cuts]) if cuts is not None else None
    ns, nt, sb, sm, fill = C.c_int(), C.c_int(), C.c_longlong(), C.c_int(), C.c_double()
    i64p = C.POINTER(C.c_longlong)
    st, tp, prod, meta, moff, base = i64p(), i32p(), i32p(), i32p(), i64p(), i64p()
    check(L.g4s_bsr3_plan_inspect_host(C.c_int(mb), C.c_int(kb), rp.ctypes.data_as(i32p), ci.ctypes.data_as(i32p), C.c_int(n),
                                       sp.ctypes.data_as(i32p) if sp is not None else None,
                                       sr.ctypes.data_as(i32p) if sr is not None else None, C.c_int(world), cc,
                                       C.c_int(grid), C.c_int(ctas_per_sm),
                                       C.byref(ns), C.byref(nt), C.byref(sb), C.byref(sm), C.byref(fill), C.byref(st),
                                       C.byref(tp), C.byref(prod), C.byref(meta), C.byref(moff), C.byref(base)))

    def take(ptr, count, dtype):
        out = np.ctypeslib.as_array(ptr, shape=(max(count, 1),))[:count].astype(dtype, copy=True)
        L.g4s_free(C.cast(ptr, C.c_void_p))
        return out

    nstages, nb = ns.value, int(rp[mb]) if mb else 0
    moff_a = take(moff, nstages + 1, np.int64)
    return {"nstages": nstages, "ntiles": nt.value, "stream_bytes": sb.value, "stage_smem_bytes": sm.value,
            "slot_fill": fill.value, "stage_table": take(st, 2 * nstages, np.int64).reshape(nstages, 2),
            "cta_ptr": take(tp, grid + 1, np.int32), "prod": take(prod, 64 * nstages, np.int32).reshape(nstages, 64),
            "grid": grid, "ctas_per_sm": ctas_per_sm, "meta": take(meta, int(moff_a[-1]), np.int32),
            "meta_off": moff_a, "base": take(base, nb, np.int64)}
