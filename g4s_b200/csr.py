"""Host-side mirror of the reference's CSR container and SpGEMM entry points (mm/inc/CSR.h, hash_mult.h,
mkl_mult.h), bound to the C ABI of include/g4s_b200.h.  All arithmetic happens in libg4s_b200.so on the GPU."""
import ctypes as C

import numpy as np

from ._lib import Timings, check, f64p, i32p, lib, longp

EPSILON = 0.001  # mm/inc/utility.h:16


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(i32p)


def _dp(a):
    return a.ctypes.data_as(f64p)


def _take(ptr, n, dtype):
    arr = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype=dtype)
    lib().g4s_free(C.cast(ptr, C.c_void_p))
    return arr


def _stream_ptr(stream):
    if stream is None:
        return C.c_void_p(0)
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))


class CSR:
    """CSR<int,double> (mm/inc/CSR.h:22-100): rows, cols, nnz, rowptr, colids, values, 0-based.

    Host arrays are numpy; `handle` is the device-resident twin (g4s_csr_t), created on first use.  A CSR made
    by a device generator or by a device SpGEMM has only the handle until `to_host()` is called."""

    def __init__(self, rows=0, cols=0, rowptr=None, colids=None, values=None):
        self.rows, self.cols = int(rows), int(cols)
        self.rowptr = _i32(rowptr) if rowptr is not None else None
        self.colids = _i32(colids) if colids is not None else None
        self.values = _f64(values) if values is not None else None
        self.zerobased = True
        self._h = None
        if self.rowptr is not None:
            if len(self.rowptr) != self.rows + 1:
                raise ValueError("rowptr must have rows+1 entries")
            if len(self.colids) != self.rowptr[-1] or len(self.values) != self.rowptr[-1]:
                raise ValueError("colids/values must have rowptr[rows] entries")

    # ---- constructors mirroring the reference ---------------------------------------------------------
    @classmethod
    def construct(cls, filename, cache=False):
        """CSR::construct (mm/inc/CSR.h:485-669): MatrixMarket coordinate file -> CSR.  cache=True keeps / reuses the
        binary twin `<filename>.g4scsr` (g4s_csr_read_cached); `self.cache_hit` tells which path ran."""
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        rp, ci, va = i32p(), i32p(), f64p()
        hit = C.c_int(0)
        if cache:
            check(lib().g4s_csr_read_cached(str(filename).encode(), C.byref(rows), C.byref(cols), C.byref(nnz),
                                            C.byref(rp), C.byref(ci), C.byref(va), C.byref(hit)))
        else:
            check(lib().g4s_csr_read_matrix_market(str(filename).encode(), C.byref(rows), C.byref(cols), C.byref(nnz),
                                                   C.byref(rp), C.byref(ci), C.byref(va)))
        self = cls(rows.value, cols.value, _take(rp, rows.value + 1, np.int32), _take(ci, nnz.value, np.int32),
                   _take(va, nnz.value, np.float64))
        self.cache_hit = bool(hit.value)
        return self

    @classmethod
    def load_binary(cls, filename):
        """Binary CSR cache file (g4s_csr_read_binary) -> CSR."""
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        rp, ci, va = i32p(), i32p(), f64p()
        check(lib().g4s_csr_read_binary(str(filename).encode(), C.byref(rows), C.byref(cols), C.byref(nnz),
                                        C.byref(rp), C.byref(ci), C.byref(va)))
        return cls(rows.value, cols.value, _take(rp, rows.value + 1, np.int32), _take(ci, nnz.value, np.int32),
                   _take(va, nnz.value, np.float64))

    def save_binary(self, filename):
        """CSR -> binary cache file (g4s_csr_write_binary)."""
        if self.rowptr is None:
            self.to_host()
        check(lib().g4s_csr_write_binary(str(filename).encode(), C.c_int(self.rows), C.c_int(self.cols),
                                         C.c_int(len(self.colids)), _ip(self.rowptr), _ip(self.colids), _dp(self.values)))

    @classmethod
    def from_graph(cls, n, start, end, w):
        """CSR(graph&) (mm/inc/CSR.h:255-329): edge list grouped by start vertex, duplicates summed."""
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        w = _f64(w)
        nnz = C.c_int()
        rp, ci, va = i32p(), i32p(), f64p()
        check(lib().g4s_csr_from_edge_list(C.c_long(len(start)), C.c_long(n), start.ctypes.data_as(longp),
                                           end.ctypes.data_as(longp), _dp(w), C.byref(nnz), C.byref(rp),
                                           C.byref(ci), C.byref(va)))
        return cls(n, n, _take(rp, n + 1, np.int32), _take(ci, nnz.value, np.int32), _take(va, nnz.value, np.float64))

    @classmethod
    def _from_handle(cls, h):
        self = cls()
        self._h = h
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_longlong()
        check(lib().g4s_csr_shape(h, C.byref(rows), C.byref(cols), C.byref(nnz)))
        self.rows, self.cols, self._nnz_dev = rows.value, cols.value, nnz.value
        return self

    @classmethod
    def laplacian2d(cls, n, row0=0, row1=-1, stream=None):
        """2-D 5-point Laplacian (diag 4, off-diag -1), rows [row0,row1), generated on the device."""
        h = C.c_void_p()
        check(lib().g4s_csr_generate_laplacian2d(C.byref(h), C.c_int(n), C.c_longlong(row0), C.c_longlong(row1),
                                                 _stream_ptr(stream)))
        return cls._from_handle(h)

    @classmethod
    def laplacian3d27(cls, n, row0=0, row1=-1, stream=None):
        """3-D 27-point Laplacian (diag 26, off-diag -1), rows [row0,row1), generated on the device."""
        h = C.c_void_p()
        check(lib().g4s_csr_generate_laplacian3d27(C.byref(h), C.c_int(n), C.c_longlong(row0), C.c_longlong(row1),
                                                   _stream_ptr(stream)))
        return cls._from_handle(h)

    @classmethod
    def rmat(cls, scale, edge_factor=16, seed=20240601, stream=None):
        h = C.c_void_p()
        check(lib().g4s_csr_generate_rmat(C.byref(h), C.c_int(scale), C.c_int(edge_factor), C.c_ulonglong(seed),
                                          _stream_ptr(stream)))
        return cls._from_handle(h)

    # ---- container behaviour ---------------------------------------------------------------------------
    @property
    def nnz(self):
        if self.rowptr is not None:
            return int(self.rowptr[-1])
        return int(getattr(self, "_nnz_dev", 0))

    @property
    def handle(self):
        if self._h is None:
            if self.rowptr is None:
                raise ValueError("empty CSR")
            h = C.c_void_p()
            check(lib().g4s_csr_create_host(C.byref(h), C.c_int(self.rows), C.c_int(self.cols), _ip(self.rowptr),
                                            _ip(self.colids), _dp(self.values)))
            self._h = h
        return self._h

    def to_host(self):
        if self.rowptr is None:
            n = self.nnz
            self.rowptr = np.empty(self.rows + 1, dtype=np.int32)
            self.colids = np.empty(n, dtype=np.int32)
            self.values = np.empty(n, dtype=np.float64)
            check(lib().g4s_csr_download(self._h, _ip(self.rowptr), _ip(self.colids), _dp(self.values)))
        return self

    def device_arrays(self):
        rp, ci, va = i32p(), i32p(), f64p()
        check(lib().g4s_csr_device_arrays(self.handle, C.byref(rp), C.byref(ci), C.byref(va)))
        return (C.cast(rp, C.c_void_p).value, C.cast(ci, C.c_void_p).value, C.cast(va, C.c_void_p).value)

    def make_empty(self):
        """CSR::make_empty (mm/inc/CSR.h:50-62)."""
        if self._h is not None:
            lib().g4s_csr_destroy(self._h)
            self._h = None
        self.rowptr = self.colids = self.values = None
        self.rows = self.cols = 0
        self._nnz_dev = 0

    def __del__(self):
        try:
            if self._h is not None:
                lib().g4s_csr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def isEmpty(self):
        return self.nnz == 0

    def __eq__(self, rhs):
        """CSR::operator== (mm/inc/CSR.h:343-408): shape and structure exact, values within EPSILON abs-or-rel."""
        a, b = self.to_host(), rhs.to_host()
        if a.nnz != b.nnz or a.rows != b.rows or a.cols != b.cols:
            return False
        if not (np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.colids, b.colids)):
            return False
        d = np.abs(a.values - b.values)
        m = np.maximum(np.abs(a.values), np.abs(b.values))
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = (a.values == b.values) | (d < EPSILON) | (d / m < EPSILON)
        return bool(np.all(ok))

    __hash__ = None

    def submatrix(self, M_, N_, M_start=0, N_start=0):
        """CSR(const CSR&, M_, N_, M_start, N_start) (mm/inc/CSR.h:691-733)."""
        a = self.to_host()
        nnz = C.c_int()
        rp, ci, va = i32p(), i32p(), f64p()
        check(lib().g4s_csr_submatrix(C.c_int(a.rows), C.c_int(a.cols), _ip(a.rowptr), _ip(a.colids), _dp(a.values),
                                      C.c_int(M_), C.c_int(N_), C.c_int(M_start), C.c_int(N_start), C.byref(nnz),
                                      C.byref(rp), C.byref(ci), C.byref(va)))
        return CSR(M_, N_, _take(rp, M_ + 1, np.int32), _take(ci, nnz.value, np.int32),
                   _take(va, nnz.value, np.float64))

    # ---- SpMV ------------------------------------------------------------------------------------------
    def set_tuning(self, lanes_per_row=0, variant=0):
        check(lib().g4s_spmv_set_tuning(self.handle, C.c_int(lanes_per_row), C.c_int(variant)))
        return self

    def spmv(self, x, y=None):
        """y = A x with host vectors (H2D x, kernel, D2H y inside the call)."""
        x = _f64(x)
        if len(x) != self.cols:
            raise ValueError("x must have cols entries")
        if y is None:
            y = np.empty(self.rows, dtype=np.float64)
        check(lib().g4s_spmv_host(self.handle, _dp(x), _dp(y)))
        return y

    def spmv_device(self, x_ptr, y_ptr, stream=None, row_map_ptr=None, accumulate=False):
        """y = A x on raw device pointers (ints), asynchronous on `stream`."""
        if row_map_ptr is None and not accumulate:
            check(lib().g4s_spmv_device(self.handle, C.c_void_p(x_ptr), C.c_void_p(y_ptr), _stream_ptr(stream)))
        else:
            check(lib().g4s_spmv_device_ex(self.handle, C.c_void_p(x_ptr), C.c_void_p(y_ptr),
                                           C.c_void_p(row_map_ptr or 0), C.c_int(1 if accumulate else 0),
                                           _stream_ptr(stream)))

    def spmv_cost(self):
        b, f = C.c_double(), C.c_double()
        check(lib().g4s_spmv_cost(self.handle, C.byref(b), C.byref(f)))
        return b.value, f.value


def spmv_csr_f64(rows, cols, rowptr, colids, values, x):
    """One-shot host entry: upload, multiply, download, free."""
    rowptr, colids, values, x = _i32(rowptr), _i32(colids), _f64(values), _f64(x)
    y = np.empty(rows, dtype=np.float64)
    check(lib().g4s_spmv_csr_f64(C.c_int(rows), C.c_int(cols), _ip(rowptr), _ip(colids), _dp(values), _dp(x), _dp(y)))
    return y


def compute_flop(A, B):
    """compute_flop / get_flop (mm/inc/mkl_mult.h:31-38, hash_mult.h:45-62): intermediate products of A*B."""
    total = C.c_longlong()
    check(lib().g4s_compute_flop_device(A.handle, B.handle, C.byref(total), None, C.c_void_p(0)))
    return total.value


def HashSpGEMM(a, b, stream=None):
    """HashSpGEMM(a, b, c, multiplies, plus) (mm/inc/hash_mult.h:1028-1057, :1109-1113): returns c = a b with
    sorted columns, device-resident (call .to_host() for numpy arrays)."""
    if a.cols != b.rows:
        raise ValueError("non-conformable operands")
    h = C.c_void_p()
    check(lib().g4s_spgemm_device(a.handle, b.handle, C.byref(h), _stream_ptr(stream)))
    return CSR._from_handle(h)


def OuterSpGEMM(a, b, stream=None):
    """OuterSpGEMM (mm/inc/outer_mult.h:271-542), the expand - sort - compress "join": same C as HashSpGEMM, values summed
    in the reference's sequential order for every row."""
    h = C.c_void_p()
    check(lib().g4s_spgemm_esc_device(a.handle, b.handle, C.byref(h), _stream_ptr(stream)))
    return CSR._from_handle(h)


def HeapSpGEMM(a, b, stream=None):
    """HeapSpGEMM (mm/inc/heap_mult.h:47-223): k-way heap merge of the sorted rows of b; same C as HashSpGEMM."""
    if a.cols != b.rows:
        raise ValueError("non-conformable operands")
    h = C.c_void_p()
    check(lib().g4s_spgemm_heap_device(a.handle, b.handle, C.byref(h), _stream_ptr(stream)))
    return CSR._from_handle(h)


def mkl(A, B, timing=None):
    """mkl(A, B, C, timing) (mm/inc/mkl_mult.h:113-117 -> :40-110): host CSR in, host CSR out, phases in timing."""
    a, b = A.to_host(), B.to_host()
    if a.cols != b.rows:
        raise ValueError("non-conformable operands")
    t = timing if timing is not None else Timings()
    cnnz = C.c_int()
    rp, ci, va = i32p(), i32p(), f64p()
    check(lib().g4s_mkl(_ip(a.rowptr), _ip(a.colids), _dp(a.values), _ip(b.rowptr), _ip(b.colids), _dp(b.values),
                        C.byref(rp), C.byref(ci), C.byref(va), C.c_int(a.rows), C.c_int(a.cols), C.c_int(b.cols),
                        C.byref(cnnz), C.byref(t)))
    return CSR(a.rows, b.cols, _take(rp, a.rows + 1, np.int32), _take(ci, cnnz.value, np.int32),
               _take(va, cnnz.value, np.float64))


def bsr_from_citcoms_nodes(node_map, eqn_k1, eqn_k2, eqn_k3):
    """CitcomS half-stored node format (Node_map, Eqn_k1..3; citcoms/lib/Construct_arrays.c:264-312) -> full BSR 3x3:
    returns (browptr, bcolids, blocks[nnzb, 3, 3]).  float32 coefficients (the reference's higher_precision) or float64."""
    node_map = _i32(node_map)
    nno = len(node_map) // 42
    dt = np.float32 if np.asarray(eqn_k1).dtype == np.float32 else np.float64
    ks = [np.ascontiguousarray(k, dtype=dt) for k in (eqn_k1, eqn_k2, eqn_k3)]
    if len(node_map) != nno * 42 or any(len(k) != nno * 42 for k in ks):
        raise ValueError("node_map and Eqn_k1..3 must have nno*42 entries")
    nnzb = C.c_int()
    rp, ci, va = i32p(), i32p(), f64p()
    check(lib().g4s_bsr_from_citcoms_nodes(C.c_int(nno), _ip(node_map), ks[0].ctypes.data_as(C.c_void_p),
                                           ks[1].ctypes.data_as(C.c_void_p), ks[2].ctypes.data_as(C.c_void_p),
                                           C.c_int(ks[0].itemsize), C.byref(nnzb), C.byref(rp), C.byref(ci), C.byref(va)))
    return (_take(rp, nno + 1, np.int32), _take(ci, nnzb.value, np.int32),
            _take(va, 9 * nnzb.value, np.float64).reshape(-1, 3, 3))
