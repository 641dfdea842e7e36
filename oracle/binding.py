"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle.

`Oracle`  : oracle/liboracle.so, the plain-C restatement (oracle_*.c).
`Ref`     : oracle/_ref/libg4s_ref.so, the reference's own mm/inc headers compiled unmodified
            (oracle/ref_shim.cpp), and oracle/_ref/libmv_ref.so, the reference's own mv/mv.c linked
            against scipy's OpenBLAS in place of MKL.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package g4s_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)
_longp = C.POINTER(C.c_long)


def build(quiet=True):
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref/*.so."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _dp(a):
    return a.ctypes.data_as(_f64p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _take(free, ptr, n, dtype):
    """Copy n items out of a malloc'd C array and free it."""
    if n:
        arr = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    else:
        arr = np.zeros(0, dtype=dtype)
    free(ptr)
    return arr


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_intprod.restype = C.c_longlong
        L.oracle_laplacian3d27_nnz.restype = C.c_longlong
        L.oracle_hash_spgemm_omp.restype = C.c_double
        L.oracle_free.argtypes = [C.c_void_p]

    def _free(self, p):
        self.lib.oracle_free(C.cast(p, C.c_void_p))

    def omp_max_threads(self):
        return int(self.lib.oracle_omp_max_threads())

    # ---- mv -------------------------------------------------------------------------------------
    def spmv_csr(self, rowptr, colids, values, x, omp=False):
        rowptr, colids, values, x = _i32(rowptr), _i32(colids), _f64(values), _f64(x)
        rows = len(rowptr) - 1
        y = np.empty(rows, dtype=np.float64)
        fn = self.lib.oracle_spmv_csr_omp if omp else self.lib.oracle_spmv_csr
        fn(C.c_int(rows), _ip(rowptr), _ip(colids), _dp(values), _dp(x), _dp(y))
        return y

    def spmv_csr_abs(self, rowptr, colids, values, x):
        rowptr, colids, values, x = _i32(rowptr), _i32(colids), _f64(values), _f64(x)
        rows = len(rowptr) - 1
        y = np.empty(rows, dtype=np.float64)
        self.lib.oracle_spmv_csr_abs(C.c_int(rows), _ip(rowptr), _ip(colids), _dp(values), _dp(x), _dp(y))
        return y

    def gen_laplacian3d27(self, n, row0=0, row1=None):
        """Rows [row0,row1) of the 3-D 27-point Laplacian as (rows, cols, rowptr, colids, values)."""
        row1 = n ** 3 if row1 is None else row1
        nnz = int(self.lib.oracle_laplacian3d27_nnz(C.c_int(n), C.c_longlong(row0), C.c_longlong(row1)))
        rowptr = np.empty(row1 - row0 + 1, dtype=np.int32)
        colids = np.empty(nnz, dtype=np.int32)
        values = np.empty(nnz, dtype=np.float64)
        self.lib.oracle_gen_laplacian3d27(C.c_int(n), C.c_longlong(row0), C.c_longlong(row1), _ip(rowptr),
                                          _ip(colids), _dp(values))
        return (row1 - row0, n ** 3, rowptr, colids, values)

    def ebe_matvec(self, elt_k, elem_dofs, u, neq, ends=8, dims=3):
        """CitcomS element-by-element operator (gather at Element_calculations.c:453-471), flattened dof map."""
        elt_k, u = _f64(elt_k), _f64(u)
        elem_dofs = _i32(elem_dofs)
        nel = elem_dofs.shape[0]
        Au = np.zeros(neq, dtype=np.float64)
        self.lib.oracle_ebe_matvec(C.c_int(nel), C.c_int(ends), C.c_int(dims), _dp(elt_k), _ip(elem_dofs), _dp(u), _dp(Au))
        return Au

    def citcoms_mesh(self, nox, noy, noz, elt_k):
        """CitcomS node format of a structured hex mesh: (ien [nel,8] 1-based, node_map [nno*42], k1, k2, k3 float32
        [nno*42]) from per-element 24x24 matrices, restating construct_ien / construct_node_maps / construct_node_ks."""
        nno, nel = nox * noy * noz, (nox - 1) * (noy - 1) * (noz - 1)
        ien = np.zeros((nel, 8), dtype=np.int32)
        assert self.lib.oracle_citcoms_ien(C.c_int(nox), C.c_int(noy), C.c_int(noz), _ip(ien)) == nel
        node_map = np.zeros(nno * 42, dtype=np.int32)
        self.lib.oracle_citcoms_node_maps(C.c_int(nox), C.c_int(noy), C.c_int(noz), _ip(node_map))
        elt_k = _f64(elt_k)
        ks = [np.zeros(nno * 42, dtype=np.float32) for _ in range(3)]
        fp = C.POINTER(C.c_float)
        rc = self.lib.oracle_citcoms_node_ks(C.c_int(nel), C.c_int(nno), _ip(ien), _dp(elt_k), _ip(node_map),
                                             ks[0].ctypes.data_as(fp), ks[1].ctypes.data_as(fp), ks[2].ctypes.data_as(fp))
        assert rc == 0, "construct_node_ks: slot not found"
        return ien, node_map, ks[0], ks[1], ks[2]

    def citcoms_n_assemble_del2_u(self, node_map, k1, k2, k3, u):
        """n_assemble_del2_u (citcoms/lib/Element_calculations.c:516-565): Au = K u on the half-stored node format."""
        nno = len(node_map) // 42
        fp = C.POINTER(C.c_float)
        uu = np.zeros(3 * nno + 1, dtype=np.float64)
        uu[:3 * nno] = u
        Au = np.zeros(3 * nno + 1, dtype=np.float64)
        self.lib.oracle_citcoms_n_assemble_del2_u(C.c_int(nno), _ip(_i32(node_map)), k1.ctypes.data_as(fp),
                                                  k2.ctypes.data_as(fp), k3.ctypes.data_as(fp), _dp(uu), _dp(Au))
        return Au[:3 * nno]

    def dense_mv(self, name, A, B):
        """name in dgemv|dsymv|dtrmv|dspmv; returns (B_after, C) like mv/mv.c's (A,B,C,dim) calls."""
        A = _f64(A).reshape(-1)
        B = _f64(B).copy()
        dim = len(B)
        Cv = np.zeros(dim, dtype=np.float64)
        getattr(self.lib, "oracle_" + name)(_dp(A), _dp(B), _dp(Cv), C.c_int(dim))
        return B, Cv

    def opt_matmul(self, xx, w):
        """res = xx w as the engine computes it (deepmd/source/op/opt_matmul.cc:47-53 in graph.h:21-32)."""
        xx, w = _f64(xx), _f64(w)
        M, N = xx.shape
        K = w.shape[1]
        res = np.empty((M, K), dtype=np.float64)
        self.lib.oracle_opt_matmul(C.c_int(M), C.c_int(N), C.c_int(K), _dp(xx), _dp(w), _dp(res))
        return res

    def bsr_spmm(self, browptr, bcolids, bvalues, bs, Bd):
        browptr, bcolids, bvalues, Bd = _i32(browptr), _i32(bcolids), _f64(bvalues), _f64(Bd)
        mb = len(browptr) - 1
        ncol = Bd.shape[1]
        out = np.empty((mb * bs, ncol), dtype=np.float64)
        self.lib.oracle_bsr_spmm(C.c_int(mb), C.c_int(bs), _ip(browptr), _ip(bcolids), _dp(bvalues), C.c_int(ncol),
                                 _dp(Bd), _dp(out))
        return out

    # ---- mm -------------------------------------------------------------------------------------
    def intprod(self, arpt, acol, brpt):
        arpt, acol, brpt = _i32(arpt), _i32(acol), _i32(brpt)
        rows = len(arpt) - 1
        row_nz = np.zeros(rows, dtype=np.int32)
        total = self.lib.oracle_intprod(_ip(arpt), _ip(acol), _ip(brpt), C.c_int(rows), _ip(row_nz))
        return int(total), row_nz

    def rows_offset(self, row_nz, total, parts):
        row_nz = _i32(row_nz)
        out = np.zeros(parts + 1, dtype=np.int32)
        self.lib.oracle_rows_offset(_ip(row_nz), C.c_int(len(row_nz)), C.c_longlong(total), C.c_int(parts), _ip(out))
        return out

    def bin_id(self, row_nz, cols, min_ht=8):
        row_nz = _i32(row_nz)
        out = np.zeros(len(row_nz), dtype=np.int8)
        self.lib.oracle_bin_id(_ip(row_nz), C.c_int(len(row_nz)), C.c_int(cols), C.c_int(min_ht),
                               out.ctypes.data_as(C.POINTER(C.c_byte)))
        return out

    def hash_spgemm(self, A, B, threads=0):
        """A, B = (rows, cols, rowptr, colids, values). Returns (rowptr, colids, values[, seconds])."""
        M, K, arpt, acol, aval = A[0], A[1], _i32(A[2]), _i32(A[3]), _f64(A[4])
        N, brpt, bcol, bval = B[1], _i32(B[2]), _i32(B[3]), _f64(B[4])
        cnnz = C.c_int(0)
        crpt, ccol, cval = _i32p(), _i32p(), _f64p()
        args = [C.c_int(M), C.c_int(K), C.c_int(N), _ip(arpt), _ip(acol), _dp(aval), _ip(brpt), _ip(bcol), _dp(bval),
                C.byref(cnnz), C.byref(crpt), C.byref(ccol), C.byref(cval)]
        secs = None
        if threads:
            secs = float(self.lib.oracle_hash_spgemm_omp(C.c_int(threads), *args))
        else:
            self.lib.oracle_hash_spgemm(*args)
        out = (_take(self._free, crpt, M + 1, np.int32), _take(self._free, ccol, cnnz.value, np.int32),
               _take(self._free, cval, cnnz.value, np.float64))
        return out + ((secs,) if threads else ())

    # ---- formats --------------------------------------------------------------------------------
    def mm_construct(self, path):
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        rp, ci, va = _i32p(), _i32p(), _f64p()
        err = C.create_string_buffer(256)
        st = self.lib.oracle_mm_construct(path.encode(), C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(rp),
                                          C.byref(ci), C.byref(va), err, C.c_int(256))
        if st != 0:
            raise RuntimeError(err.value.decode())
        return (rows.value, cols.value, _take(self._free, rp, rows.value + 1, np.int32),
                _take(self._free, ci, nnz.value, np.int32), _take(self._free, va, nnz.value, np.float64))

    def csr_from_graph(self, n, start, end, w):
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        w = _f64(w)
        nnz = C.c_int()
        rp, ci, va = _i32p(), _i32p(), _f64p()
        self.lib.oracle_csr_from_graph(C.c_long(len(start)), C.c_long(n), start.ctypes.data_as(_longp),
                                       end.ctypes.data_as(_longp), _dp(w), C.byref(nnz), C.byref(rp), C.byref(ci),
                                       C.byref(va))
        return (n, n, _take(self._free, rp, n + 1, np.int32), _take(self._free, ci, nnz.value, np.int32),
                _take(self._free, va, nnz.value, np.float64))

    def csr_submatrix(self, A, M_, N_, M_start=0, N_start=0):
        rows, cols, rpt, col, val = A[0], A[1], _i32(A[2]), _i32(A[3]), _f64(A[4])
        nnz = C.c_int()
        rp, ci, va = _i32p(), _i32p(), _f64p()
        st = self.lib.oracle_csr_submatrix(C.c_int(rows), C.c_int(cols), _ip(rpt), _ip(col), _dp(val), C.c_int(M_),
                                           C.c_int(N_), C.c_int(M_start), C.c_int(N_start), C.byref(nnz),
                                           C.byref(rp), C.byref(ci), C.byref(va))
        if st != 0:
            raise ValueError("matrix subsect error")
        return (M_, N_, _take(self._free, rp, M_ + 1, np.int32), _take(self._free, ci, nnz.value, np.int32),
                _take(self._free, va, nnz.value, np.float64))


class Ref:
    """The reference's own code (oracle/_ref).  `available` is False when the .so files were never built."""

    def __init__(self):
        p1 = os.path.join(HERE, "_ref", "libg4s_ref.so")
        p2 = os.path.join(HERE, "_ref", "libmv_ref.so")
        self.available = os.path.exists(p1)
        self.mv_available = os.path.exists(p2)
        if self.available:
            self.lib = L = C.CDLL(p1)
            L.ref_hash_spgemm.restype = C.c_double
            L.ref_get_flop.restype = C.c_longlong
            L.ref_bin.restype = C.c_longlong
            L.ref_last_error.restype = C.c_char_p
            L.ref_free.argtypes = [C.c_void_p]
        if self.mv_available:
            try:
                self.mv = C.CDLL(p2)
            except OSError:
                self.mv_available = False

    def _free(self, p):
        self.lib.ref_free(C.cast(p, C.c_void_p))

    def omp_max_threads(self):
        return int(self.lib.ref_omp_max_threads())

    def set_threads(self, n):
        self.lib.ref_omp_set_threads(C.c_int(n))

    def _ab(self, A, B):
        M, K, arpt, acol, aval = A[0], A[1], _i32(A[2]), _i32(A[3]), _f64(A[4])
        N, brpt, bcol, bval = B[1], _i32(B[2]), _i32(B[3]), _f64(B[4])
        keep = (arpt, acol, aval, brpt, bcol, bval)
        args = [C.c_int(M), C.c_int(K), C.c_int(N), C.c_int(len(acol)), _ip(arpt), _ip(acol), _dp(aval),
                C.c_int(len(bcol)), _ip(brpt), _ip(bcol), _dp(bval)]
        return M, args, keep

    def hash_spgemm(self, A, B, variant=0, want_output=True):
        """Reference HashSpGEMM; variant 0 = <false,true> (the parity oracle). Returns (rpt, col, val, seconds)."""
        M, args, keep = self._ab(A, B)
        cnnz = C.c_int(0)
        crpt, ccol, cval = _i32p(), _i32p(), _f64p()
        if want_output:
            secs = self.lib.ref_hash_spgemm(C.c_int(variant), *args, C.byref(cnnz), C.byref(crpt), C.byref(ccol),
                                            C.byref(cval))
        else:
            secs = self.lib.ref_hash_spgemm(C.c_int(variant), *args, C.byref(cnnz), None, None, None)
        if secs < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        if not want_output:
            return None, None, cnnz.value, float(secs)
        return (_take(self._free, crpt, M + 1, np.int32), _take(self._free, ccol, cnnz.value, np.int32),
                _take(self._free, cval, cnnz.value, np.float64), float(secs))

    def heap_spgemm(self, A, B):
        M, args, keep = self._ab(A, B)
        cnnz = C.c_int(0)
        crpt, ccol, cval = _i32p(), _i32p(), _f64p()
        st = self.lib.ref_heap_spgemm(*args, C.byref(cnnz), C.byref(crpt), C.byref(ccol), C.byref(cval))
        if st != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return (_take(self._free, crpt, M + 1, np.int32), _take(self._free, ccol, cnnz.value, np.int32),
                _take(self._free, cval, cnnz.value, np.float64))

    def get_flop(self, A, B):
        M, args, keep = self._ab(A, B)
        return int(self.lib.ref_get_flop(*args))

    def bin(self, threads, A, B):
        rows, arpt, acol, brpt = A[0], _i32(A[2]), _i32(A[3]), _i32(B[2])
        row_nz = np.zeros(rows, dtype=np.int32)
        offs = np.zeros(threads + 1, dtype=np.int32)
        bid = np.zeros(rows, dtype=np.int8)
        total = self.lib.ref_bin(C.c_int(threads), C.c_int(rows), C.c_int(B[1]), _ip(arpt), _ip(acol), _ip(brpt),
                                 _ip(row_nz), _ip(offs), bid.ctypes.data_as(C.POINTER(C.c_byte)))
        return int(total), row_nz, offs, bid

    def _csr_out(self, st, rows, cols, nnz, rp, ci, va):
        if st != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return (rows.value, cols.value, _take(self._free, rp, rows.value + 1, np.int32),
                _take(self._free, ci, nnz.value, np.int32), _take(self._free, va, nnz.value, np.float64))

    def csr_construct(self, path):
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        rp, ci, va = _i32p(), _i32p(), _f64p()
        st = self.lib.ref_csr_construct(path.encode(), C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(rp),
                                        C.byref(ci), C.byref(va))
        return self._csr_out(st, rows, cols, nnz, rp, ci, va)

    def csr_from_graph(self, n, start, end, w):
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        w = _f64(w)
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        rp, ci, va = _i32p(), _i32p(), _f64p()
        st = self.lib.ref_csr_from_graph(C.c_long(len(start)), C.c_long(n), start.ctypes.data_as(_longp),
                                         end.ctypes.data_as(_longp), _dp(w), C.byref(rows), C.byref(cols),
                                         C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(va))
        return self._csr_out(st, rows, cols, nnz, rp, ci, va)

    def csr_equal(self, A, B):
        """CSR::operator== of the reference (structure exact, values within 1e-3)."""
        if A[0] != B[0] or A[1] != B[1] or len(A[3]) != len(B[3]):
            return False
        r1, c1, v1, r2, c2, v2 = _i32(A[2]), _i32(A[3]), _f64(A[4]), _i32(B[2]), _i32(B[3]), _f64(B[4])
        return bool(self.lib.ref_csr_equal(C.c_int(A[0]), C.c_int(A[1]), C.c_int(len(c1)), _ip(r1), _ip(c1), _dp(v1),
                                           _ip(r2), _ip(c2), _dp(v2)))

    def opt_matmul(self, xx, w):
        """res = xx w through the reference's own GraphProcess (deepmd/source/op/graph.h, compiled where it lies)."""
        xx, w = _f64(xx), _f64(w)
        M, N = xx.shape
        K = w.shape[1]
        res = np.empty((M, K), dtype=np.float64)
        self.lib.ref_opt_matmul(C.c_int(M), C.c_int(N), C.c_int(K), _dp(xx), _dp(w), _dp(res))
        return res

    def dense_mv(self, name, A, B):
        """mv/mv.c's own matrix_multiply_{dsymv,dtrmv,sspmv,dgemv}(A,B,C,dim) on OpenBLAS."""
        A = _f64(A).reshape(-1)
        B = _f64(B).copy()
        dim = len(B)
        Cv = np.zeros(dim, dtype=np.float64)
        getattr(self.mv, "matrix_multiply_" + name)(_dp(A), _dp(B), _dp(Cv), C.c_int(dim))
        return B, Cv
