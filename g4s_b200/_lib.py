"""ctypes loader for g4s_b200/libg4s_b200.so (the C ABI declared in include/g4s_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libg4s_b200.so")

i32p = C.POINTER(C.c_int)
f64p = C.POINTER(C.c_double)
longp = C.POINTER(C.c_long)
i64p = C.POINTER(C.c_longlong)


class G4SError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("g4s_b200 error %d: %s" % (status, message))
        self.status = status


class Timings(C.Structure):
    """C twin of the reference's class Timings (mm/inc/Timings.h:4-22)."""
    _fields_ = [("measure_separate", C.c_ubyte), ("measure_total", C.c_ubyte), ("create", C.c_double),
                ("spmm", C.c_double), ("convert", C.c_double), ("order", C.c_double), ("export_csr", C.c_double),
                ("destroy", C.c_double), ("total", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C g4s_b200/csrc). g4s_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.g4s_last_error.restype = C.c_char_p
        L.g4s_version.restype = C.c_char_p
        L.g4s_kernel_launch_count.restype = C.c_longlong
        L.g4s_free.argtypes = [C.c_void_p]
        for name in ("g4s_laplacian2d_nnz", "g4s_laplacian3d27_nnz", "compute_flop_host"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_longlong
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise G4SError(status, lib().g4s_last_error().decode())
