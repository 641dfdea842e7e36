// C++ convenience layer over the C ABI (include/g4s_b200.h) with the reference's own call shapes, for a
// maintainer who swaps the MKL-backed SpGEMM of mm/ for the GPU one.  Header-only; include it AFTER the
// reference's CSR.h / Timings.h / utility.h (it uses their CSR<int,double>, Timings and my_malloc so that the
// caller's my_free in CSR::make_empty (mm/inc/CSR.h:50-62) matches the allocation).
//
//   mkl(arpt, acol, aval, brpt, bcol, bval, &crpt, &ccol, &cval, M, K, N, &cnnz, timing)   mm/inc/mkl_mult.h:40-43
//   mkl(A, B, C, timing) / mkl(A, B, C)                                                      mm/inc/mkl_mult.h:113-124
//   HashSpGEMM<sortOutput>(a, b, c, multiplies, plus)                                        mm/inc/hash_mult.h:1103-1113
#ifndef G4S_B200_HPP
#define G4S_B200_HPP

#include <cstdio>
#include <cstdlib>
#include <functional>

#include "g4s_b200.h"

namespace g4s_detail {
inline void *alloc_int(size_t bytes, void *) { return my_malloc<int>(bytes / sizeof(int) + 1); }
inline void *alloc_double(size_t bytes, void *) { return my_malloc<double>(bytes / sizeof(double) + 1); }
inline void die(const char *what) {  // the reference asserts on every MKL status (mm/inc/mkl_mult.h:51-107)
    std::fprintf(stderr, "g4s_b200: %s failed: %s\n", what, g4s_last_error());
    std::abort();
}
static_assert(sizeof(Timings) == sizeof(g4s_timings), "Timings layout differs from g4s_timings");
}  // namespace g4s_detail

inline void mkl(int *arpt, int *acol, double *aval, int *brpt, int *bcol, double *bval, int **crpt_, int **ccol_,
                double **cval_, int M, int K, int N, int *cnnz_, Timings &timing) {
    if (g4s_mkl_alloc(arpt, acol, aval, brpt, bcol, bval, crpt_, ccol_, cval_, M, K, N, cnnz_,
                      reinterpret_cast<g4s_timings *>(&timing), g4s_detail::alloc_int, g4s_detail::alloc_double,
                      nullptr) != G4S_OK)
        g4s_detail::die("mkl");
}

inline void mkl(const CSR<int, double> &A, const CSR<int, double> &B, CSR<int, double> &C, Timings &timing) {
    C.rows = A.rows;
    C.cols = B.cols;
    mkl(A.rowptr, A.colids, A.values, B.rowptr, B.colids, B.values, &C.rowptr, &C.colids, &C.values, A.rows, A.cols,
        B.cols, &C.nnz, timing);
}

inline void mkl(const CSR<int, double> &A, const CSR<int, double> &B, CSR<int, double> &C) {
    Timings timing;
    mkl(A, B, C, timing);
}

// GPU twin of HashSpGEMM for the only instantiation the reference uses (CSR<int,double>, multiplies, plus);
// the output is always column-sorted (sortOutput = true costs nothing extra on the GPU).
inline void G4sHashSpGEMM(const CSR<int, double> &a, const CSR<int, double> &b, CSR<int, double> &c) {
    Timings t;
    c.zerobased = true;
    mkl(a, b, c, t);
}

#endif
