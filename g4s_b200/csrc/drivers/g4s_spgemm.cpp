// g4s_spgemm — the reference's mm/ SpGEMM driver (mm/src/mkl_spgemm.cpp:5-86) on top of libg4s_b200.so.
//
//   g4s_spgemm <A.mtx> [B.mtx]        (a name without '/' and '.mtx' is resolved like the reference does:
//                                      ../matrix/{ER,G500,suite_sparse/<name>}/<name>.mtx, mkl_spgemm.cpp:19-37)
//
// Same flow: construct A (and B, or B = A), trim to conformable shapes with the sub-matrix constructor
// (mkl_spgemm.cpp:42-58), count the intermediate products, one warm-up multiply, ten timed mkl() calls averaged,
// Timings::print(2 * products).  Every step goes through the C ABI; the arithmetic runs on the GPU.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "g4s_b200.h"

namespace {
struct Csr {
    int rows = 0, cols = 0, nnz = 0;
    int *rowptr = nullptr, *colids = nullptr;
    double *values = nullptr;
    void clear() {
        g4s_free(rowptr);
        g4s_free(colids);
        g4s_free(values);
        rowptr = colids = nullptr;
        values = nullptr;
        rows = cols = nnz = 0;
    }
};
std::string resolve(const std::string &name) {
    if (name.find('/') != std::string::npos || name.find(".mtx") != std::string::npos) return name;
    if (name.find("ER") != std::string::npos) return "../matrix/ER/" + name + ".mtx";
    if (name.find("G500") != std::string::npos) return "../matrix/G500/" + name + ".mtx";
    return "../matrix/suite_sparse/" + name + "/" + name + ".mtx";
}
void die(const char *what) {
    std::fprintf(stderr, "g4s_spgemm: %s failed: %s\n", what, g4s_last_error());
    std::exit(1);
}
// construct() through the binary cache `<path>.g4scsr` (G4S_NO_CACHE=1: parse the text file every time, as the reference)
void construct(Csr &m, const std::string &path) {
    int hit = 0;
    const int rc = std::getenv("G4S_NO_CACHE")
                       ? g4s_csr_read_matrix_market(path.c_str(), &m.rows, &m.cols, &m.nnz, &m.rowptr, &m.colids, &m.values)
                       : g4s_csr_read_cached(path.c_str(), &m.rows, &m.cols, &m.nnz, &m.rowptr, &m.colids, &m.values, &hit);
    if (rc != G4S_OK) die(("construct(" + path + ")").c_str());
    if (hit) std::printf("(loaded from %s.g4scsr)\n", path.c_str());
}
void trim(Csr &m, int M_, int N_) {
    Csr t;
    t.rows = M_;
    t.cols = N_;
    if (g4s_csr_submatrix(m.rows, m.cols, m.rowptr, m.colids, m.values, M_, N_, 0, 0, &t.nnz, &t.rowptr, &t.colids,
                          &t.values) != G4S_OK)
        die("submatrix");
    m.clear();
    m = t;
}
}  // namespace

int main(int argc, char **argv) {
    std::string mat1 = "can_24", mat2 = "can_24";
    if (argc == 2) mat1 = mat2 = argv[1];
    if (argc >= 3) {
        mat1 = argv[1];
        mat2 = argv[2];
    }
    std::printf("reading matrix A from %s\n", resolve(mat1).c_str());
    Csr A, B;
    construct(A, resolve(mat1));
    const bool same = mat1 == mat2;
    if (!same) {
        construct(B, resolve(mat2));
        if (A.cols < B.rows) trim(B, A.cols, B.cols);
        else if (A.cols > B.rows) trim(A, A.rows, B.rows);
    }
    const Csr &Bm = same ? A : B;
    const long long total_flop = compute_flop_host(A.rowptr, A.colids, Bm.rowptr, A.rows);

    g4s_timings timing, bench;
    g4s_timings_init(&timing);
    g4s_timings_init(&bench);
    int *crpt = nullptr, *ccol = nullptr, cnnz = 0;
    double *cval = nullptr;
    auto multiply = [&]() {
        if (g4s_mkl(A.rowptr, A.colids, A.values, Bm.rowptr, Bm.colids, Bm.values, &crpt, &ccol, &cval, A.rows, A.cols,
                    Bm.cols, &cnnz, &timing) != G4S_OK)
            die("mkl");
    };
    auto release = [&]() {
        g4s_free(crpt);
        g4s_free(ccol);
        g4s_free(cval);
    };
    multiply();  // warm-up (mkl_spgemm.cpp:67-69)
    release();
    const int iter = 10;
    for (int i = 0; i < iter; ++i) {
        multiply();
        g4s_timings_add(&bench, &timing);
        release();
    }
    g4s_timings_div(&bench, iter);
    std::printf("C: %d x %d, nnz %d\n", A.rows, Bm.cols, cnnz);
    g4s_timings_print(&bench, 2.0 * (double)total_flop);
    A.clear();
    B.clear();
    return 0;
}
