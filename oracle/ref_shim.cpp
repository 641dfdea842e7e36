// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Thin extern "C" wrapper around the UNMODIFIED reference headers under
// /root/reference/mm/inc (compiled where they lie by oracle/Makefile into
// oracle/_ref/libg4s_ref.so).  It lets tests/, bench.py's cpu_baseline /
// --impl reference leg and __graft_entry__.smoke() call the reference's own
//   HashSpGEMM<false,true>   (mm/inc/hash_mult.h:1028-1057, :1109-1113)
//   HashSpGEMM<true,true>    (AVX2 chunk probing, mm/inc/hash_mult.h:305-398, :822-921)
//   HeapSpGEMM               (mm/inc/heap_mult.h:47-223)
//   CSR::construct           (mm/inc/CSR.h:485-669)
//   CSR(graph&)              (mm/inc/CSR.h:255-329)
//   BIN::set_max_bin         (mm/inc/BIN.h:77-186)
//   get_flop                 (mm/inc/hash_mult.h:45-62)
// through ctypes.  No reference source is copied here.
#include "all.h"

#include <chrono>
#include <cstring>
#include <functional>
#include <string>

namespace {
template <class T>
T *dup_out(const T *src, size_t n) {
    T *p = (T *)malloc(sizeof(T) * (n ? n : 1));
    if (n) memcpy(p, src, sizeof(T) * n);
    return p;
}
void export_csr(CSR<int, double> &c, int *rows, int *cols, int *nnz, int **rpt, int **col, double **val) {
    *rows = c.rows;
    *cols = c.cols;
    *nnz = c.nnz;
    *rpt = dup_out(c.rowptr, (size_t)c.rows + 1);
    *col = dup_out(c.colids, (size_t)c.nnz);
    *val = dup_out(c.values, (size_t)c.nnz);
}
std::string g_err;

// Reference defect guard: BIN::set_rows_offset (BIN.h:108-117) cuts at ceil(total/T)*(tid+1), which exceeds
// the total work when total < T*(T-1); lower_bound then returns rows+1 and a thread walks past the last row
// (segfault on e.g. the 3x3 tridiagonal case with 8 threads).  The wrapper lowers the OpenMP thread count
// for such tiny inputs instead of touching the reference; results do not depend on the thread count.
struct ThreadClamp {
    int saved;
    explicit ThreadClamp(long long total) : saved(omp_get_max_threads()) {
        int t = saved;
        while (t > 1 && (long long)t * (t - 1) > total) --t;
        omp_set_num_threads(t);
    }
    ~ThreadClamp() { omp_set_num_threads(saved); }
};
}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
void ref_free(void *p) { free(p); }
int ref_omp_max_threads() { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

// variant: 0 = HashSpGEMM<false,true> (scalar probing, sorted; the oracle),
//          1 = HashSpGEMM<true,true>  (vector probing, sorted),
//          2 = HashSpGEMM<false,false> (scalar probing, hash-table order)
// Returns wall seconds of the multiply itself (steady_clock), <0 on error.
double ref_hash_spgemm(int variant, int M, int K, int N, int annz, int *arpt, int *acol, double *aval, int bnnz,
                       int *brpt, int *bcol, double *bval, int *cnnz, int **crpt, int **ccol, double **cval) {
    try {
        CSR<int, double> A(arpt, acol, aval, M, K, annz, 0);
        CSR<int, double> B(brpt, bcol, bval, K, N, bnnz, 0);
        CSR<int, double> C;
        ThreadClamp clamp(get_flop(A, B));
        auto t0 = std::chrono::steady_clock::now();
        if (variant == 0)
            HashSpGEMM<false, true>(A, B, C, std::multiplies<double>(), std::plus<double>());
        else if (variant == 1)
            HashSpGEMM<true, true>(A, B, C, std::multiplies<double>(), std::plus<double>());
        else
            HashSpGEMM<false, false>(A, B, C, std::multiplies<double>(), std::plus<double>());
        auto t1 = std::chrono::steady_clock::now();
        int r, c;
        if (crpt) export_csr(C, &r, &c, cnnz, crpt, ccol, cval);
        else *cnnz = C.nnz;
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (std::exception &e) {
        g_err = e.what();
        return -1.0;
    }
}

// Same product through the reference's heap (k-way merge on CSC) variant, used only to
// cross-check the hash oracle.  HeapSpGEMM takes CSC inputs and yields CSC.
int ref_heap_spgemm(int M, int K, int N, int annz, int *arpt, int *acol, double *aval, int bnnz, int *brpt,
                    int *bcol, double *bval, int *cnnz, int **crpt, int **ccol, double **cval) {
    try {
        CSR<int, double> A(arpt, acol, aval, M, K, annz, 0);
        CSR<int, double> B(brpt, bcol, bval, K, N, bnnz, 0);
        CSC<int, double> Ac, Bc, Cc;
        convert(Ac, A);
        convert(Bc, B);
        HeapSpGEMM(Ac, Bc, Cc, std::multiplies<double>(), std::plus<double>());
        CSR<int, double> C;
        convert(C, Cc);
        int r, c;
        export_csr(C, &r, &c, cnnz, crpt, ccol, cval);
        return 0;
    } catch (std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

long long ref_get_flop(int M, int K, int N, int annz, int *arpt, int *acol, double *aval, int bnnz, int *brpt,
                       int *bcol, double *bval) {
    CSR<int, double> A(arpt, acol, aval, M, K, annz, 0);
    CSR<int, double> B(brpt, bcol, bval, K, N, bnnz, 0);
    return get_flop(A, B);
}

// BIN::set_max_bin with `threads` OpenMP threads: per-row intermediate products (row_nz),
// work-balanced row cuts (rows_offset[threads+1]) and hash-table size class (bin_id).
long long ref_bin(int threads, int rows, int cols, int *arpt, int *acol, int *brpt, int *row_nz, int *rows_offset,
                  signed char *bin_id) {
    int saved = omp_get_max_threads();
    omp_set_num_threads(threads);
    long long total;
    {
        BIN<int, double> bin(rows, 8);
        bin.set_max_bin(arpt, acol, brpt, rows, cols);
        memcpy(row_nz, bin.row_nz, sizeof(int) * rows);
        memcpy(rows_offset, bin.rows_offset, sizeof(int) * (threads + 1));
        for (int i = 0; i < rows; ++i) bin_id[i] = bin.bin_id[i];
        total = bin.total_intprod;
        // ~BIN frees one table per thread; they must exist first.
        bin.create_local_hash_table(cols);
    }
    omp_set_num_threads(saved);
    return total;
}

int ref_csr_construct(const char *path, int *rows, int *cols, int *nnz, int **rpt, int **col, double **val) {
    try {
        CSR<int, double> A;
        A.construct(std::string(path));
        export_csr(A, rows, cols, nnz, rpt, col, val);
        return 0;
    } catch (std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// CSR(graph&): edges must arrive grouped by start vertex; duplicates are summed.
int ref_csr_from_graph(long m, long n, const long *start, const long *end, const double *w, int *rows, int *cols,
                       int *nnz, int **rpt, int **col, double **val) {
    try {
        graph G;  // ~graph() free()s the three arrays, so hand it malloc'd copies
        G.m = m;
        G.n = n;
        G.start = dup_out(start, (size_t)m);
        G.end = dup_out(end, (size_t)m);
        G.w = dup_out(w, (size_t)m);
        CSR<int, double> A(G);
        export_csr(A, rows, cols, nnz, rpt, col, val);
        return 0;
    } catch (std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// CSR::operator== (structure exact, values within EPSILON=1e-3; CSR.h:343-408)
int ref_csr_equal(int rows, int cols, int nnz, int *rpt1, int *col1, double *val1, int *rpt2, int *col2,
                  double *val2) {
    CSR<int, double> A(rpt1, col1, val1, rows, cols, nnz, 0);
    CSR<int, double> B(rpt2, col2, val2, rows, cols, nnz, 0);
    return (A == B) ? 1 : 0;
}

}  // extern "C"

// ---- OptMatmul through the reference's own engine loop ------------------------------------------------------------------
// deepmd/source/op/graph.h is included unmodified (GraphProcess, struct Graph); the TensorFlow op around it
// (opt_matmul.cc) cannot be built here, so its gather lambda (opt_matmul.cc:47-53) is repeated below word for word except
// for the captured Nsize, which is a global in the reference too.
#include <omp.h>
#include <stdio.h>

#include <vector>
namespace deepmd_ref {
#include G4S_REF_DEEPMD_GRAPH_H  // -D from oracle/Makefile: $(REF)/deepmd/source/op/graph.h (mm/inc has a graph.h of its own)
int Nsize;
}  // namespace deepmd_ref
extern "C" int ref_opt_matmul(int M, int N, int K, const double *xx, const double *w, double *result) {
    using namespace deepmd_ref;
    Graph graph;
    Nsize = N;
    graph.states = w;
    graph.numNodes = M;
    graph.degree = K;
    std::vector<const double *> A((size_t)M);
    for (int i = 0; i < M; i++) A[i] = xx + (size_t)i * N;
    graph.edgeWeight = A.data();
    GraphProcess(&graph, result,
                 [&](int e, int a, struct Graph *graph, double *result) {
                     int Col = getNeighbors(graph, e);
                     result[e * Col + a] = 0;
                     for (int k = 0; k < Nsize; k++) {
                         result[e * Col + a] += graph->edgeWeight[e][k] * graph->states[k * Col + a];
                     }
                 },
                 [&](int e, struct Graph *graph, double *Au) {});
    return 0;
}
