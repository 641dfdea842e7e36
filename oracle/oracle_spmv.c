/* TEST INFRASTRUCTURE ONLY — see oracle.h for scope and pinning status.
 *
 * Matrix-vector side of the oracle.  The reference's mv/ driver (mv/mv.c) calls four MKL CBLAS level-2
 * routines on a dense buffer; MKL is third-party and absent, so each routine is restated here from its
 * published BLAS definition, with the exact argument choices of the reference call sites. */
#include "oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

int oracle_omp_max_threads(void) { return omp_get_max_threads(); }

/* y = A x on the reference's CSR container (mm/inc/CSR.h:22-100: 0-based rowptr[rows+1], colids, values),
 * alpha = 1, beta = 0 as every call in mv/mv.c:6-27.  One left-to-right sum per row, separate multiply
 * and add (built with -ffp-contract=off), so the result is a fixed function of the stored order. */
void oracle_spmv_csr(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                     double *y) {
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (long j = rowptr[i]; j < rowptr[i + 1]; ++j) s += values[j] * x[colids[j]];
        y[i] = s;
    }
}

/* Same product, rows split over OpenMP threads (each row still summed left to right, so the result is
 * bit-identical to oracle_spmv_csr).  This is the CPU baseline bench.py times. */
void oracle_spmv_csr_omp(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                         double *y) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (long j = rowptr[i]; j < rowptr[i + 1]; ++j) s += values[j] * x[colids[j]];
        y[i] = s;
    }
}

/* sum_j |a_ij| |x_j| : the scale against which tests state the 1e-12 relative tolerance. */
void oracle_spmv_csr_abs(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                         double *yabs) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (long j = rowptr[i]; j < rowptr[i + 1]; ++j) s += fabs(values[j]) * fabs(x[colids[j]]);
        yabs[i] = s;
    }
}

/* mv/mv.c:23-27  cblas_dgemv(ColMajor, NoTrans, dim, dim, 1.0, A, dim, B, 1, 0.0, C, 1).
 * Column-major element (i,j) = A[i + j*dim]; the driver filled A row-major (mv/mv.c:62), so this is
 * y = (file matrix)^T x. */
void oracle_dgemv(const double *A, const double *B, double *C, int dim) {
    for (int i = 0; i < dim; ++i) C[i] = 0.0;
    for (int j = 0; j < dim; ++j) {
        const double xj = B[j];
        const double *col = A + (size_t)j * dim;
        for (int i = 0; i < dim; ++i) C[i] += col[i] * xj;
    }
}

/* mv/mv.c:6-10  cblas_dsymv(ColMajor, Upper, dim, 1.0, A, dim, B, 1, 0.0, C, 1).
 * Only a(i,j), i <= j, = A[i + j*dim] is referenced; the strictly lower part is its mirror. */
void oracle_dsymv(const double *A, const double *B, double *C, int dim) {
    for (int i = 0; i < dim; ++i) C[i] = 0.0;
    for (int j = 0; j < dim; ++j) {
        const double *col = A + (size_t)j * dim;
        double t = 0.0;
        for (int i = 0; i < j; ++i) {
            C[i] += col[i] * B[j];
            t += col[i] * B[i];
        }
        C[j] += col[j] * B[j] + t;
    }
}

/* mv/mv.c:12-15  cblas_dtrmv(ColMajor, Upper, Trans, NonUnit, dim, A, dim, B, 1):  B <- U^T B in place,
 * U(i,j) = A[i + j*dim] for i <= j.  C is not touched. */
void oracle_dtrmv(const double *A, double *B, double *C, int dim) {
    (void)C;
    for (int j = dim - 1; j >= 0; --j) { /* descending j only reads B[i], i <= j: safe in place */
        const double *col = A + (size_t)j * dim;
        double t = 0.0;
        for (int i = 0; i <= j; ++i) t += col[i] * B[i];
        B[j] = t;
    }
}

/* mv/mv.c:17-21  cblas_dspmv(ColMajor, Upper, dim, 1.0, A, B, 1, 0.0, C, 1): the dense buffer is read AS IF
 * it were packed-upper storage: a(i,j), i <= j, = A[i + j(j+1)/2] (first dim(dim+1)/2 doubles). */
void oracle_dspmv(const double *A, const double *B, double *C, int dim) {
    for (int i = 0; i < dim; ++i) C[i] = 0.0;
    for (int j = 0; j < dim; ++j) {
        const double *col = A + (size_t)j * (j + 1) / 2;
        double t = 0.0;
        for (int i = 0; i < j; ++i) {
            C[i] += col[i] * B[j];
            t += col[i] * B[i];
        }
        C[j] += col[j] * B[j] + t;
    }
}

/* Block-CSR (bs x bs row-major blocks) times a row-major dense block of ncol columns.  There is no such
 * routine in the reference; this restates C = A B for BASELINE config 5 so the
 * tensor-core path has a checker.  PARITY UNPINNED. */
void oracle_bsr_spmm(int mb, int bs, const int *browptr, const int *bcolids, const double *bvalues, int ncol,
                     const double *B, double *C) {
#pragma omp parallel for schedule(static)
    for (int I = 0; I < mb; ++I) {
        double *crow = C + (size_t)I * bs * ncol;
        memset(crow, 0, sizeof(double) * bs * ncol);
        for (long p = browptr[I]; p < browptr[I + 1]; ++p) {
            const double *blk = bvalues + (size_t)p * bs * bs;
            const double *brow = B + (size_t)bcolids[p] * bs * ncol;
            for (int r = 0; r < bs; ++r)
                for (int s = 0; s < bs; ++s) {
                    const double a = blk[r * bs + s];
                    for (int c = 0; c < ncol; ++c) crow[r * ncol + c] += a * brow[s * ncol + c];
                }
        }
    }
}

/* Synthetic input of BASELINE config 2 on the host (rows [row0,row1) of the 3-D 27-point Laplacian on an n^3
 * grid, natural ordering, diag 26, off-diag -1, global column ids).  Written independently of the product's
 * device generator so that tests can cross-check the two, and used by bench.py's CPU legs to build their
 * bounded sample without touching the GPU. */
static long long o_cnt(int n, long long t) { return 1 + (t > 0) + (t < n - 1); }
long long oracle_laplacian3d27_nnz(int n, long long row0, long long row1) {
    long long total = 0;
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (long long row = row0; row < row1; ++row)
        total += o_cnt(n, row % n) * o_cnt(n, (row / n) % n) * o_cnt(n, row / ((long long)n * n));
    return total;
}
void oracle_gen_laplacian3d27(int n, long long row0, long long row1, int *rowptr, int *colids, double *values) {
    const long long nrows = row1 - row0;
    rowptr[0] = 0;
    for (long long r = 0; r < nrows; ++r) { /* sequential prefix: the sample is a few million rows */
        const long long row = row0 + r;
        rowptr[r + 1] = rowptr[r] +
                        (int)(o_cnt(n, row % n) * o_cnt(n, (row / n) % n) * o_cnt(n, row / ((long long)n * n)));
    }
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < nrows; ++r) {
        const long long row = row0 + r;
        const int i = (int)(row % n), j = (int)((row / n) % n), k = (int)(row / ((long long)n * n));
        long p = rowptr[r];
        for (int dk = -1; dk <= 1; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                    if (k + dk < 0 || k + dk >= n || j + dj < 0 || j + dj >= n || i + di < 0 || i + di >= n) continue;
                    const long long col = ((long long)(k + dk) * n + (j + dj)) * n + (i + di);
                    colids[p] = (int)col;
                    values[p] = col == row ? 26.0 : -1.0;
                    ++p;
                }
    }
}

/* The G4S engine loop, GraphProcess (deepmd/source/op/graph.h:21-32): for every vertex, gather() per neighbour, then
 * apply().  Restated sequentially; the reference's copy runs it under `omp parallel for schedule(dynamic,1)`. */
typedef void (*oracle_gather_fn)(int, int, const double **, const double *, double *);
typedef void (*oracle_apply_fn)(int, const double **, const double *, double *);
void oracle_graph_process(int num_nodes, int degree, const double **edge_weight, const double *states, double *result,
                          oracle_gather_fn gather, oracle_apply_fn apply) {
    for (int vi = 0; vi < num_nodes; ++vi) {
        for (int nb = 0; nb < degree; ++nb) gather(vi, nb, edge_weight, states, result);
        if (apply) apply(vi, edge_weight, states, result);
    }
}

/* CitcomS's gather callback (citcoms/lib/Element_calculations.c:453-471) summed over a whole mesh, i.e. what
 * e_assemble_del2_u computes through the engine (:475-510): for element e, node a, direction i,
 *     Au[dof(e,a,i)] += sum_b sum_j elt_k[e][ii + j] * u[dof(e,b,j)],   ii = (a*n+b)*dims - (dims*n+dims) + (i-1)*n
 * (1-based a, b, i; n = loc_mat_size = ends*dims), which is row 3(a-1)+(i-1), columns 3(b-1)+j of a row-major n x n
 * block.  The IEN / ID indirection is flattened into elem_dofs[e][3(a-1)+(i-1)].  PARITY UNPINNED: CitcomS is not
 * buildable here; the index arithmetic is restated literally and cross-checked against an assembled sparse matrix. */
void oracle_ebe_matvec(int nel, int ends, int dims, const double *elt_k, const int *elem_dofs, const double *u,
                       double *Au) {
    const int n = ends * dims;
    for (int e = 0; e < nel; ++e) {
        const double *k = elt_k + (size_t)e * n * n;
        const int *dof = elem_dofs + (size_t)e * n;
        for (int a = 1; a <= ends; ++a)
            for (int i = 1; i <= dims; ++i) {
                const int aa = dof[dims * (a - 1) + (i - 1)];
                for (int b = 1; b <= ends; ++b) {
                    const int ii = (a * n + b) * dims - (dims * n + dims) + (i - 1) * n;
                    double s = 0.0;
                    for (int j = 0; j < dims; ++j) s += k[ii + j] * u[dof[dims * (b - 1) + j]];
                    Au[aa] += s;
                }
            }
    }
}

/* OptMatmul: res[M,K] = xx[M,N] w[N,K], row-major, restating the gather of
 * deepmd/source/op/opt_matmul.cc:47-53 inside the engine loop of deepmd/source/op/graph.h:21-32 — one vertex per row of
 * xx, one gather per output column, "result[e*Col+a] = 0; for k: result[e*Col+a] += edgeWeight[e][k] * states[k*Col+a]"
 * (k ascending, one product and one addition per term; -ffp-contract=off keeps them separate).  The reference pins the
 * thread count to 8 with schedule(dynamic,1); the result does not depend on it.  Pinned against the reference's own
 * GraphProcess (oracle/_ref, ref_opt_matmul) in tests/test_opt_matmul.py. */
void oracle_opt_matmul(int M, int N, int K, const double *xx, const double *w, double *res) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int e = 0; e < M; ++e)
        for (int a = 0; a < K; ++a) {
            res[(size_t)e * K + a] = 0;
            for (int k = 0; k < N; ++k) res[(size_t)e * K + a] += xx[(size_t)e * N + k] * w[(size_t)k * K + a];
        }
}
