/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's mv/ and mm/ algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so; the product (g4s_b200/) never does and has no CPU fallback.
 *
 * Pinning status:
 *  - SpGEMM / BIN / loaders: PINNED against the reference's own code, compiled unmodified into
 *    oracle/_ref/libg4s_ref.so (tests/test_oracle.py compares every function below with it, and
 *    tests/golden/ holds outputs generated from it by tests/golden/make_golden.py).
 *  - dense mv (dgemv/dsymv/dtrmv/dspmv): the reference's arithmetic is Intel MKL CBLAS, which is neither
 *    vendored nor installed: PARITY UNPINNED against MKL.  The restatement is checked against the
 *    reference's own mv/mv.c call sites linked to OpenBLAS (oracle/_ref/libmv_ref.so) instead.
 *  - CSR SpMV: the reference has no sparse matrix-vector product; oracle_spmv_csr is a restatement of
 *    "y = A x, alpha = 1, beta = 0" (mv/mv.c:23-27) on the CSR container of mm/inc/CSR.h:22-100.
 *    PARITY UNPINNED by any reference test; cross-checked against the dense path at small dim.
 */
#ifndef G4S_ORACLE_H
#define G4S_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- oracle_spmv.c ---- */
void oracle_spmv_csr(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                     double *y);
void oracle_spmv_csr_omp(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                         double *y);
void oracle_spmv_csr_abs(int rows, const int *rowptr, const int *colids, const double *values, const double *x,
                         double *yabs);
void oracle_dgemv(const double *A, const double *B, double *C, int dim);
void oracle_dsymv(const double *A, const double *B, double *C, int dim);
void oracle_dtrmv(const double *A, double *B, double *C, int dim);
void oracle_dspmv(const double *A, const double *B, double *C, int dim);
int oracle_omp_max_threads(void);
long long oracle_laplacian3d27_nnz(int n, long long row0, long long row1);
void oracle_gen_laplacian3d27(int n, long long row0, long long row1, int *rowptr, int *colids, double *values);
void oracle_opt_matmul(int M, int N, int K, const double *xx, const double *w, double *res);
void oracle_bsr_spmm(int mb, int bs, const int *browptr, const int *bcolids, const double *bvalues, int ncol,
                     const double *B, double *C);

/* ---- oracle_citcoms.c: CitcomS node-format operator, PARITY UNPINNED ---- */
int oracle_citcoms_ien(int nox, int noy, int noz, int *ien);
void oracle_citcoms_node_maps(int nox, int noy, int noz, int *node_map);
int oracle_citcoms_node_ks(int nel, int nno, const int *ien, const double *elt_k, const int *node_map, float *k1,
                           float *k2, float *k3);
void oracle_citcoms_n_assemble_del2_u(int nno, const int *node_map, const float *k1, const float *k2, const float *k3,
                                      double *u, double *Au);

/* ---- oracle_spgemm.c ---- */
long long oracle_intprod(const int *arpt, const int *acol, const int *brpt, int rows, int *row_nz);
void oracle_rows_offset(const int *row_nz, int rows, long long total_intprod, int parts, int *rows_offset);
void oracle_bin_id(const int *row_nz, int rows, int cols, int min_ht, signed char *bin_id);
void oracle_hash_symbolic(const int *arpt, const int *acol, const int *brpt, const int *bcol, int rows, int cols,
                          int *crpt, int *cnnz);
void oracle_hash_numeric(const int *arpt, const int *acol, const double *aval, const int *brpt, const int *bcol,
                         const double *bval, int rows, int cols, const int *crpt, int *ccol, double *cval,
                         int sort_output);
int oracle_hash_spgemm(int M, int K, int N, const int *arpt, const int *acol, const double *aval, const int *brpt,
                       const int *bcol, const double *bval, int *cnnz, int **crpt, int **ccol, double **cval);
double oracle_hash_spgemm_omp(int threads, int M, int K, int N, const int *arpt, const int *acol,
                              const double *aval, const int *brpt, const int *bcol, const double *bval, int *cnnz,
                              int **crpt, int **ccol, double **cval);
void oracle_free(void *p);

/* ---- oracle_formats.c ---- */
int oracle_mm_construct(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids,
                        double **values, char *err, int errlen);
int oracle_csr_from_graph(long m, long n, const long *start, const long *end, const double *w, int *nnz,
                          int **rowptr, int **colids, double **values);
int oracle_csr_submatrix(int rows, int cols, const int *rowptr, const int *colids, const double *values, int M_,
                         int N_, int M_start, int N_start, int *nnz, int **orpt, int **ocol, double **oval);

#ifdef __cplusplus
}
#endif
#endif
