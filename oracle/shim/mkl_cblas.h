/* Stand-in for Intel MKL's <mkl_cblas.h> (not vendored in the reference, not
 * installed here).  Maps the four CBLAS level-2 calls of mv/mv.c:9,14,20,26 onto the
 * LP64 OpenBLAS that scipy bundles (symbols carry a scipy_ prefix).
 * TEST INFRASTRUCTURE ONLY: used to build oracle/_ref/libmv_ref.so. */
#ifndef G4S_ORACLE_MKL_CBLAS_SHIM_H
#define G4S_ORACLE_MKL_CBLAS_SHIM_H
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
void scipy_cblas_dgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE, int m, int n, double alpha, const double *a, int lda,
                       const double *x, int incx, double beta, double *y, int incy);
void scipy_cblas_dsymv(CBLAS_LAYOUT, CBLAS_UPLO, int n, double alpha, const double *a, int lda,
                       const double *x, int incx, double beta, double *y, int incy);
void scipy_cblas_dtrmv(CBLAS_LAYOUT, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int n, const double *a, int lda,
                       double *x, int incx);
void scipy_cblas_dspmv(CBLAS_LAYOUT, CBLAS_UPLO, int n, double alpha, const double *ap,
                       const double *x, int incx, double beta, double *y, int incy);
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_dsymv scipy_cblas_dsymv
#define cblas_dtrmv scipy_cblas_dtrmv
#define cblas_dspmv scipy_cblas_dspmv
#endif
