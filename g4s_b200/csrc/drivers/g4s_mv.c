/* g4s_mv — the reference's mv/ driver (mv/mv.c:29-97) on top of libg4s_b200.so.
 *
 *   g4s_mv <matrix.mtx>
 *
 * Like the reference it reads a MatrixMarket coordinate file as a PATTERN (row, column pairs; any value columns are
 * ignored), fills a dense dim x dim row-major buffer with rand() at the listed positions (mv/mv.c:59-63; the rest is
 * zero here, the reference leaves it uninitialised), sets B = 1.0 (mv/mv.c:65-67) and times the four entry points
 * in the reference's order (dsymv, dtrmv, sspmv, dgemv; dtrmv overwrites B for the two calls after it).  Differences,
 * on purpose: the size line is parsed from the first non-comment line (the reference discards that line and reads the
 * next one), and times are wall-clock milliseconds (the reference uses clock(), i.e. CPU time).
 * With G4S_MV_DUMP=<path> the four results (C after dsymv, B after dtrmv, C after sspmv, C after dgemv; dim doubles
 * each) are written to <path>: tests/test_drivers.py compares them with the reference's own mv.c functions. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "g4s_b200.h"

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

int main(int argc, char *argv[]) {
    if (argc != 2) {
        fprintf(stderr, "usage: %s <matrix.mtx>\n", argv[0]);
        return 1;
    }
    FILE *file = fopen(argv[1], "r");
    if (!file) {
        fprintf(stderr, "cannot open %s\n", argv[1]);
        return 1;
    }
    printf("reading matrix A from %s\n", argv[1]);
    char line[1024];
    int dim = 0, cols = 0, nnz = 0;
    while (fgets(line, sizeof line, file))
        if (line[0] != '%') break;
    if (sscanf(line, "%d %d %d", &dim, &cols, &nnz) != 3 || dim <= 0 || nnz < 0) {
        fprintf(stderr, "bad size line: %s", line);
        return 1;
    }
    double *A = (double *)calloc((size_t)dim * dim, sizeof(double));
    double *B = (double *)malloc(sizeof(double) * dim);
    double *C = (double *)malloc(sizeof(double) * dim);
    if (!A || !B || !C) {
        fprintf(stderr, "out of memory for a dense %d x %d matrix\n", dim, dim);
        return 1;
    }
    for (int i = 0; i < nnz; i++) {
        int row, col;
        if (!fgets(line, sizeof line, file) || sscanf(line, "%d %d", &row, &col) != 2) {
            fprintf(stderr, "entry %d is malformed\n", i);
            return 1;
        }
        if (row >= 1 && row <= dim && col >= 1 && col <= dim) A[(size_t)(row - 1) * dim + col - 1] = rand();
    }
    fclose(file);
    for (int i = 0; i < dim; i++) B[i] = 1.0;

    matrix_multiply_dgemv(A, B, C, 1);  /* warm-up: CUDA context creation is not part of any timing */
    for (int i = 0; i < dim; i++) B[i] = 1.0;
    const char *dump_path = getenv("G4S_MV_DUMP");
    FILE *dump = dump_path ? fopen(dump_path, "wb") : NULL;
    double t0 = now_ms();
    matrix_multiply_dsymv(A, B, C, dim);
    printf("matrix_multiply_dsymv time: %f ms\n", now_ms() - t0);
    if (dump) fwrite(C, sizeof(double), dim, dump);
    t0 = now_ms();
    matrix_multiply_dtrmv(A, B, C, dim);
    printf("matrix_multiply_dtrmv time: %f ms\n", now_ms() - t0);
    if (dump) fwrite(B, sizeof(double), dim, dump);
    t0 = now_ms();
    matrix_multiply_sspmv(A, B, C, dim);
    printf("matrix_multiply_sspmv time: %f ms\n", now_ms() - t0);
    if (dump) fwrite(C, sizeof(double), dim, dump);
    t0 = now_ms();
    matrix_multiply_dgemv(A, B, C, dim);
    printf("matrix_multiply_dgemv time: %f ms\n", now_ms() - t0);
    if (dump) {
        fwrite(C, sizeof(double), dim, dump);
        fclose(dump);
    }
    double sum = 0.0;
    for (int i = 0; i < dim; i++) sum += C[i];
    printf("checksum(C) %.17g\n", sum);
    free(A);
    free(B);
    free(C);
    return g4s_last_error()[0] ? 2 : 0;
}
