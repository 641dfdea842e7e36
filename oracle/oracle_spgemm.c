/* TEST INFRASTRUCTURE ONLY — see oracle.h for scope and pinning status.
 *
 * Plain-C restatement of the reference's two-phase hash SpGEMM (C = A B on CSR<int,double>):
 *   work count + work-balanced row cut + table size class : mm/inc/BIN.h:77-177
 *   symbolic (hash-set insert, nnz per C row)             : mm/inc/hash_mult.h:64-109, :495-508
 *   numeric  (hash-map accumulate, compact, sort by col)  : mm/inc/hash_mult.h:525-608
 *   driver                                                : mm/inc/hash_mult.h:1028-1057
 * Pinned against the reference itself (oracle/_ref/libg4s_ref.so) by tests/test_oracle.py. */
#include "oracle.h"

#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define HASH_MULTIPLIER 107 /* mm/inc/hash_mult.h:23 */
#define MIN_TABLE 8         /* mm/inc/hash_mult.h:34-35 */

void oracle_free(void *p) { free(p); }

/* BIN::set_intprod_num (BIN.h:77-95): row_nz[i] = sum over stored A(i,k) of nnz(B(k,:)). */
long long oracle_intprod(const int *arpt, const int *acol, const int *brpt, int rows, int *row_nz) {
    long long total = 0;
    for (int i = 0; i < rows; ++i) {
        int w = 0;
        for (long j = arpt[i]; j < arpt[i + 1]; ++j) w += brpt[acol[j] + 1] - brpt[acol[j]];
        row_nz[i] = w;
        total += w;
    }
    return total;
}

/* BIN::set_rows_offset (BIN.h:100-122): exclusive prefix sum of the per-row work, then part p ends at the
 * first prefix entry >= ceil(total/parts)*(p+1) (std::lower_bound); the last cut is forced to `rows`.
 * The reference does the target arithmetic in `int` (BIN.h:108,116); 64-bit here, identical whenever the
 * reference does not overflow. */
void oracle_rows_offset(const int *row_nz, int rows, long long total, int parts, int *rows_offset) {
    long long *ps = (long long *)malloc(sizeof(long long) * ((size_t)rows + 1));
    ps[0] = 0;
    for (int i = 0; i < rows; ++i) ps[i + 1] = ps[i] + row_nz[i];
    long long avg = (total + parts - 1) / parts;
    rows_offset[0] = 0;
    for (int p = 0; p < parts; ++p) {
        long long target = avg * (p + 1);
        long lo = 0, hi = (long)rows + 1; /* lower_bound over ps[0..rows] */
        while (lo < hi) {
            long mid = lo + (hi - lo) / 2;
            if (ps[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        rows_offset[p + 1] = (int)lo;
    }
    rows_offset[parts] = rows;
    free(ps);
}

/* BIN::set_bin_id (BIN.h:157-177): class 0 for empty rows, else 1 + smallest j with
 * min(work, cols) <= min_ht << j; the row's table then has min_ht << (class-1) slots. */
void oracle_bin_id(const int *row_nz, int rows, int cols, int min_ht, signed char *bin_id) {
    for (int i = 0; i < rows; ++i) {
        int w = row_nz[i] > cols ? cols : row_nz[i];
        if (w == 0) {
            bin_id[i] = 0;
            continue;
        }
        int j = 0;
        while (w > (min_ht << j)) ++j;
        bin_id[i] = (signed char)(j + 1);
    }
}

static inline int table_size_for(int work, int cols) {
    int w = work > cols ? cols : work;
    if (w == 0) return 0;
    int s = MIN_TABLE;
    while (s < w) s <<= 1;
    return s;
}

/* One row of the symbolic phase: number of distinct columns among the intermediate products. */
static int symbolic_row(const int *arpt, const int *acol, const int *brpt, const int *bcol, int i, int *keys,
                        int tsize) {
    int nz = 0;
    for (int s = 0; s < tsize; ++s) keys[s] = -1;
    for (long j = arpt[i]; j < arpt[i + 1]; ++j) {
        int k = acol[j];
        for (long p = brpt[k]; p < brpt[k + 1]; ++p) {
            int key = bcol[p];
            int h = (key * HASH_MULTIPLIER) & (tsize - 1);
            for (;;) {
                if (keys[h] == key) break;
                if (keys[h] == -1) {
                    keys[h] = key;
                    ++nz;
                    break;
                }
                h = (h + 1) & (tsize - 1);
            }
        }
    }
    return nz;
}

static int cmp_col(const void *a, const void *b) {
    const int ca = *(const int *)a, cb = *(const int *)b;
    return (ca > cb) - (ca < cb);
}
typedef struct {
    int col;
    int pad;
    double val;
} colval_t;

/* One row of the numeric phase.  Products are formed in stored order (j over A's row, p over B's row) and
 * accumulated as  new = product + old  (hash_mult.h:584: addop(t_val, ht_value[hash])); the table is then
 * compacted in slot order and, when sort_output, sorted by column (keys are distinct, so any comparison
 * sort gives the same result as the reference's std::sort). */
static void numeric_row(const int *arpt, const int *acol, const double *aval, const int *brpt, const int *bcol,
                        const double *bval, int i, int *keys, double *vals, int tsize, int *ccol, double *cval,
                        int nz, int sort_output, colval_t *scratch) {
    for (int s = 0; s < tsize; ++s) keys[s] = -1;
    for (long j = arpt[i]; j < arpt[i + 1]; ++j) {
        int k = acol[j];
        double a = aval[j];
        for (long p = brpt[k]; p < brpt[k + 1]; ++p) {
            double t = a * bval[p];
            int key = bcol[p];
            int h = (key * HASH_MULTIPLIER) & (tsize - 1);
            for (;;) {
                if (keys[h] == key) {
                    vals[h] = t + vals[h];
                    break;
                }
                if (keys[h] == -1) {
                    keys[h] = key;
                    vals[h] = t;
                    break;
                }
                h = (h + 1) & (tsize - 1);
            }
        }
    }
    int n = 0;
    if (sort_output) {
        for (int s = 0; s < tsize; ++s)
            if (keys[s] != -1) {
                scratch[n].col = keys[s];
                scratch[n].val = vals[s];
                ++n;
            }
        qsort(scratch, (size_t)n, sizeof(colval_t), cmp_col);
        for (int s = 0; s < n; ++s) {
            ccol[s] = scratch[s].col;
            cval[s] = scratch[s].val;
        }
    } else {
        for (int s = 0; s < tsize; ++s)
            if (keys[s] != -1) {
                ccol[n] = keys[s];
                cval[n] = vals[s];
                ++n;
            }
    }
    (void)nz;
}

void oracle_hash_symbolic(const int *arpt, const int *acol, const int *brpt, const int *bcol, int rows, int cols,
                          int *crpt, int *cnnz) {
    int *work = (int *)malloc(sizeof(int) * (size_t)(rows ? rows : 1));
    oracle_intprod(arpt, acol, brpt, rows, work);
    int maxw = 0;
    for (int i = 0; i < rows; ++i)
        if (work[i] > maxw) maxw = work[i];
    int cap = table_size_for(maxw, cols);
    int *keys = (int *)malloc(sizeof(int) * (size_t)(cap ? cap : 1));
    crpt[0] = 0;
    for (int i = 0; i < rows; ++i) {
        int ts = table_size_for(work[i], cols);
        int nz = ts ? symbolic_row(arpt, acol, brpt, bcol, i, keys, ts) : 0;
        crpt[i + 1] = crpt[i] + nz; /* scan(row_nz -> crpt), hash_mult.h:506 */
    }
    *cnnz = crpt[rows];
    free(keys);
    free(work);
}

void oracle_hash_numeric(const int *arpt, const int *acol, const double *aval, const int *brpt, const int *bcol,
                         const double *bval, int rows, int cols, const int *crpt, int *ccol, double *cval,
                         int sort_output) {
    int *work = (int *)malloc(sizeof(int) * (size_t)(rows ? rows : 1));
    oracle_intprod(arpt, acol, brpt, rows, work);
    int maxw = 0;
    for (int i = 0; i < rows; ++i)
        if (work[i] > maxw) maxw = work[i];
    int cap = table_size_for(maxw, cols);
    int *keys = (int *)malloc(sizeof(int) * (size_t)(cap ? cap : 1));
    double *vals = (double *)malloc(sizeof(double) * (size_t)(cap ? cap : 1));
    colval_t *scratch = (colval_t *)malloc(sizeof(colval_t) * (size_t)(cap ? cap : 1));
    for (int i = 0; i < rows; ++i) {
        int ts = table_size_for(work[i], cols);
        if (ts)
            numeric_row(arpt, acol, aval, brpt, bcol, bval, i, keys, vals, ts, ccol + crpt[i], cval + crpt[i],
                        crpt[i + 1] - crpt[i], sort_output, scratch);
    }
    free(scratch);
    free(vals);
    free(keys);
    free(work);
}

/* HashSpGEMM<false,true> (hash_mult.h:1028-1057, :1109-1113).  Output arrays are malloc'd; free with
 * oracle_free.  Returns 0. */
int oracle_hash_spgemm(int M, int K, int N, const int *arpt, const int *acol, const double *aval, const int *brpt,
                       const int *bcol, const double *bval, int *cnnz, int **crpt, int **ccol, double **cval) {
    (void)K;
    *crpt = (int *)malloc(sizeof(int) * ((size_t)M + 1));
    oracle_hash_symbolic(arpt, acol, brpt, bcol, M, N, *crpt, cnnz);
    *ccol = (int *)malloc(sizeof(int) * (size_t)(*cnnz ? *cnnz : 1));
    *cval = (double *)malloc(sizeof(double) * (size_t)(*cnnz ? *cnnz : 1));
    oracle_hash_numeric(arpt, acol, aval, brpt, bcol, bval, M, N, *crpt, *ccol, *cval, 1);
    return 0;
}

/* The same algorithm with the reference's thread structure: rows cut by oracle_rows_offset into
 * `threads` work-balanced ranges, one reusable table per thread (BIN.h:128-151).  Each C row is still
 * produced by one thread in stored order, so the output is bit-identical to oracle_hash_spgemm.
 * Returns wall seconds of the multiply (used as the "port" CPU baseline when oracle/_ref is absent). */
double oracle_hash_spgemm_omp(int threads, int M, int K, int N, const int *arpt, const int *acol,
                              const double *aval, const int *brpt, const int *bcol, const double *bval, int *cnnz,
                              int **crpt, int **ccol, double **cval) {
    (void)K;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int *work = (int *)malloc(sizeof(int) * (size_t)(M ? M : 1));
    long long total = oracle_intprod(arpt, acol, brpt, M, work);
    int *cut = (int *)malloc(sizeof(int) * ((size_t)threads + 1));
    oracle_rows_offset(work, M, total, threads, cut);
    /* the reference's cut can land on rows+1 for tiny inputs (total < T(T-1)), where it then reads past the
     * last row; the restatement clamps instead of reproducing the out-of-bounds access */
    for (int p = 0; p <= threads; ++p)
        if (cut[p] > M) cut[p] = M;
    int *rownz = (int *)malloc(sizeof(int) * ((size_t)M + 1));
    int **tkeys = (int **)calloc((size_t)threads, sizeof(int *));
    double **tvals = (double **)calloc((size_t)threads, sizeof(double *));
    colval_t **tscr = (colval_t **)calloc((size_t)threads, sizeof(colval_t *));
#pragma omp parallel num_threads(threads)
    {
        int t = omp_get_thread_num();
        int maxw = 0;
        for (int i = cut[t]; i < cut[t + 1]; ++i)
            if (work[i] > maxw) maxw = work[i];
        int cap = table_size_for(maxw, N);
        tkeys[t] = (int *)malloc(sizeof(int) * (size_t)(cap ? cap : 1));
        tvals[t] = (double *)malloc(sizeof(double) * (size_t)(cap ? cap : 1));
        tscr[t] = (colval_t *)malloc(sizeof(colval_t) * (size_t)(cap ? cap : 1));
        for (int i = cut[t]; i < cut[t + 1]; ++i) {
            int ts = table_size_for(work[i], N);
            rownz[i] = ts ? symbolic_row(arpt, acol, brpt, bcol, i, tkeys[t], ts) : 0;
        }
    }
    *crpt = (int *)malloc(sizeof(int) * ((size_t)M + 1));
    (*crpt)[0] = 0;
    for (int i = 0; i < M; ++i) (*crpt)[i + 1] = (*crpt)[i] + rownz[i];
    *cnnz = (*crpt)[M];
    *ccol = (int *)malloc(sizeof(int) * (size_t)(*cnnz ? *cnnz : 1));
    *cval = (double *)malloc(sizeof(double) * (size_t)(*cnnz ? *cnnz : 1));
    const int *rp = *crpt;
    int *cc = *ccol;
    double *cv = *cval;
#pragma omp parallel num_threads(threads)
    {
        int t = omp_get_thread_num();
        for (int i = cut[t]; i < cut[t + 1]; ++i) {
            int ts = table_size_for(work[i], N);
            if (ts)
                numeric_row(arpt, acol, aval, brpt, bcol, bval, i, tkeys[t], tvals[t], ts, cc + rp[i], cv + rp[i],
                            rp[i + 1] - rp[i], 1, tscr[t]);
        }
        free(tkeys[t]);
        free(tvals[t]);
        free(tscr[t]);
    }
    free(tkeys);
    free(tvals);
    free(tscr);
    free(rownz);
    free(cut);
    free(work);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
