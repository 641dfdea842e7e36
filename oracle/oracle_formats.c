/* TEST INFRASTRUCTURE ONLY — see oracle.h for scope and pinning status.
 *
 * Plain-C restatement of the reference's input formats:
 *   MatrixMarket coordinate file -> CSR : CSR<IT,NT>::construct, mm/inc/CSR.h:485-669 (banner :440-478)
 *   sorted edge list (class graph)  -> CSR : CSR<IT,NT>::CSR(graph&), mm/inc/CSR.h:255-329, graph.h:4-25
 *   sub-matrix extraction               : CSR(const CSR&, M_, N_, M_start, N_start), mm/inc/CSR.h:691-733
 * Pinned against the reference itself (oracle/_ref/libg4s_ref.so) by tests/test_oracle.py. */
#include "oracle.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int fail(char *err, int errlen, const char *msg) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg);
    return -1;
}

static int split_ws(char *line, char **tok, int maxtok) {
    int n = 0;
    char *p = line;
    while (*p) {
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        if (n < maxtok) tok[n] = p;
        ++n;
        while (*p && !isspace((unsigned char)*p)) ++p;
        if (*p) *p++ = '\0';
    }
    return n;
}

typedef struct {
    long key;
    double val;
} keyval_t;

/* stable merge sort by key.  The reference uses std::sort (CSR.h:645), which leaves the relative order of
 * duplicate (row,col) entries unspecified; stable order is one of the outcomes it permits. */
static void msort(keyval_t *a, keyval_t *tmp, long n) {
    if (n < 2) return;
    long h = n / 2;
    msort(a, tmp, h);
    msort(a + h, tmp, n - h);
    long i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(keyval_t) * (size_t)n);
}

/* Returns 0, or -1 with a message in err for every input the reference rejects with std::runtime_error
 * (CSR.h:451-476, :490, :510, :557, :562). */
int oracle_mm_construct(const char *path, int *rows, int *cols, int *nnz, int **rowptr, int **colids,
                        double **values, char *err, int errlen) {
    FILE *f = fopen(path, "r");
    if (!f) return fail(err, errlen, "unable to open file");
    char line[1024];
    char *tok[8];
    /* banner: exactly five tokens "%%MatrixMarket matrix <storage> <type> <symmetry>" */
    if (!fgets(line, sizeof line, f)) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket banner");
    }
    int nt = split_ws(line, tok, 8);
    if (nt != 5 || strcmp(tok[0], "%%MatrixMarket") || strcmp(tok[1], "matrix")) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket banner");
    }
    char storage[32], type[32], symmetry[32];
    snprintf(storage, sizeof storage, "%s", tok[2]);
    snprintf(type, sizeof type, "%s", tok[3]);
    snprintf(symmetry, sizeof symmetry, "%s", tok[4]);
    if (strcmp(storage, "array") && strcmp(storage, "coordinate")) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket storage format");
    }
    if (!strcmp(storage, "array")) {
        fclose(f);
        return fail(err, errlen, "not impl storage type array");
    }
    int is_pattern = !strcmp(type, "pattern"), is_complex = !strcmp(type, "complex");
    if (!is_pattern && !is_complex && strcmp(type, "real") && strcmp(type, "integer")) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket data type");
    }
    int general = !strcmp(symmetry, "general"), symm = !strcmp(symmetry, "symmetric"),
        skew = !strcmp(symmetry, "skew-symmetric"), herm = !strcmp(symmetry, "hermitian");
    if (!general && !symm && !skew && !herm) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket symmetry");
    }
    if (herm) {
        fclose(f);
        return fail(err, errlen, "not impl matrix type: hermitian");
    }
    /* skip comment lines; the first non-'%' line holds "rows cols entries" */
    do {
        if (!fgets(line, sizeof line, f)) line[0] = '\0';
    } while (line[0] == '%');
    nt = split_ws(line, tok, 8);
    if (nt != 3) {
        fclose(f);
        return fail(err, errlen, "invalid MatrixMarket coordinate format");
    }
    long R = atol(tok[0]), C = atol(tok[1]), E = atol(tok[2]);
    if (E <= 0) {
        fclose(f);
        return fail(err, errlen, "something wrong: nnz is 0");
    }
    long cap = general ? E : 2 * E;
    keyval_t *kv = (keyval_t *)malloc(sizeof(keyval_t) * (size_t)cap);
    long n = 0, nread = 0;
    for (; nread < E; ++nread) {
        long i, j;
        double v = 1.0, im;
        if (fscanf(f, "%ld %ld", &i, &j) != 2) break;
        if (!is_pattern && fscanf(f, "%lf", &v) != 1) break;
        if (is_complex && fscanf(f, "%lf", &im) != 1) break;
        i -= 1;
        j -= 1;
        kv[n].key = C * i + j; /* (row, col) order as one 64-bit key, CSR.h:642 */
        kv[n].val = v;
        ++n;
        if (!general && i != j) { /* mirror off-diagonals right after their source entry, CSR.h:586-624 */
            kv[n].key = C * j + i;
            kv[n].val = skew ? -v : v;
            ++n;
        }
    }
    fclose(f);
    if (nread != E) {
        free(kv);
        return fail(err, errlen, "read nnz not equal to declared nnz");
    }
    keyval_t *tmp = (keyval_t *)malloc(sizeof(keyval_t) * (size_t)n);
    msort(kv, tmp, n);
    free(tmp);
    int *rp = (int *)calloc((size_t)R + 1, sizeof(int));
    int *ci = (int *)malloc(sizeof(int) * (size_t)n);
    double *va = (double *)malloc(sizeof(double) * (size_t)n);
    for (long e = 0; e < n; ++e) {
        rp[kv[e].key / C + 1]++;
        ci[e] = (int)(kv[e].key % C);
        va[e] = kv[e].val;
    }
    for (long r = 1; r <= R; ++r) rp[r] += rp[r - 1];
    free(kv);
    *rows = (int)R;
    *cols = (int)C;
    *nnz = (int)n;
    *rowptr = rp;
    *colids = ci;
    *values = va;
    return 0;
}

typedef struct {
    long s, e;
    double w;
} edge_t;
static int cmp_edge(const void *a, const void *b) { /* lexicographic (start, end, w) like std::pair */
    const edge_t *x = (const edge_t *)a, *y = (const edge_t *)b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    if (x->w != y->w) return x->w < y->w ? -1 : 1;
    return 0;
}

/* CSR(graph&): consecutive edges with the same start form a group; each group is sorted by (end, weight),
 * equal (start,end) pairs are summed left to right, and the merged triples are then bucketed by row in
 * arrival order.  (If the same start shows up in two separate groups the row keeps both runs unmerged,
 * exactly as the reference does.) */
int oracle_csr_from_graph(long m, long n, const long *start, const long *end, const double *w, int *nnz,
                          int **rowptr, int **colids, double **values) {
    edge_t *grp = (edge_t *)malloc(sizeof(edge_t) * (size_t)(m ? m : 1));
    edge_t *out = (edge_t *)malloc(sizeof(edge_t) * (size_t)(m ? m : 1));
    long nout = 0, g0 = 0;
    while (g0 < m) {
        long g1 = g0 + 1;
        while (g1 < m && start[g1] == start[g0]) ++g1;
        long len = g1 - g0;
        for (long k = 0; k < len; ++k) {
            grp[k].s = start[g0 + k];
            grp[k].e = end[g0 + k];
            grp[k].w = w[g0 + k];
        }
        qsort(grp, (size_t)len, sizeof(edge_t), cmp_edge);
        out[nout++] = grp[0];
        for (long k = 1; k < len; ++k) {
            if (grp[k].e == grp[k - 1].e) out[nout - 1].w += grp[k].w;
            else out[nout++] = grp[k];
        }
        g0 = g1;
    }
    int *rp = (int *)calloc((size_t)n + 1, sizeof(int));
    int *ci = (int *)malloc(sizeof(int) * (size_t)(nout ? nout : 1));
    double *va = (double *)malloc(sizeof(double) * (size_t)(nout ? nout : 1));
    for (long k = 0; k < nout; ++k) rp[out[k].s + 1]++;
    for (long r = 1; r <= n; ++r) rp[r] += rp[r - 1];
    int *fill = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    memcpy(fill, rp, sizeof(int) * (size_t)n);
    for (long k = 0; k < nout; ++k) {
        int pos = fill[out[k].s]++;
        ci[pos] = (int)out[k].e;
        va[pos] = out[k].w;
    }
    free(fill);
    free(out);
    free(grp);
    *nnz = (int)nout;
    *rowptr = rp;
    *colids = ci;
    *values = va;
    return 0;
}

/* Rows [M_start, M_start+M_) x columns [N_start, N_start+N_) of A, column ids rebased (CSR.h:691-733);
 * the driver uses it to trim A or B to conformable shapes (mm/src/mkl_spgemm.cpp:42-58). */
int oracle_csr_submatrix(int rows, int cols, const int *rowptr, const int *colids, const double *values, int M_,
                         int N_, int M_start, int N_start, int *nnz, int **orpt, int **ocol, double **oval) {
    if (M_ + M_start > rows || N_ + N_start > cols) return -1;
    int *rp = (int *)calloc((size_t)M_ + 1, sizeof(int));
    for (int i = 0; i < M_; ++i) {
        int c = 0;
        for (long j = rowptr[i + M_start]; j < rowptr[i + M_start + 1]; ++j)
            c += (colids[j] >= N_start && colids[j] < N_start + N_);
        rp[i + 1] = rp[i] + c;
    }
    int total = rp[M_];
    int *ci = (int *)malloc(sizeof(int) * (size_t)(total ? total : 1));
    double *va = (double *)malloc(sizeof(double) * (size_t)(total ? total : 1));
    for (int i = 0; i < M_; ++i) {
        int q = rp[i];
        for (long j = rowptr[i + M_start]; j < rowptr[i + M_start + 1]; ++j)
            if (colids[j] >= N_start && colids[j] < N_start + N_) {
                ci[q] = colids[j] - N_start;
                va[q++] = values[j];
            }
    }
    *nnz = total;
    *orpt = rp;
    *ocol = ci;
    *oval = va;
    return 0;
}
