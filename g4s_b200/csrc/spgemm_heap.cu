// HeapSpGEMM on the GPU: C = A B by a k-way HEAP MERGE of the sorted rows of B, for rows of A of ANY
// length.  The reference merges column by column on CSC with std::make_heap / pop_heap (mm/inc/heap_mult.h:47-223,
// HeapEntry.h); on CSR the same algorithm runs row by row: row i of C is the merge of the rows B(k,:), k in A(i,:), taken
// through a binary min-heap of (head column, list) entries — O(flop log k) comparisons, no table, output sorted by
// construction.  The merge class of g4s_spgemm_device covers rows with at most 8 lists in registers; this is its
// general twin: one thread per row, the heap of row i lives in global scratch at [arpt[i], arpt[i+1]) (a heap never holds
// more entries than the row of A has, so nnz(A) slots serve every row at once; short heaps stay in L1).
// Two passes like every SpGEMM here (count, scan, fill).  Ties on the column are broken by the list number, so equal
// columns leave the heap in A's stored order and the sums are taken exactly as HashSpGEMM<false,true> takes them
// (j ascending, one product and one addition per term, mm/inc/hash_mult.h:579-600): values are bit-identical to the oracle.
// Needs B's rows sorted by column (G4S_ERR_INVALID otherwise: use g4s_spgemm_device, which hashes).
#include <algorithm>

#include "common.cuh"

namespace g4s {

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host, cudaStream_t stream);

struct HeapArgs {
    const int *arpt, *acol;
    const double *aval;
    const int *brpt, *bcol;
    const double *bval;
    int M;
    int *hkey, *hlist, *pos;  // scratch, nnz(A) entries each
    int *row_nnz;             // symbolic: out
    const int *crpt;          // numeric: in
    int *ccol;
    double *cval;
};

__device__ __forceinline__ bool heap_less(int c1, int j1, int c2, int j2) { return c1 < c2 || (c1 == c2 && j1 < j2); }

template <bool NUMERIC>
__global__ void __launch_bounds__(128) spgemm_heap_row_kernel(const HeapArgs a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.M) return;
    const int as = __ldg(a.arpt + row), k = __ldg(a.arpt + row + 1) - as;
    int *hk = a.hkey + as, *hl = a.hlist + as, *pos = a.pos + as;
    int size = 0;
    auto sift_down = [&](int i, int c, int j) {  // place (c, j) starting at hole i
        for (;;) {
            int ch = 2 * i + 1;
            if (ch >= size) break;
            if (ch + 1 < size && heap_less(hk[ch + 1], hl[ch + 1], hk[ch], hl[ch])) ++ch;
            if (!heap_less(hk[ch], hl[ch], c, j)) break;
            hk[i] = hk[ch];
            hl[i] = hl[ch];
            i = ch;
        }
        hk[i] = c;
        hl[i] = j;
    };
    // build: one entry per non-empty row of B, sifted up
    for (int j = 0; j < k; ++j) {
        const int kk = __ldg(a.acol + as + j);
        const int p = __ldg(a.brpt + kk);
        pos[j] = p;
        if (p < __ldg(a.brpt + kk + 1)) {
            const int c = __ldg(a.bcol + p);
            int i = size++;
            while (i > 0) {
                const int par = (i - 1) >> 1;
                if (!heap_less(c, j, hk[par], hl[par])) break;
                hk[i] = hk[par];
                hl[i] = hl[par];
                i = par;
            }
            hk[i] = c;
            hl[i] = j;
        }
    }
    const int out = NUMERIC ? __ldg(a.crpt + row) : 0;
    int n = 0;
    while (size > 0) {
        const int c = hk[0];
        double v = 0.0;
        bool first = true;
        while (size > 0 && hk[0] == c) {  // every list that carries column c, in list order
            const int j = hl[0];
            const int p = pos[j];
            if (NUMERIC) {
                const double prod = __dmul_rn(__ldg(a.aval + as + j), __ldg(a.bval + p));
                v = first ? prod : __dadd_rn(prod, v);
                first = false;
            }
            pos[j] = p + 1;
            if (p + 1 < __ldg(a.brpt + __ldg(a.acol + as + j) + 1)) {
                sift_down(0, __ldg(a.bcol + p + 1), j);  // the list's next head replaces the root
            } else {
                --size;
                if (size > 0) sift_down(0, hk[size], hl[size]);
            }
        }
        if (NUMERIC) {
            a.ccol[out + n] = c;
            a.cval[out + n] = v;
        }
        ++n;
    }
    if (!NUMERIC) a.row_nnz[row] = n;
}

__global__ void heap_rows_sorted_kernel(const int *__restrict__ rowptr, const int *__restrict__ colids, int rows,
                                        int *__restrict__ unsorted) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int s = rowptr[warp], e = rowptr[warp + 1];
    int bad = 0;
    for (int k = s + 1 + lane; k < e; k += 32) bad |= __ldg(colids + k) <= __ldg(colids + k - 1);
    if (__any_sync(0xffffffffu, bad) && lane == 0) *unsorted = 1;
}

int spgemm_heap_run(g4s_csr *A, g4s_csr *B, g4s_csr **Cout, cudaStream_t stream) {
    if (A->cols != B->rows) return fail(G4S_ERR_SHAPE, "g4s_spgemm_heap: A.cols != B.rows");
    const int M = A->rows;
    int *scratch = nullptr, *row_nnz = nullptr, *flag = nullptr;
    g4s_csr *C = new (std::nothrow) g4s_csr();
    if (!C) return fail(G4S_ERR_ALLOC, "host allocation failed");
    C->rows = M;
    C->cols = B->cols;
    C->owns = true;
    C->pooled = true;
    auto bail = [&](int rc) {
        if (scratch) cudaFreeAsync(scratch, stream);
        if (row_nnz) cudaFreeAsync(row_nnz, stream);
        if (flag) cudaFreeAsync(flag, stream);
        g4s_csr_destroy(C);
        return rc;
    };
#define HEAP_CUDA(expr)                                                                                       \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return bail(fail(G4S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
    } while (0)
    if (B->sorted_cols < 0) {
        int h = 0;
        HEAP_CUDA(cudaMallocAsync(&flag, sizeof(int), stream));
        HEAP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
        if (B->rows > 0) {
            heap_rows_sorted_kernel<<<(int)(((long long)B->rows * 32 + 255) / 256), 256, 0, stream>>>(B->rowptr, B->colids, B->rows, flag);
            count_launch();
        }
        HEAP_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        HEAP_CUDA(cudaStreamSynchronize(stream));
        B->sorted_cols = h ? 0 : 1;
    }
    if (B->sorted_cols != 1)
        return bail(fail(G4S_ERR_INVALID, "g4s_spgemm_heap: the rows of B must be sorted by column (g4s_spgemm_device has no such need)"));
    const size_t na = (size_t)std::max<long long>(A->nnz, 1);
    HEAP_CUDA(cudaMallocAsync(&scratch, sizeof(int) * 3 * na, stream));
    HEAP_CUDA(cudaMallocAsync(&row_nnz, sizeof(int) * ((size_t)M + 1), stream));
    HEAP_CUDA(cudaMallocAsync(&C->rowptr, sizeof(int) * ((size_t)M + 1) + 64, stream));
    HeapArgs a;
    a.arpt = A->rowptr;
    a.acol = A->colids;
    a.aval = A->values;
    a.brpt = B->rowptr;
    a.bcol = B->colids;
    a.bval = B->values;
    a.M = M;
    a.hkey = scratch;
    a.hlist = scratch + na;
    a.pos = scratch + 2 * na;
    a.row_nnz = row_nnz;
    a.crpt = nullptr;
    a.ccol = nullptr;
    a.cval = nullptr;
    const int grid = (M + 127) / 128;
    if (M > 0) {
        spgemm_heap_row_kernel<false><<<grid, 128, 0, stream>>>(a);
        count_launch();
    }
    long long cnnz = 0;
    int rc = exclusive_scan_i32(row_nnz, C->rowptr, M, 1, &cnnz, stream);
    if (rc == G4S_OK && cnnz > 2147483647LL) rc = fail(G4S_ERR_INVALID, "g4s_spgemm_heap: nnz(C) exceeds int32 row pointers");
    if (rc) return bail(rc);
    C->nnz = cnnz;
    HEAP_CUDA(cudaMallocAsync(&C->colids, sizeof(int) * (size_t)cnnz + 64, stream));
    HEAP_CUDA(cudaMallocAsync(&C->values, sizeof(double) * (size_t)cnnz + 64, stream));
    a.crpt = C->rowptr;
    a.ccol = C->colids;
    a.cval = C->values;
    if (M > 0) {
        spgemm_heap_row_kernel<true><<<grid, 128, 0, stream>>>(a);
        count_launch();
    }
    HEAP_CUDA(cudaGetLastError());
    cudaFreeAsync(scratch, stream);
    cudaFreeAsync(row_nnz, stream);
    if (flag) cudaFreeAsync(flag, stream);
    scratch = row_nnz = flag = nullptr;
    HEAP_CUDA(cudaStreamSynchronize(stream));
#undef HEAP_CUDA
    C->sorted_cols = 1;
    *Cout = C;
    return G4S_OK;
}

}  // namespace g4s

extern "C" int g4s_spgemm_heap_device(g4s_csr_t A, g4s_csr_t B, g4s_csr_t *C, void *stream) {
    if (!A || !B || !C) return g4s::fail(G4S_ERR_INVALID, "g4s_spgemm_heap_device: null argument");
    int rc = g4s::ensure_device();
    if (rc) return rc;
    return g4s::spgemm_heap_run(A, B, C, (cudaStream_t)stream);
}
