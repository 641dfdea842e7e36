// SpGEMM by expand - sort - compress, the GPU form of the reference's OuterSpGEMM "join" (mm/inc/outer_mult.h:271-542):
// every intermediate product becomes a (row, col, value) triple (the reference's tuple<int,int,double>, :206-211), the
// triples are radix-sorted by the 64-bit key (row << 32) | col (its ExtractKey, :213-222) and equal keys are summed
// (its doMerge, :239-255).  Same C as the hash path: CSR, columns ascending.  The reference expands by columns of A (CSC)
// x rows of B; expanding by rows of A produces the same multiset of triples.  The sort is stable and the compress sums a
// run left to right, so the order inside a run is the expansion order (j ascending over A's row) — HashSpGEMM's sequential
// accumulation order: the VALUES of this path are bit-identical to the reference's HashSpGEMM<false,true> for every row.
// This is the SECOND, independent SpGEMM of the library: 16 bytes of traffic per product and per sort pass make it slower
// than the hash path on every input measured (profiles/), and it needs 32 bytes of scratch per product; it exists for
// the reference's algorithm inventory and as a device-side cross-check of the hash kernels.
// The sort is CUB's DeviceRadixSort (library code, as in rmat.cu); expansion, compress and row pointers are ours.

#include "common.cuh"

namespace g4s {

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host,
                       cudaStream_t stream);

// products of every row of A (BIN::set_intprod_num, mm/inc/BIN.h:77-95), one warp per row
__global__ void esc_row_products_kernel(const int *__restrict__ arpt, const int *__restrict__ acol,
                                        const int *__restrict__ brpt, int M, int *__restrict__ row_products,
                                        int *__restrict__ overflow) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= M) return;
    long long w = 0;
    for (int j = __ldg(arpt + row) + lane; j < __ldg(arpt + row + 1); j += 32) {
        const int k = __ldg(acol + j);
        w += __ldg(brpt + k + 1) - __ldg(brpt + k);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) {
        if (w > 2147483647LL) {
            *overflow = 1;
            w = 0;
        }
        row_products[row] = (int)w;
    }
}

// expand: one warp per row of A; entry j of the row owns the slots [base, base + len(B row)) in expansion order
__global__ void esc_expand_kernel(const int *__restrict__ arpt, const int *__restrict__ acol,
                                  const double *__restrict__ aval, const int *__restrict__ brpt,
                                  const int *__restrict__ bcol, const double *__restrict__ bval, int M,
                                  const int *__restrict__ offset, unsigned long long *__restrict__ keys,
                                  double *__restrict__ vals) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= M) return;
    long long base = offset[row];
    const unsigned long long hi = (unsigned long long)(unsigned)row << 32;
    for (int j = __ldg(arpt + row); j < __ldg(arpt + row + 1); ++j) {
        const int k = __ldg(acol + j);
        const double av = __ldg(aval + j);
        const int bs = __ldg(brpt + k), be = __ldg(brpt + k + 1);
        for (int p = bs + lane; p < be; p += 32) {
            keys[base + (p - bs)] = hi | (unsigned)__ldg(bcol + p);
            vals[base + (p - bs)] = __dmul_rn(av, __ldg(bval + p));
        }
        base += be - bs;
    }
}

// compress, pass 1: head[i] = 1 when triple i starts a run of equal keys
__global__ void esc_heads_kernel(const unsigned long long *__restrict__ keys, long long n, int *__restrict__ head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// compress, pass 2: the thread at the head of a run sums it left to right (the reference's doMerge order), writes the
// column and the value, and counts the entry for its row
__global__ void esc_compress_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
                                    long long n, const int *__restrict__ head, const int *__restrict__ slot,
                                    int *__restrict__ ccol, double *__restrict__ cval, int *__restrict__ row_nnz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const unsigned long long key = keys[i];
    double v = vals[i];
    for (long long q = i + 1; q < n && keys[q] == key; ++q) v = __dadd_rn(vals[q], v);
    const int o = slot[i];
    ccol[o] = (int)(unsigned)(key & 0xffffffffull);
    cval[o] = v;
    atomicAdd(&row_nnz[(int)(key >> 32)], 1);
}

int spgemm_esc_run(g4s_csr *A, g4s_csr *B, g4s_csr **Cout, cudaStream_t stream) {
    if (A->cols != B->rows) return fail(G4S_ERR_SHAPE, "g4s_spgemm_esc: A.cols != B.rows");
    const int M = A->rows;
    g4s_csr *C = new (std::nothrow) g4s_csr();
    if (!C) return fail(G4S_ERR_ALLOC, "host allocation failed");
    C->rows = M;
    C->cols = B->cols;
    C->owns = true;
    C->pooled = true;
    int *row_products = nullptr, *offset = nullptr, *overflow = nullptr, *head = nullptr, *slot = nullptr, *row_nnz = nullptr;
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    double *v0 = nullptr, *v1 = nullptr;
    void *tmp = nullptr;
    auto release = [&]() {
        for (void *p : {(void *)row_products, (void *)offset, (void *)overflow, (void *)head, (void *)slot, (void *)row_nnz,
                        (void *)k0, (void *)k1, (void *)v0, (void *)v1, tmp})
            if (p) cudaFreeAsync(p, stream);
    };
#define ESC_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            release();                                                                             \
            g4s_csr_destroy(C);                                                                    \
            return fail(G4S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
        }                                                                                          \
    } while (0)
    ESC_CUDA(cudaMallocAsync(&C->rowptr, sizeof(int) * ((size_t)M + 1) + 64, stream));
    ESC_CUDA(cudaMallocAsync(&row_products, sizeof(int) * ((size_t)M + 1), stream));
    ESC_CUDA(cudaMallocAsync(&offset, sizeof(int) * ((size_t)M + 1), stream));
    ESC_CUDA(cudaMallocAsync(&row_nnz, sizeof(int) * ((size_t)M + 1), stream));
    ESC_CUDA(cudaMallocAsync(&overflow, sizeof(int), stream));
    ESC_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), stream));
    ESC_CUDA(cudaMemsetAsync(row_nnz, 0, sizeof(int) * ((size_t)M + 1), stream));
    const int wblocks = (int)(((long long)M * 32 + 255) / 256);
    long long total = 0, cnnz = 0;
    int over = 0;
    if (M > 0) {
        esc_row_products_kernel<<<wblocks, 256, 0, stream>>>(A->rowptr, A->colids, B->rowptr, M, row_products, overflow);
        count_launch();
    }
    int rc = exclusive_scan_i32(row_products, offset, M, 1, &total, stream);
    if (rc == G4S_OK) {
        ESC_CUDA(cudaMemcpyAsync(&over, overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ESC_CUDA(cudaStreamSynchronize(stream));
        if (over || total > 2147483647LL)
            rc = fail(G4S_ERR_INVALID, "g4s_spgemm_esc: more than 2^31-1 intermediate products (use g4s_spgemm_device)");
    }
    if (rc) {
        release();
        g4s_csr_destroy(C);
        return rc;
    }
    const long long n = total;
    if (n > 0) {
        ESC_CUDA(cudaMallocAsync(&k0, sizeof(unsigned long long) * (size_t)n, stream));
        ESC_CUDA(cudaMallocAsync(&k1, sizeof(unsigned long long) * (size_t)n, stream));
        ESC_CUDA(cudaMallocAsync(&v0, sizeof(double) * (size_t)n, stream));
        ESC_CUDA(cudaMallocAsync(&v1, sizeof(double) * (size_t)n, stream));
        esc_expand_kernel<<<wblocks, 256, 0, stream>>>(A->rowptr, A->colids, A->values, B->rowptr, B->colids, B->values, M,
                                                       offset, k0, v0);
        count_launch();
        int rbits = 1;
        while ((1LL << rbits) < M) ++rbits;
        // the library's own stable LSD radix sort (radix_sort.cu): (row << 32 | col) keys, values ride along
        rc = radix_sort_pairs_u64(k0, k1, reinterpret_cast<unsigned long long *>(v0), reinterpret_cast<unsigned long long *>(v1), n,
                                  32 + rbits, stream);
        if (rc) {
            release();
            g4s_csr_destroy(C);
            return rc;
        }
        ESC_CUDA(cudaMallocAsync(&head, sizeof(int) * (size_t)n, stream));
        ESC_CUDA(cudaMallocAsync(&slot, sizeof(int) * (size_t)n, stream));
        const int nblocks = (int)((n + 255) / 256);
        esc_heads_kernel<<<nblocks, 256, 0, stream>>>(k1, n, head);
        count_launch();
        rc = exclusive_scan_i32(head, slot, n, 0, &cnnz, stream);
        if (rc) {
            release();
            g4s_csr_destroy(C);
            return rc;
        }
        ESC_CUDA(cudaMallocAsync(&C->colids, sizeof(int) * (size_t)cnnz + 64, stream));
        ESC_CUDA(cudaMallocAsync(&C->values, sizeof(double) * (size_t)cnnz + 64, stream));
        esc_compress_kernel<<<nblocks, 256, 0, stream>>>(k1, v1, n, head, slot, C->colids, C->values, row_nnz);
        count_launch();
    } else {
        ESC_CUDA(cudaMallocAsync(&C->colids, 64, stream));
        ESC_CUDA(cudaMallocAsync(&C->values, 64, stream));
    }
    C->nnz = cnnz;
    long long check_total = 0;
    rc = exclusive_scan_i32(row_nnz, C->rowptr, M, 1, &check_total, stream);
    release();
    ESC_CUDA(cudaStreamSynchronize(stream));
#undef ESC_CUDA
    if (rc == G4S_OK && check_total != cnnz) rc = fail(G4S_ERR_CUDA, "g4s_spgemm_esc: row counts do not add up to nnz(C)");
    if (rc) {
        g4s_csr_destroy(C);
        return rc;
    }
    *Cout = C;
    return G4S_OK;
}

}  // namespace g4s

extern "C" int g4s_spgemm_esc_device(g4s_csr_t A, g4s_csr_t B, g4s_csr_t *C, void *stream) {
    if (!A || !B || !C) return g4s::fail(G4S_ERR_INVALID, "g4s_spgemm_esc_device: null argument");
    int rc = g4s::ensure_device();
    if (rc) return rc;
    return g4s::spgemm_esc_run(A, B, C, (cudaStream_t)stream);
}
