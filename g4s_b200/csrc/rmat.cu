// Graph-to-CSR loader on the device, and the R-MAT input of BASELINE config 3.
//
//   g4s_csr_from_edges_device : the reference's CSR(graph&) (mm/inc/CSR.h:255-329; edge list laid out as class
//       graph, mm/inc/graph.h:4-25) for edge lists that already live on the GPU: edges are ordered by
//       (start, end), equal pairs are SUMMED, rows are counted into rowptr.  Unlike the reference the input
//       need not arrive grouped by start vertex.  Duplicates are added in arrival order (the reference adds
//       them in ascending-weight order), so merged values agree to rounding, the pattern exactly.
//   g4s_rmat_edges_device     : Graph500 R-MAT edges (a,b,c,d = .57,.19,.19,.05), counter-based: edge e at level
//       l draws from splitmix64(seed, e, l), so any slice of the list can be regenerated anywhere.
//   g4s_csr_generate_rmat     : the two composed.
//
// The ordering step is a key-value radix sort from CUB (header-only, ships with the CUDA toolkit).  It runs
// once per input in set-up code, outside every timed region; the measured hot paths (SpMV, SpGEMM) contain no
// library kernels.

#include <algorithm>

#include "common.cuh"

namespace g4s {

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host,
                       cudaStream_t stream);
int alloc_csr(g4s_csr **out, int rows, int cols, long long nnz);

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void rmat_edges_kernel(int scale, long long m, unsigned long long seed, long *__restrict__ start,
                                  long *__restrict__ end, double *__restrict__ w) {
    const double a = 0.57, ab = 0.57 + 0.19, abc = 0.57 + 0.19 + 0.19;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x) {
        const unsigned long long base = splitmix64(seed ^ splitmix64((unsigned long long)e));
        long r = 0, c = 0;
        for (int l = 0; l < scale; ++l) {
            const double u = (double)(splitmix64(base + (unsigned long long)l) >> 11) * (1.0 / 9007199254740992.0);
            const int rb = u >= ab, cb = (u >= a && u < ab) || u >= abc;
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        start[e] = r;
        end[e] = c;
        w[e] = (double)(splitmix64(base + 4096ULL) >> 11) * (1.0 / 9007199254740992.0);
    }
}

__global__ void pack_keys_kernel(const long *__restrict__ start, const long *__restrict__ end, long long m,
                                 unsigned long long *__restrict__ keys) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x)
        keys[e] = ((unsigned long long)start[e] << 32) | (unsigned long long)(unsigned int)end[e];
}
__global__ void check_range_kernel(const long *__restrict__ start, const long *__restrict__ end, long long m, long n,
                                   int *__restrict__ bad) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x)
        if (start[e] < 0 || start[e] >= n || end[e] < 0 || end[e] >= n) *bad = 1;
}
__global__ void head_flags_kernel(const unsigned long long *__restrict__ keys, long long m, int *__restrict__ head) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x)
        head[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1 : 0;
}
// one thread per sorted edge: heads write their merged entry and the row pointers of every row that starts
// between the previous entry's row and their own
__global__ void merge_edges_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ w,
                                   const int *__restrict__ head, const int *__restrict__ idx, long long m, int n,
                                   int unique, int *__restrict__ rowptr, int *__restrict__ colids,
                                   double *__restrict__ values) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[e];
        const int row = (int)(key >> 32);
        if (head[e]) {
            const int j = idx[e];
            double s = w[e];
            for (long long f = e + 1; f < m && keys[f] == key; ++f) s += w[f];
            colids[j] = (int)(key & 0xffffffffULL);
            values[j] = s;
            const int prev_row = e ? (int)(keys[e - 1] >> 32) : -1;
            for (int r = prev_row + 1; r <= row; ++r) rowptr[r] = j;
        }
        if (e == m - 1)
            for (int r = row + 1; r <= n; ++r) rowptr[r] = unique;
    }
}

int edges_to_csr(long long m, long n, const long *start, const long *end, const double *w, g4s_csr **out,
                 cudaStream_t stream) {
    if (n > 2147483646L) return fail(G4S_ERR_INVALID, "edge list: more than 2^31-2 vertices");
    g4s_csr *h = nullptr;
    int rc;
    if (m == 0) {
        if ((rc = alloc_csr(&h, (int)n, (int)n, 0))) return rc;
        G4S_CUDA(cudaMemsetAsync(h->rowptr, 0, sizeof(int) * ((size_t)n + 1), stream));
        *out = h;
        return G4S_OK;
    }
    const int grid = sm_count() * 8;
    int *bad = nullptr;
    G4S_CUDA(cudaMalloc(&bad, sizeof(int)));
    G4S_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    check_range_kernel<<<grid, 256, 0, stream>>>(start, end, m, n, bad);
    G4S_CHECK_LAUNCH("check_range_kernel");
    int hbad = 0;
    G4S_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    G4S_CUDA(cudaStreamSynchronize(stream));
    cudaFree(bad);
    if (hbad) return fail(G4S_ERR_FORMAT, "edge list: vertex id out of range");

    unsigned long long *k0 = nullptr, *k1 = nullptr;
    double *w1 = nullptr;
    G4S_CUDA(cudaMalloc(&k0, sizeof(unsigned long long) * (size_t)m));
    G4S_CUDA(cudaMalloc(&k1, sizeof(unsigned long long) * (size_t)m));
    G4S_CUDA(cudaMalloc(&w1, sizeof(double) * (size_t)m));
    pack_keys_kernel<<<grid, 256, 0, stream>>>(start, end, m, k0);
    G4S_CHECK_LAUNCH("pack_keys_kernel");
    int vbits = 1;
    while ((1L << vbits) < n) ++vbits;
    // the library's own stable radix sort (radix_sort.cu) ping-pongs between two pairs of buffers: the caller's weights are
    // copied first, because the even passes write into the source pair
    double *w0 = nullptr;
    G4S_CUDA(cudaMalloc(&w0, sizeof(double) * (size_t)m));
    G4S_CUDA(cudaMemcpyAsync(w0, w, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, stream));
    if ((rc = radix_sort_pairs_u64(k0, k1, reinterpret_cast<unsigned long long *>(w0), reinterpret_cast<unsigned long long *>(w1), m,
                                   32 + vbits, stream)))
        return rc;
    int *head = nullptr, *idx = nullptr;
    G4S_CUDA(cudaMalloc(&head, sizeof(int) * ((size_t)m + 1)));
    G4S_CUDA(cudaMalloc(&idx, sizeof(int) * ((size_t)m + 1)));
    head_flags_kernel<<<grid, 256, 0, stream>>>(k1, m, head);
    G4S_CHECK_LAUNCH("head_flags_kernel");
    long long unique = 0;
    if ((rc = exclusive_scan_i32(head, idx, m, 0, &unique, stream))) return rc;
    if (unique > 2147483647LL || n > 2147483647L)
        return fail(G4S_ERR_INVALID, "edge list: more than 2^31-1 distinct edges or vertices (int32 CSR)");
    if ((rc = alloc_csr(&h, (int)n, (int)n, unique))) return rc;
    merge_edges_kernel<<<grid, 256, 0, stream>>>(k1, w1, head, idx, m, (int)n, (int)unique, h->rowptr, h->colids, h->values);
    G4S_CHECK_LAUNCH("merge_edges_kernel");
    G4S_CUDA(cudaStreamSynchronize(stream));
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(w1);
    cudaFree(w0);
    cudaFree(head);
    cudaFree(idx);
    *out = h;
    return G4S_OK;
}

}  // namespace g4s

using namespace g4s;

extern "C" {

int g4s_rmat_edges_device(int scale, long long m, unsigned long long seed, long *start_dev, long *end_dev,
                          double *w_dev, void *stream) {
    if (scale < 1 || scale > 30 || m < 0 || (m && (!start_dev || !end_dev || !w_dev)))
        return fail(G4S_ERR_INVALID, "g4s_rmat_edges_device: bad arguments (1 <= scale <= 30)");
    int rc = ensure_device();
    if (rc) return rc;
    if (m == 0) return G4S_OK;
    rmat_edges_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(scale, m, seed, start_dev, end_dev, w_dev);
    G4S_CHECK_LAUNCH("rmat_edges_kernel");
    return G4S_OK;
}

int g4s_csr_from_edges_device(long m, long n, const long *start_dev, const long *end_dev, const double *w_dev,
                              g4s_csr_t *out, void *stream) {
    if (m < 0 || n < 0 || !out || (m && (!start_dev || !end_dev || !w_dev)))
        return fail(G4S_ERR_INVALID, "g4s_csr_from_edges_device: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    return edges_to_csr(m, n, start_dev, end_dev, w_dev, out, (cudaStream_t)stream);
}

int g4s_csr_generate_rmat(g4s_csr_t *out, int scale, int edge_factor, unsigned long long seed, void *stream) {
    if (!out || scale < 1 || scale > 30 || edge_factor < 1) return fail(G4S_ERR_INVALID, "g4s_csr_generate_rmat: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    const long n = 1L << scale;
    const long long m = (long long)edge_factor << scale;
    long *s = nullptr, *e = nullptr;
    double *w = nullptr;
    G4S_CUDA(cudaMalloc(&s, sizeof(long) * (size_t)m));
    G4S_CUDA(cudaMalloc(&e, sizeof(long) * (size_t)m));
    G4S_CUDA(cudaMalloc(&w, sizeof(double) * (size_t)m));
    rc = g4s_rmat_edges_device(scale, m, seed, s, e, w, stream);
    if (rc == G4S_OK) rc = edges_to_csr(m, n, s, e, w, out, (cudaStream_t)stream);
    cudaFree(s);
    cudaFree(e);
    cudaFree(w);
    return rc;
}

}  // extern "C"
