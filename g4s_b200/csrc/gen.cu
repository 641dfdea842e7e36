// Synthetic inputs of BASELINE.json's configs, generated directly in CSR on the device.
// Closed-form row pointers: with c(t) = 2 on a grid boundary and 3 inside (n >= 2; 1 when n == 1),
// P(t) = sum_{t'<t} c(t') and S = P(n), the 27-point row (k,j,i) starts at
//     P(k) S^2 + c(k) (P(j) S + c(j) P(i))
// and the 5-point row (j,i) has 1 + [i>0] + [i<n-1] + [j>0] + [j<n-1] entries.
#include "common.cuh"

namespace g4s {

__host__ __device__ inline long long cnt1(int n, long long t) {  // neighbours of t inside [0,n), incl. itself
    return 1 + (t > 0) + (t < n - 1);
}
__host__ __device__ inline long long pre1(int n, long long t) {  // sum_{t'<t} cnt1(t')
    if (t <= 0) return 0;
    if (n == 1) return 1;
    if (t >= n) return 3LL * n - 2;
    return 3 * t - 1;
}
__host__ __device__ inline long long lap3d_row_start(int n, long long row) {
    const long long S = pre1(n, n);
    const long long i = row % n, j = (row / n) % n, k = row / ((long long)n * n);
    if (k >= n) return S * S * S;
    return pre1(n, k) * S * S + cnt1(n, k) * (pre1(n, j) * S + cnt1(n, j) * pre1(n, i));
}
// 5-point: entries before row (j,i).  Row (j,i) holds [j>0] + [i>0] + 1 + [i<n-1] + [j<n-1].
__host__ __device__ inline long long lap2d_row_start(int n, long long row) {
    const long long N = (long long)n * n;
    if (row >= N) row = N;
    const long long j = row / n, i = row % n;
    // full rows j' < j: each has sum_i (1 + [i>0] + [i<n-1]) = 3n-2 horizontal+diag, plus n*[j'>0] + n*[j'<n-1]
    long long full = j * (3LL * n - 2);
    full += (j > 0 ? (j - 1) * (long long)n : 0);                         // "up" neighbours: rows 1..j-1
    full += (j <= n - 1 ? j : (long long)n - 1) * (long long)n;           // "down" neighbours: rows 0..min(j,n-1)-1
    if (j >= n) return full;
    // partial row j: i entries before column i
    long long part = pre1(n, i) + (j > 0 ? i : 0) + (j < n - 1 ? i : 0);
    return full + part;
}

__global__ void gen_lap3d27_kernel(int n, long long row0, long long row1, int *__restrict__ rowptr,
                                   int *__restrict__ colids, double *__restrict__ values) {
    const long long base = lap3d_row_start(n, row0);
    for (long long row = row0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; row <= row1;
         row += (long long)gridDim.x * blockDim.x) {
        long long p = lap3d_row_start(n, row) - base;
        rowptr[row - row0] = (int)p;
        if (row == row1) break;
        const int i = (int)(row % n), j = (int)((row / n) % n), k = (int)(row / ((long long)n * n));
        for (int dk = -1; dk <= 1; ++dk) {
            if (k + dk < 0 || k + dk >= n) continue;
            for (int dj = -1; dj <= 1; ++dj) {
                if (j + dj < 0 || j + dj >= n) continue;
                for (int di = -1; di <= 1; ++di) {
                    if (i + di < 0 || i + di >= n) continue;
                    const long long col = ((long long)(k + dk) * n + (j + dj)) * n + (i + di);
                    colids[p] = (int)col;
                    values[p] = (col == row) ? 26.0 : -1.0;
                    ++p;
                }
            }
        }
    }
}

__global__ void gen_lap2d_kernel(int n, long long row0, long long row1, int *__restrict__ rowptr,
                                 int *__restrict__ colids, double *__restrict__ values) {
    const long long base = lap2d_row_start(n, row0);
    for (long long row = row0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; row <= row1;
         row += (long long)gridDim.x * blockDim.x) {
        long long p = lap2d_row_start(n, row) - base;
        rowptr[row - row0] = (int)p;
        if (row == row1) break;
        const int i = (int)(row % n), j = (int)(row / n);
        if (j > 0) { colids[p] = (int)(row - n); values[p++] = -1.0; }
        if (i > 0) { colids[p] = (int)(row - 1); values[p++] = -1.0; }
        colids[p] = (int)row; values[p++] = 4.0;
        if (i < n - 1) { colids[p] = (int)(row + 1); values[p++] = -1.0; }
        if (j < n - 1) { colids[p] = (int)(row + n); values[p++] = -1.0; }
    }
}

long long lap3d_nnz_range(int n, long long row0, long long row1) {
    return lap3d_row_start(n, row1) - lap3d_row_start(n, row0);
}
long long lap2d_nnz_range(int n, long long row0, long long row1) {
    return lap2d_row_start(n, row1) - lap2d_row_start(n, row0);
}

int alloc_csr(g4s_csr **out, int rows, int cols, long long nnz);  // capi.cu

static int gen_common(g4s_csr_t *out, int n, long long row0, long long row1, bool three_d, cudaStream_t stream) {
    if (!out || n < 1) return fail(G4S_ERR_INVALID, "generator: bad arguments");
    const long long N = three_d ? (long long)n * n * n : (long long)n * n;
    if (row1 < 0) row1 = N;
    if (row0 < 0 || row0 > row1 || row1 > N) return fail(G4S_ERR_INVALID, "generator: bad row range");
    if (N > 2147483647LL) return fail(G4S_ERR_INVALID, "generator: more than 2^31-1 columns");
    const long long nnz = three_d ? lap3d_nnz_range(n, row0, row1) : lap2d_nnz_range(n, row0, row1);
    if (nnz > 2147483647LL) return fail(G4S_ERR_INVALID, "generator: nnz does not fit int32 row pointers");
    int rc = ensure_device();
    if (rc) return rc;
    g4s_csr *h = nullptr;
    rc = alloc_csr(&h, (int)(row1 - row0), (int)N, nnz);
    if (rc) return rc;
    const int threads = 256;
    const long long want = (row1 - row0 + 1 + threads - 1) / threads;
    const int grid = (int)std::min<long long>(want, (long long)sm_count() * 32);
    if (three_d) gen_lap3d27_kernel<<<grid, threads, 0, stream>>>(n, row0, row1, h->rowptr, h->colids, h->values);
    else gen_lap2d_kernel<<<grid, threads, 0, stream>>>(n, row0, row1, h->rowptr, h->colids, h->values);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g4s_csr_destroy(h);
        return fail(G4S_ERR_CUDA, std::string("generator launch: ") + cudaGetErrorString(e));
    }
    count_launch();
    *out = h;
    return G4S_OK;
}

}  // namespace g4s

extern "C" {
long long g4s_laplacian2d_nnz(int n, long long row0, long long row1) {
    if (row1 < 0) row1 = (long long)n * n;
    return g4s::lap2d_nnz_range(n, row0, row1);
}
long long g4s_laplacian3d27_nnz(int n, long long row0, long long row1) {
    if (row1 < 0) row1 = (long long)n * n * n;
    return g4s::lap3d_nnz_range(n, row0, row1);
}
int g4s_csr_generate_laplacian2d(g4s_csr_t *out, int n, long long row0, long long row1, void *stream) {
    return g4s::gen_common(out, n, row0, row1, false, (cudaStream_t)stream);
}
int g4s_csr_generate_laplacian3d27(g4s_csr_t *out, int n, long long row0, long long row1, void *stream) {
    return g4s::gen_common(out, n, row0, row1, true, (cudaStream_t)stream);
}
}
