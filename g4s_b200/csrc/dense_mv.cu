// The reference's four dense mv entry points (mv/mv.c:6-27) on the GPU.
//
// The reference passes a row-major-filled dim x dim buffer to column-major CBLAS calls; element (i,j) of the
// matrix the BLAS sees is A[i + j*dim] (full storage) or A[i + j(j+1)/2], i <= j (packed-upper view of the
// same buffer, dspmv).  Each routine is HBM-bound and reads every stored element it needs exactly ONCE:
//   dgemv : row_part — thread per row i, coalesced along i for a fixed column j: sum_j a(i,j) x[j]
//                                                       (mv/mv.c:23-27: ColMajor, NoTrans, alpha 1, beta 0)
//   dsymv : upper_tile — one pass over the upper triangle; element a(i,j), i <= j, feeds BOTH y[i] += a x[j] and (i < j)
//           y[j] += a x[i]                              (mv/mv.c:6-10 : ColMajor, Upper)
//   dtrmv : upper_tile, column sums only: B <- sum_{i <= j} a(i,j) B[i]
//                                                       (mv/mv.c:12-15: ColMajor, Upper, Trans, NonUnit; in place)
//   dspmv : dsymv on the packed view                    (mv/mv.c:17-21: the driver calls it matrix_multiply_sspmv)
// All sums are taken in a fixed order (partial results per tile, reduced by a last kernel): deterministic, no atomics.
#include <algorithm>

#include "common.cuh"

namespace g4s {

__device__ __forceinline__ size_t col_offset(int j, int dim, bool packed) {
    return packed ? (size_t)j * (j + 1) / 2 : (size_t)j * dim;
}

// partial[c][i] = sum over this CTA-column-chunk c of a(i,j) x[j], j restricted to j >= i when upper
template <bool UPPER, bool PACKED>
__global__ void __launch_bounds__(256) row_part_kernel(const double *__restrict__ A, const double *__restrict__ x,
                                                       double *__restrict__ partial, int dim, int jchunk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * jchunk, j1 = min(dim, j0 + jchunk);
    __shared__ double xs[256];
    double acc = 0.0;
    for (int jb = j0; jb < j1; jb += 256) {
        __syncthreads();
        if (jb + threadIdx.x < j1) xs[threadIdx.x] = x[jb + threadIdx.x];
        __syncthreads();
        const int je = min(256, j1 - jb);
        if (i < dim) {
            int js = 0;
            if (UPPER) js = max(0, i - jb);
#pragma unroll 4
            for (int jj = js; jj < je; ++jj) acc = fma(__ldg(A + col_offset(jb + jj, dim, PACKED) + i), xs[jj], acc);
        }
    }
    if (i < dim) partial[(size_t)blockIdx.y * dim + i] = acc;
}

// out[j] = sum_{i < j + incl} a(i,j) x[i]; one warp per column
template <bool PACKED>
__global__ void __launch_bounds__(256) col_dot_kernel(const double *__restrict__ A, const double *__restrict__ x,
                                                      double *__restrict__ out, int dim, int incl) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= dim) return;
    const int lane = threadIdx.x & 31;
    const double *col = A + col_offset(j, dim, PACKED);
    const int lim = j + incl;
    double acc = 0.0;
    for (int i = lane; i < lim; i += 32) acc = fma(__ldg(col + i), __ldg(x + i), acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[j] = acc;
}

// One pass over the upper triangle.  CTA = 8 warps = 256 rows [256 rb, +256) x one chunk of columns [jc0, jc1); lane = row.
// Columns are taken 32 at a time: 32 coalesced loads a[c] = a(i, j0 + c) per lane, then
//   rows   : racc += a[c] x[j0 + c]                       for j >= i   (kept in a register over the whole chunk)
//   columns: v[c]  = a[c] x[i]  for i < j (i <= j for dtrmv), summed over the warp's 32 rows by a halving exchange
//            (31 shuffles for 32 columns; lane c ends with column j0 + c), then over the CTA's 8 warps in shared memory.
// part_row[chunk][i] and part_col[rb][j] are summed by reduce_sym_kernel.  Tiles below the diagonal are skipped.
template <bool PACKED, bool ROWS, bool DIAG_IN_COLS>
__global__ void __launch_bounds__(256) upper_tile_kernel(const double *__restrict__ A, const double *__restrict__ x,
                                                         double *__restrict__ part_row, double *__restrict__ part_col,
                                                         int dim, int jchunk) {
    constexpr unsigned FULL = 0xffffffffu;
    const int rb = blockIdx.x, chunk = blockIdx.y;
    const int jc0 = chunk * jchunk, jc1 = min(dim, jc0 + jchunk);
    if (jc1 <= rb * 256) return;  // wholly below the diagonal
    __shared__ double cs[8][32];
    __shared__ double xs[256];  // x over the tile's columns (tiles are 256 wide unless dim is huge: then the tail reads global)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (ROWS) {
        if (jc0 + (int)threadIdx.x < jc1) xs[threadIdx.x] = __ldg(x + jc0 + threadIdx.x);
        __syncthreads();
    }
    const int i = rb * 256 + threadIdx.x;
    const bool row_ok = i < dim;
    const double xi = row_ok ? __ldg(x + i) : 0.0;
    double racc = 0.0;
    for (int j0 = max(jc0, (rb * 256) & ~31); j0 < jc1; j0 += 32) {
        double v[32];
        // all 32 loads of the step first (32 independent 256-byte requests per warp in flight), arithmetic afterwards
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int j = j0 + c;
            v[c] = (row_ok && j < jc1 && i <= j) ? __ldg(A + col_offset(j, dim, PACKED) + i) : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int j = j0 + c;
            // x[j] is the same for the whole warp: a shared-memory broadcast (a shuffle per column made dsymv 70 % slower than
            // dtrmv, which needs none)
            if (ROWS && j < jc1) racc = fma(v[c], j - jc0 < 256 ? xs[j - jc0] : __ldg(x + j), racc);
            v[c] = (DIAG_IN_COLS || i < j) ? v[c] * xi : 0.0;
        }
        // halving exchange: after the step with offset o a lane keeps the columns whose bit o equals its own
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int c = 0; c < o; ++c) {
                const double send = up ? v[c] : v[c + o], keep = up ? v[c + o] : v[c];
                v[c] = keep + __shfl_xor_sync(FULL, send, o);
            }
        }
        // lane L now holds the sum over the warp's rows of column j0 + L (bits consumed high to low = its own index)
        cs[warp][lane] = v[0];
        __syncthreads();
        if (warp == 0 && j0 + lane < jc1) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += cs[w][lane];
            part_col[(size_t)rb * dim + j0 + lane] = t;
        }
        __syncthreads();
    }
    if (ROWS && row_ok) part_row[(size_t)chunk * dim + i] = racc;
}

// out[i] = sum_chunk part_row[chunk][i] + sum_rb part_col[rb][i]; entries no tile wrote are zero (the buffers are cleared)
__global__ void reduce_sym_kernel(const double *__restrict__ part_row, const double *__restrict__ part_col,
                                  double *__restrict__ out, int dim, int nchunks, int nrb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    double s = 0.0;
    if (part_row)
        for (int c = 0; c < nchunks; ++c) s += part_row[(size_t)c * dim + i];
    for (int r = 0; r < nrb; ++r) s += part_col[(size_t)r * dim + i];
    out[i] = s;
}

// out[i] = sum_c partial[c][i] (+ extra[i])
__global__ void reduce_partials_kernel(const double *__restrict__ partial, const double *__restrict__ extra,
                                       double *__restrict__ out, int dim, int nchunks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(size_t)c * dim + i];
    if (extra) s += extra[i];
    out[i] = s;
}

int dense_mv_run(int op, const double *A, double *B, double *C, int dim, cudaStream_t stream) {
    if (dim <= 0) return G4S_OK;
    const int row_blocks = (dim + 255) / 256;
    // enough column chunks to fill the machine with 256-row CTAs
    int nchunks = std::max(1, std::min((dim + 255) / 256, (sm_count() * 8 + row_blocks - 1) / row_blocks));
    const int jchunk = (((dim + nchunks - 1) / nchunks) + 255) / 256 * 256;
    nchunks = (dim + jchunk - 1) / jchunk;
    double *partial = nullptr, *tmp = nullptr;
    G4S_CUDA(cudaMallocAsync(&partial, sizeof(double) * (size_t)nchunks * dim, stream));
    G4S_CUDA(cudaMallocAsync(&tmp, sizeof(double) * (size_t)dim, stream));
    const dim3 grid(row_blocks, nchunks);
    const int col_blocks = (dim + 7) / 8;
    switch (op) {
        case 0:  // dgemv
            row_part_kernel<false, false><<<grid, 256, 0, stream>>>(A, B, partial, dim, jchunk);
            G4S_CHECK_LAUNCH("row_part_kernel");
            reduce_partials_kernel<<<row_blocks, 256, 0, stream>>>(partial, nullptr, C, dim, nchunks);
            G4S_CHECK_LAUNCH("reduce_partials_kernel");
            break;
        case 1:    // dsymv
        case 3:    // dspmv
        case 2: {  // dtrmv: B <- U^T B, through a temporary because every column reads the old B
            const int nrb = row_blocks;
            // square 256 x 256 tiles (the triangle is then ~ nrb^2 / 2 CTAs of equal work, diagonal tiles half of it); for very
            // large dim wider tiles keep the two partial buffers below ~64 chunks x dim doubles
            const int sj = std::max(256, ((dim + 63) / 64 + 255) / 256 * 256);
            const int sch = (dim + sj - 1) / sj;
            double *prow = nullptr, *pcol = nullptr;
            G4S_CUDA(cudaMallocAsync(&prow, sizeof(double) * (size_t)sch * dim, stream));
            G4S_CUDA(cudaMallocAsync(&pcol, sizeof(double) * (size_t)nrb * dim, stream));
            G4S_CUDA(cudaMemsetAsync(prow, 0, sizeof(double) * (size_t)sch * dim, stream));
            G4S_CUDA(cudaMemsetAsync(pcol, 0, sizeof(double) * (size_t)nrb * dim, stream));
            const dim3 sgrid(nrb, sch);
            if (op == 1) upper_tile_kernel<false, true, false><<<sgrid, 256, 0, stream>>>(A, B, prow, pcol, dim, sj);
            else if (op == 3) upper_tile_kernel<true, true, false><<<sgrid, 256, 0, stream>>>(A, B, prow, pcol, dim, sj);
            else upper_tile_kernel<false, false, true><<<sgrid, 256, 0, stream>>>(A, B, nullptr, pcol, dim, sj);
            G4S_CHECK_LAUNCH("upper_tile_kernel");
            reduce_sym_kernel<<<row_blocks, 256, 0, stream>>>(op == 2 ? nullptr : prow, pcol, op == 2 ? B : C, dim, sch, nrb);
            G4S_CHECK_LAUNCH("reduce_sym_kernel");
            G4S_CUDA(cudaFreeAsync(prow, stream));
            G4S_CUDA(cudaFreeAsync(pcol, stream));
            break;
        }
        default:
            cudaFreeAsync(partial, stream);
            cudaFreeAsync(tmp, stream);
            return fail(G4S_ERR_INVALID, "g4s_dense_mv_device: op must be 0..3");
    }
    G4S_CUDA(cudaFreeAsync(partial, stream));
    G4S_CUDA(cudaFreeAsync(tmp, stream));
    return G4S_OK;
}

// host-buffer wrapper shared by the four reference-named entry points
static void dense_mv_host(const char *name, int op, double *A, double *B, double *C, int dim) {
    auto run = [&]() -> int {
        if (!A || !B || (!C && op != 2) || dim < 0) return fail(G4S_ERR_INVALID, "null argument");
        if (dim == 0) return G4S_OK;
        int rc = ensure_device();
        if (rc) return rc;
        const size_t n2 = (size_t)dim * dim;
        double *dA = nullptr, *dB = nullptr, *dC = nullptr;
        G4S_CUDA(cudaMalloc(&dA, sizeof(double) * n2));
        G4S_CUDA(cudaMalloc(&dB, sizeof(double) * dim));
        G4S_CUDA(cudaMalloc(&dC, sizeof(double) * dim));
        G4S_CUDA(cudaMemcpy(dA, A, sizeof(double) * n2, cudaMemcpyHostToDevice));
        G4S_CUDA(cudaMemcpy(dB, B, sizeof(double) * dim, cudaMemcpyHostToDevice));
        rc = dense_mv_run(op, dA, dB, dC, dim, 0);
        if (rc == G4S_OK) {
            if (op == 2) G4S_CUDA(cudaMemcpy(B, dB, sizeof(double) * dim, cudaMemcpyDeviceToHost));
            else G4S_CUDA(cudaMemcpy(C, dC, sizeof(double) * dim, cudaMemcpyDeviceToHost));
        }
        cudaFree(dA);
        cudaFree(dB);
        cudaFree(dC);
        return rc;
    };
    const int rc = run();
    if (rc != G4S_OK) fprintf(stderr, "g4s_b200: %s failed (%d): %s\n", name, rc, g4s_last_error());
}

}  // namespace g4s

using namespace g4s;

extern "C" {

void matrix_multiply_dgemv(double *A, double *B, double *C, int dim) { dense_mv_host("matrix_multiply_dgemv", 0, A, B, C, dim); }
void matrix_multiply_dsymv(double *A, double *B, double *C, int dim) { dense_mv_host("matrix_multiply_dsymv", 1, A, B, C, dim); }
void matrix_multiply_dtrmv(double *A, double *B, double *C, int dim) { dense_mv_host("matrix_multiply_dtrmv", 2, A, B, C, dim); }
void matrix_multiply_sspmv(double *A, double *B, double *C, int dim) { dense_mv_host("matrix_multiply_sspmv", 3, A, B, C, dim); }

int g4s_dense_mv_device(int op, const double *A_dev, double *B_dev, double *C_dev, int dim, void *stream) {
    if (!A_dev || !B_dev || (!C_dev && op != 2) || dim < 0) return fail(G4S_ERR_INVALID, "g4s_dense_mv_device: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    return dense_mv_run(op, A_dev, B_dev, C_dev, dim, (cudaStream_t)stream);
}

}  // extern "C"
