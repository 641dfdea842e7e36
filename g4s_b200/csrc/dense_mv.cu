// The reference's four dense mv entry points (mv/mv.c:6-27) on the GPU.
//
// The reference passes a row-major-filled dim x dim buffer to column-major CBLAS calls; element (i,j) of the
// matrix the BLAS sees is A[i + j*dim] (full storage) or A[i + j(j+1)/2], i <= j (packed-upper view of the
// same buffer, dspmv).  Each routine is HBM-bound (dim^2 or dim^2/2 doubles read once or twice), built from
// two deterministic primitives:
//   row_part : thread per row i, coalesced along i for a fixed column j:  sum_{j in [jlo(i), jhi)} a(i,j) x[j]
//   col_dot  : warp per column j, coalesced along the column:             sum_{i <  lim(j)} a(i,j) x[i]
//   dgemv : C = row_part(all j)                         (mv/mv.c:23-27: ColMajor, NoTrans, alpha 1, beta 0)
//   dsymv : C = row_part(j >= i) + col_dot(i < j)       (mv/mv.c:6-10 : ColMajor, Upper)
//   dtrmv : B <- col_dot(i <= j)                        (mv/mv.c:12-15: ColMajor, Upper, Trans, NonUnit; in place)
//   dspmv : dsymv on the packed view                    (mv/mv.c:17-21: the driver calls it matrix_multiply_sspmv)
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
        case 1:  // dsymv
        case 3:  // dspmv
            if (op == 1) row_part_kernel<true, false><<<grid, 256, 0, stream>>>(A, B, partial, dim, jchunk);
            else row_part_kernel<true, true><<<grid, 256, 0, stream>>>(A, B, partial, dim, jchunk);
            G4S_CHECK_LAUNCH("row_part_kernel");
            if (op == 1) col_dot_kernel<false><<<col_blocks, 256, 0, stream>>>(A, B, tmp, dim, 0);
            else col_dot_kernel<true><<<col_blocks, 256, 0, stream>>>(A, B, tmp, dim, 0);
            G4S_CHECK_LAUNCH("col_dot_kernel");
            reduce_partials_kernel<<<row_blocks, 256, 0, stream>>>(partial, tmp, C, dim, nchunks);
            G4S_CHECK_LAUNCH("reduce_partials_kernel");
            break;
        case 2:  // dtrmv: B <- U^T B, through a temporary because every column reads the old B
            col_dot_kernel<false><<<col_blocks, 256, 0, stream>>>(A, B, tmp, dim, 1);
            G4S_CHECK_LAUNCH("col_dot_kernel");
            G4S_CUDA(cudaMemcpyAsync(B, tmp, sizeof(double) * (size_t)dim, cudaMemcpyDeviceToDevice, stream));
            break;
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
