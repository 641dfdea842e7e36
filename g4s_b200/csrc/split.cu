// Building blocks of the row-partitioned multi-GPU product:
//   g4s_csr_split_columns   : a GPU's row block A[r0:r1, :] -> diagonal block (columns it owns, rebased) and
//                             off-diagonal block (all other columns), the latter ROW-COMPRESSED: only rows that
//                             have an off-diagonal entry are kept, with a row map back to local rows.
//   g4s_csr_compact_columns : the sorted list of distinct columns an (off-diagonal) block references — the
//                             entries of x that must come from other GPUs — and a rewrite of its column ids to
//                             positions in that list (the halo buffer).
//   g4s_gather_f64          : dst[k] = src[idx[k]] — packs the x entries other GPUs asked for.
// The row cut itself is g4s_partition_rows_* (BIN::set_rows_offset, mm/inc/BIN.h:100-122, on nnz or work).
#include <algorithm>

#include "common.cuh"

namespace g4s {

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host,
                       cudaStream_t stream);
int alloc_csr(g4s_csr **out, int rows, int cols, long long nnz);

// per row: entries inside / outside the column window [c0, c1)
__global__ void split_count_kernel(const int *__restrict__ rowptr, const int *__restrict__ colids, int rows, int c0,
                                   int c1, int *__restrict__ in_cnt, int *__restrict__ out_cnt,
                                   int *__restrict__ has_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    int ni = 0;
    const int s = rowptr[warp], e = rowptr[warp + 1];
    for (int k = s + lane; k < e; k += 32) {
        const int c = __ldg(colids + k);
        ni += (c >= c0 && c < c1);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ni += __shfl_xor_sync(0xffffffffu, ni, o);
    if (lane == 0) {
        in_cnt[warp] = ni;
        out_cnt[warp] = (e - s) - ni;
        has_out[warp] = (e - s) > ni;
    }
}

// One warp per row, order-preserving: lanes ballot which entries go where.
__global__ void split_fill_kernel(const int *__restrict__ rowptr, const int *__restrict__ colids,
                                  const double *__restrict__ values, int rows, int c0, int c1,
                                  const int *__restrict__ in_ptr, const int *__restrict__ out_slot,
                                  const int *__restrict__ out_cnt, int *__restrict__ d_col, double *__restrict__ d_val,
                                  int *__restrict__ o_rowcnt, int *__restrict__ o_rowmap,
                                  const int *__restrict__ o_ptr, int *__restrict__ o_col, double *__restrict__ o_val,
                                  int pass) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int s = rowptr[warp], e = rowptr[warp + 1];
    const bool keep_out = out_cnt[warp] > 0;
    const int slot = out_slot[warp];  // compressed row index of this row in the off-diagonal block
    if (pass == 0) {                  // publish the compressed rows' lengths and map
        if (lane == 0 && keep_out) {
            o_rowcnt[slot] = out_cnt[warp];
            o_rowmap[slot] = warp;
        }
        return;
    }
    int di = in_ptr[warp], oi = keep_out ? o_ptr[slot] : 0;
    for (int base = s; base < e; base += 32) {
        const int k = base + lane;
        int c = -1;
        double v = 0.0;
        if (k < e) {
            c = __ldg(colids + k);
            v = __ldg(values + k);
        }
        const bool inside = k < e && c >= c0 && c < c1;
        const bool outside = k < e && !inside;
        const unsigned mi = __ballot_sync(0xffffffffu, inside), mo = __ballot_sync(0xffffffffu, outside);
        const unsigned below = (1u << lane) - 1;
        if (inside) {
            const int p = di + __popc(mi & below);
            d_col[p] = c - c0;
            d_val[p] = v;
        } else if (outside) {
            const int p = oi + __popc(mo & below);
            o_col[p] = c;
            o_val[p] = v;
        }
        di += __popc(mi);
        oi += __popc(mo);
    }
}

__global__ void flag_columns_kernel(const int *__restrict__ colids, long long nnz, int *__restrict__ flags) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
        flags[__ldg(colids + k)] = 1;
}
__global__ void list_columns_kernel(const int *__restrict__ flags, const int *__restrict__ pos, int cols,
                                    int *__restrict__ list) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cols && flags[c]) list[pos[c]] = c;
}
__global__ void remap_columns_kernel(int *__restrict__ colids, long long nnz, const int *__restrict__ pos) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
        colids[k] = pos[colids[k]];
}
__global__ void gather_f64_kernel(double *__restrict__ dst, const double *__restrict__ src,
                                  const int *__restrict__ idx, long long n) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        dst[k] = __ldg(src + __ldg(idx + k));
}

}  // namespace g4s

using namespace g4s;

extern "C" {

int g4s_csr_split_columns(g4s_csr_t A, int c0, int c1, g4s_csr_t *diag, g4s_csr_t *off, void *stream_) {
    if (!A || !diag || !off || c0 < 0 || c1 < c0 || c1 > A->cols)
        return fail(G4S_ERR_INVALID, "g4s_csr_split_columns: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rows = A->rows;
    int *in_cnt = nullptr, *out_cnt = nullptr, *has_out = nullptr, *o_cnt = nullptr;
    const size_t n1 = (size_t)rows + 1;
    G4S_CUDA(cudaMalloc(&in_cnt, sizeof(int) * n1));
    G4S_CUDA(cudaMalloc(&out_cnt, sizeof(int) * n1));
    G4S_CUDA(cudaMalloc(&has_out, sizeof(int) * n1));
    const int threads = 256;
    const int blocks = (int)(((long long)rows * 32 + threads - 1) / threads);
    if (rows) {
        split_count_kernel<<<blocks, threads, 0, stream>>>(A->rowptr, A->colids, rows, c0, c1, in_cnt, out_cnt, has_out);
        G4S_CHECK_LAUNCH("split_count_kernel");
    }
    long long nnz_in = 0, n_off_rows = 0, nnz_out = 0;
    int *in_ptr = nullptr, *slot = nullptr;
    G4S_CUDA(cudaMalloc(&in_ptr, sizeof(int) * n1));
    G4S_CUDA(cudaMalloc(&slot, sizeof(int) * n1));
    if ((rc = exclusive_scan_i32(in_cnt, in_ptr, rows, 1, &nnz_in, stream))) return rc;
    if ((rc = exclusive_scan_i32(has_out, slot, rows, 1, &n_off_rows, stream))) return rc;
    g4s_csr *D = nullptr, *O = nullptr;
    if ((rc = alloc_csr(&D, rows, c1 - c0, nnz_in))) return rc;
    G4S_CUDA(cudaMemcpyAsync(D->rowptr, in_ptr, sizeof(int) * n1, cudaMemcpyDeviceToDevice, stream));
    // off-diagonal block: n_off_rows compressed rows
    G4S_CUDA(cudaMalloc(&o_cnt, sizeof(int) * ((size_t)n_off_rows + 1)));
    int *row_map = nullptr;
    G4S_CUDA(cudaMalloc(&row_map, sizeof(int) * (size_t)std::max<long long>(n_off_rows, 1)));
    if (rows) {
        split_fill_kernel<<<blocks, threads, 0, stream>>>(A->rowptr, A->colids, A->values, rows, c0, c1, in_ptr, slot,
                                                         out_cnt, nullptr, nullptr, o_cnt, row_map, nullptr, nullptr,
                                                         nullptr, 0);
        G4S_CHECK_LAUNCH("split_fill_kernel<map>");
    }
    int *o_ptr = nullptr;
    G4S_CUDA(cudaMalloc(&o_ptr, sizeof(int) * ((size_t)n_off_rows + 1)));
    if ((rc = exclusive_scan_i32(o_cnt, o_ptr, n_off_rows, 1, &nnz_out, stream))) return rc;
    if ((rc = alloc_csr(&O, (int)n_off_rows, A->cols, nnz_out))) return rc;
    G4S_CUDA(cudaMemcpyAsync(O->rowptr, o_ptr, sizeof(int) * ((size_t)n_off_rows + 1), cudaMemcpyDeviceToDevice, stream));
    if (rows) {
        split_fill_kernel<<<blocks, threads, 0, stream>>>(A->rowptr, A->colids, A->values, rows, c0, c1, in_ptr, slot,
                                                         out_cnt, D->colids, D->values, o_cnt, row_map, O->rowptr,
                                                         O->colids, O->values, 1);
        G4S_CHECK_LAUNCH("split_fill_kernel<fill>");
    }
    G4S_CUDA(cudaStreamSynchronize(stream));
    O->row_map = row_map;
    O->full_rows = rows;
    cudaFree(in_cnt);
    cudaFree(out_cnt);
    cudaFree(has_out);
    cudaFree(in_ptr);
    cudaFree(slot);
    cudaFree(o_cnt);
    cudaFree(o_ptr);
    *diag = D;
    *off = O;
    return G4S_OK;
}

int g4s_csr_row_map(g4s_csr_t h, const int **row_map_dev, int *full_rows) {
    if (!h) return fail(G4S_ERR_INVALID, "null handle");
    if (row_map_dev) *row_map_dev = h->row_map;
    if (full_rows) *full_rows = h->row_map ? h->full_rows : h->rows;
    return G4S_OK;
}

int g4s_csr_compact_columns(g4s_csr_t A, int **needed_cols_dev, int *n_needed, void *stream_) {
    if (!A || !needed_cols_dev || !n_needed) return fail(G4S_ERR_INVALID, "g4s_csr_compact_columns: null argument");
    if (!A->owns) return fail(G4S_ERR_INVALID, "g4s_csr_compact_columns: handle must own its arrays (they are rewritten)");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cols = A->cols;
    int *flags = nullptr, *pos = nullptr, *list = nullptr;
    G4S_CUDA(cudaMalloc(&flags, sizeof(int) * ((size_t)cols + 1)));
    G4S_CUDA(cudaMalloc(&pos, sizeof(int) * ((size_t)cols + 1)));
    G4S_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)cols + 1), stream));
    const int grid = sm_count() * 8;
    if (A->nnz) {
        flag_columns_kernel<<<grid, 256, 0, stream>>>(A->colids, A->nnz, flags);
        G4S_CHECK_LAUNCH("flag_columns_kernel");
    }
    long long n = 0;
    if ((rc = exclusive_scan_i32(flags, pos, cols, 1, &n, stream))) return rc;
    G4S_CUDA(cudaMalloc(&list, sizeof(int) * (size_t)std::max<long long>(n, 1)));
    if (cols) {
        list_columns_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(flags, pos, cols, list);
        G4S_CHECK_LAUNCH("list_columns_kernel");
    }
    if (A->nnz) {
        remap_columns_kernel<<<grid, 256, 0, stream>>>(A->colids, A->nnz, pos);
        G4S_CHECK_LAUNCH("remap_columns_kernel");
    }
    G4S_CUDA(cudaStreamSynchronize(stream));
    cudaFree(flags);
    cudaFree(pos);
    A->cols = (int)n;
    *needed_cols_dev = list;
    *n_needed = (int)n;
    return G4S_OK;
}

int g4s_device_free(void *p) {
    if (p) G4S_CUDA(cudaFree(p));
    return G4S_OK;
}

int g4s_gather_f64(double *dst_dev, const double *src_dev, const int *idx_dev, long long n, void *stream) {
    if (n < 0 || (n > 0 && (!dst_dev || !src_dev || !idx_dev))) return fail(G4S_ERR_INVALID, "g4s_gather_f64: bad arguments");
    if (n == 0) return G4S_OK;
    int rc = ensure_device();
    if (rc) return rc;
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
    gather_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dst_dev, src_dev, idx_dev, n);
    G4S_CHECK_LAUNCH("gather_f64_kernel");
    return G4S_OK;
}

}  // extern "C"
