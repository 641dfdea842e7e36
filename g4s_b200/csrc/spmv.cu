// CSR SpMV for sm_100a:  y = A x  (fp64 values, int32 indices).
//
// Replaces, for sparse A, the reference's dense matrix_multiply_dgemv (mv/mv.c:23-27; alpha=1, beta=0) on
// the reference's own CSR container (mm/inc/CSR.h:22-100).
//
// Design:
//  * Inspector (once per matrix, like mkl_sparse_optimize): the merge list "row ends ∪ nnz indices" is cut
//    into tiles of TILE items; tile t starts at row tile_row[t] and nnz t*TILE - tile_row[t] (merge-path
//    diagonal search).  Every tile therefore owns at most TILE rows AND at most TILE nonzeros, whatever the
//    row-length distribution (power-law rows, millions of empty rows).
//  * Executor: persistent CTAs; each walks tiles blockIdx.x, +gridDim.x, ...  A tile's colids / values /
//    rowptr slices are contiguous in HBM and are brought into shared memory by the TMA engine as 1-D bulk
//    async copies (cp.async.bulk ... mbarrier::complete_tx, L2 evict_first) through a STAGES-deep ring, so
//    the HBM stream never waits for arithmetic.  x is gathered with ld.global.nc and lives in L1/L2.
//  * Inside a tile, L = 2^k lanes cooperate on a row (L from the tile's mean row length), reading the
//    staged arrays from shared memory (conflict-free for odd row lengths) and reducing with shuffles;
//    rows far longer than the mean are swept by the whole CTA.
//  * A row cut by a tile boundary leaves its partial sum in carry[t]; a tiny fix-up kernel adds the carries
//    in tile order (deterministic: no atomics anywhere).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace g4s {

// ------------------------------------------------------------------------------------------------------
// Inspector: merge-path diagonal search, one thread per tile boundary.
// Item d of the merge is "row end i" if rowptr[i+1] <= d - i - 1 ... (CUB-style coordinate search).
// ------------------------------------------------------------------------------------------------------
__global__ void spmv_tile_search_kernel(const int *__restrict__ rowptr, int rows, long long nnz, int tile_items,
                                        int ntiles, int *__restrict__ tile_row) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    long long total = (long long)rows + nnz;
    long long d = (long long)t * tile_items;
    if (d > total) d = total;
    long long lo = d > nnz ? d - nnz : 0;
    long long hi = d < rows ? d : rows;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        // row `mid` ends before nnz index (d - mid - 1) is consumed?
        if ((long long)__ldg(rowptr + mid + 1) <= d - mid - 1) lo = mid + 1;
        else hi = mid;
    }
    tile_row[t] = (int)lo;
}

__global__ void max_row_len_kernel(const int *__restrict__ rowptr, int rows, int *__restrict__ out) {
    int m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __ldg(rowptr + i + 1) - __ldg(rowptr + i));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------------------
// Executor
// ------------------------------------------------------------------------------------------------------
template <int TILE>
struct alignas(32) SpmvStage {
    double vals[TILE + 8];
    int cols[TILE + 8];
    int rp[TILE + 16];
    int desc[8];  // row0, row1, nnz0, nnz1
};

struct SpmvArgs {
    const int *rowptr;
    const int *colids;
    const double *values;
    const double *x;
    double *y;
    const int *tile_row;
    double *carry;
    const int *row_map;  // optional: local row r is y[row_map[r]] (row-compressed off-diagonal blocks)
    int rows;
    long long nnz;
    int ntiles;
    int lanes_log2;  // -1: per-tile automatic
};

template <int L, int THREADS>
__device__ __forceinline__ double group_reduce(double v) {
#pragma unroll
    for (int o = L >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ACCUM: y[row] += sum instead of y[row] = sum (second pass of the split multi-GPU product).
template <int TILE, int THREADS, int L, bool ACCUM>
__device__ __forceinline__ void tile_rows(const SpmvStage<TILE> &st, const SpmvArgs &a, int tile, int *long_rows,
                                          int *n_long, double *red) {
    const int row0 = st.desc[0], row1 = st.desc[1], nnz0 = st.desc[2], nnz1 = st.desc[3];
    const int a0 = nnz0 & ~3;
    const int rbase = row0 - (row0 & ~3);
    const int R = row1 - row0;
    const int nloc = R + (row1 < a.rows ? 1 : 0);  // rows ending here + the row left open at the end
    constexpr int GROUPS = THREADS / L;
    const int gid = threadIdx.x / L, lane = threadIdx.x % L;
    constexpr int LONG = 64 * L;
    const double *__restrict__ x = a.x;

    for (int base = 0; base < nloc; base += GROUPS) {
        const int r = base + gid;
        int s = 0, e = 0;
        if (r < nloc) {
            s = max(st.rp[rbase + r], nnz0);
            e = min(st.rp[rbase + r + 1], nnz1);
        }
        bool is_long = (e - s) > LONG;
        if (is_long) {
            if (lane == 0) long_rows[atomicAdd(n_long, 1)] = r;
            e = s;
        }
        double acc = 0.0;
        int k = s + lane - a0;
        const int ke = e - a0;
#pragma unroll 4
        for (; k < ke; k += L) acc = fma(st.vals[k], __ldg(x + st.cols[k]), acc);
        if (L > 1) acc = group_reduce<L, THREADS>(acc);
        if (lane == 0 && r < nloc && !is_long) {
            if (r < R) {
                const int row = a.row_map ? a.row_map[row0 + r] : row0 + r;
                if (ACCUM) a.y[row] += acc;
                else a.y[row] = acc;
            } else {
                a.carry[tile] = acc;
            }
        }
    }
    // rows far above the tile's mean length: the whole CTA sweeps each one
    __syncthreads();
    const int nl = *n_long;
    for (int i = 0; i < nl; ++i) {
        const int r = long_rows[i];
        const int s = max(st.rp[rbase + r], nnz0) - a0, e = min(st.rp[rbase + r + 1], nnz1) - a0;
        double acc = 0.0;
        for (int k = s + threadIdx.x; k < e; k += THREADS) acc = fma(st.vals[k], __ldg(x + st.cols[k]), acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) t += red[w];
            if (r < R) {
                const int row = a.row_map ? a.row_map[row0 + r] : row0 + r;
                if (ACCUM) a.y[row] += t;
                else a.y[row] = t;
            } else {
                a.carry[tile] = t;
            }
        }
        __syncthreads();
    }
}

template <int TILE, int STAGES, int THREADS, bool ACCUM>
__global__ void __launch_bounds__(THREADS) spmv_tile_kernel(const SpmvArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvStage<TILE> *stages = reinterpret_cast<SpmvStage<TILE> *>(smem_raw);
    __shared__ uint64_t full[STAGES];
    __shared__ int long_rows[TILE / 64 + 64];
    __shared__ int n_long;
    __shared__ double red[THREADS / 32];

    const int tid = threadIdx.x;
    const long long total_items = (long long)a.rows + a.nnz;
    uint64_t policy = 0;

    // producer (thread 0): stage one tile
    auto issue = [&](int stage, int tile) {
        SpmvStage<TILE> &st = stages[stage];
        const int row0 = __ldg(a.tile_row + tile), row1 = __ldg(a.tile_row + tile + 1);
        const long long d0 = (long long)tile * TILE;
        const long long d1 = min(d0 + TILE, total_items);
        const int nnz0 = (int)(d0 - row0), nnz1 = (int)(d1 - row1);
        st.desc[0] = row0;
        st.desc[1] = row1;
        st.desc[2] = nnz0;
        st.desc[3] = nnz1;
        // nonzeros [nnz0, nnz1): bulk-copy the 16-byte-aligned cover, scalar-copy what sticks out past the
        // last aligned element of the arrays (only ever the final tile)
        const int a0 = nnz0 & ~3;
        long long a1 = ((long long)nnz1 + 3) & ~3LL;
        const long long amax = a.nnz & ~3LL;
        if (a1 > amax) a1 = amax;
        if (a1 < a0) a1 = a0;
        const uint32_t ncopy = (uint32_t)(a1 - a0);
        for (long long k = a1; k < nnz1; ++k) {
            st.cols[k - a0] = a.colids[k];
            st.vals[k - a0] = a.values[k];
        }
        // rowptr[row0 .. min(row1+1, rows)]
        const int r0 = row0 & ~3;
        const int rend = min(row1 + 1, a.rows) + 1;  // one past the last needed entry
        long long r1 = ((long long)rend + 3) & ~3LL;
        const long long rmax = ((long long)a.rows + 1) & ~3LL;
        if (r1 > rmax) r1 = rmax;
        if (r1 < r0) r1 = r0;
        const uint32_t nrp = (uint32_t)(r1 - r0);
        for (long long k = r1; k < rend; ++k) st.rp[k - r0] = a.rowptr[k];
        mbar_arrive_expect_tx(&full[stage], ncopy * 12u + nrp * 4u);
        if (ncopy) {
            bulk_g2s(st.vals, a.values + a0, ncopy * 8u, &full[stage], policy);
            bulk_g2s(st.cols, a.colids + a0, ncopy * 4u, &full[stage], policy);
        }
        if (nrp) bulk_g2s(st.rp, a.rowptr + r0, nrp * 4u, &full[stage], policy);
    };

    if (tid == 0) {
        policy = policy_evict_first();
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        n_long = 0;
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < a.ntiles) issue(s, tile);
        }
    }

    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        mbar_wait(&full[stage], parity);
        const SpmvStage<TILE> &st = stages[stage];

        int lg = a.lanes_log2;
        if (lg < 0) {  // automatic: about 4 nonzeros per lane
            const int R1 = st.desc[1] - st.desc[0] + 1;
            const int avg = (st.desc[3] - st.desc[2]) / R1;
            lg = avg <= 6 ? 0 : avg <= 12 ? 1 : avg <= 24 ? 2 : avg <= 48 ? 3 : avg <= 96 ? 4 : 5;
        }
        switch (lg) {
            case 0: tile_rows<TILE, THREADS, 1, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
            case 1: tile_rows<TILE, THREADS, 2, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
            case 2: tile_rows<TILE, THREADS, 4, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
            case 3: tile_rows<TILE, THREADS, 8, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
            case 4: tile_rows<TILE, THREADS, 16, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
            default: tile_rows<TILE, THREADS, 32, ACCUM>(st, a, tile, long_rows, &n_long, red); break;
        }
        // tile_rows ends with every thread past its last read of the stage (it syncs before the long-row
        // sweep and after each swept row); one more barrier covers the no-long-row path
        __syncthreads();
        if (tid == 0) {
            n_long = 0;
            const int next = tile + STAGES * gridDim.x;
            if (next < a.ntiles) issue(stage, next);
        }
    }
}

// carry[t] belongs to row tile_row[t+1]; consecutive tiles inside one long row form a chain that one
// thread adds up in tile order.
__global__ void spmv_fixup_kernel(const int *__restrict__ tile_row, const double *__restrict__ carry,
                                  const int *__restrict__ row_map, double *__restrict__ y, int ntiles, int rows) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int row = tile_row[t + 1];
    if (row >= rows) return;
    if (t > 0 && tile_row[t] == row) return;  // not the head of its chain
    double s = carry[t];
    for (int u = t + 1; u < ntiles && tile_row[u + 1] == row; ++u) s += carry[u];
    y[row_map ? row_map[row] : row] += s;
}

// Baseline kernel kept for comparison and for tiny matrices: one warp per row straight from global memory.
__global__ void spmv_warp_row_kernel(const int *__restrict__ rowptr, const int *__restrict__ colids,
                                     const double *__restrict__ values, const double *__restrict__ x,
                                     double *__restrict__ y, int rows) {
    const int lane = threadIdx.x & 31;
    for (long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; row < rows;
         row += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
        double acc = 0.0;
        for (int k = s + lane; k < e; k += 32) acc = fma(__ldg(values + k), __ldg(x + __ldg(colids + k)), acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[row] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
template <int TILE, int STAGES, int THREADS>
static int launch_tile_kernel(const SpmvArgs &args, bool accum, int ctas_per_sm, cudaStream_t stream) {
    const size_t smem = sizeof(SpmvStage<TILE>) * STAGES;
    auto k0 = spmv_tile_kernel<TILE, STAGES, THREADS, false>;
    auto k1 = spmv_tile_kernel<TILE, STAGES, THREADS, true>;
    static bool configured = false;
    if (!configured) {
        G4S_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        G4S_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int grid = std::min(args.ntiles, sm_count() * ctas_per_sm);
    if (grid < 1) grid = 1;
    if (accum) k1<<<grid, THREADS, smem, stream>>>(args);
    else k0<<<grid, THREADS, smem, stream>>>(args);
    G4S_CHECK_LAUNCH("spmv_tile_kernel");
    return G4S_OK;
}

int spmv_build_plan(g4s_csr *h, cudaStream_t stream) {
    SpmvPlan &p = h->plan;
    if (p.tile_row) return G4S_OK;
    p.tile_items = 2048;
    const long long items = (long long)h->rows + h->nnz;
    p.ntiles = (int)((items + p.tile_items - 1) / p.tile_items);
    if (p.ntiles < 1) p.ntiles = 1;
    G4S_CUDA(cudaMalloc(&p.tile_row, sizeof(int) * ((size_t)p.ntiles + 1)));
    G4S_CUDA(cudaMalloc(&p.carry, sizeof(double) * (size_t)p.ntiles));
    G4S_CUDA(cudaMemsetAsync(p.carry, 0, sizeof(double) * (size_t)p.ntiles, stream));
    const int threads = 256;
    spmv_tile_search_kernel<<<(p.ntiles + 1 + threads - 1) / threads, threads, 0, stream>>>(
        h->rowptr, h->rows, h->nnz, p.tile_items, p.ntiles, p.tile_row);
    G4S_CHECK_LAUNCH("spmv_tile_search_kernel");
    return G4S_OK;
}

void spmv_free_plan(g4s_csr *h) {
    if (h->plan.tile_row) cudaFree(h->plan.tile_row);
    if (h->plan.carry) cudaFree(h->plan.carry);
    h->plan = SpmvPlan();
}

int spmv_run(g4s_csr *h, const double *x, double *y, const int *row_map, bool accum, cudaStream_t stream) {
    if (h->rows == 0) return G4S_OK;
    int rc = spmv_build_plan(h, stream);
    if (rc) return rc;
    const SpmvPlan &p = h->plan;
    if (p.variant == 9) {  // comparison baseline
        if (accum || row_map) return fail(G4S_ERR_INVALID, "warp-per-row baseline supports plain y = A x only");
        int grid = sm_count() * 8;
        spmv_warp_row_kernel<<<grid, 256, 0, stream>>>(h->rowptr, h->colids, h->values, x, y, h->rows);
        G4S_CHECK_LAUNCH("spmv_warp_row_kernel");
        return G4S_OK;
    }
    SpmvArgs a;
    a.rowptr = h->rowptr;
    a.colids = h->colids;
    a.values = h->values;
    a.x = x;
    a.y = y;
    a.tile_row = p.tile_row;
    a.carry = p.carry;
    a.row_map = row_map;
    a.rows = h->rows;
    a.nnz = h->nnz;
    a.ntiles = p.ntiles;
    a.lanes_log2 = -1;
    if (p.lanes_per_row > 0) {
        int lg = 0;
        while ((1 << lg) < p.lanes_per_row && lg < 5) ++lg;
        a.lanes_log2 = lg;
    }
    // variants: ring depth x CTAs per SM (TILE is fixed by the plan)
    switch (p.variant) {
        case 1: rc = launch_tile_kernel<2048, 2, 256>(a, accum, 3, stream); break;
        case 2: rc = launch_tile_kernel<2048, 4, 256>(a, accum, 1, stream); break;
        case 3: rc = launch_tile_kernel<2048, 3, 128>(a, accum, 2, stream); break;
        case 4: rc = launch_tile_kernel<2048, 2, 512>(a, accum, 2, stream); break;
        default: rc = launch_tile_kernel<2048, 3, 256>(a, accum, 2, stream); break;
    }
    if (rc) return rc;
    const int threads = 256;
    spmv_fixup_kernel<<<(p.ntiles + threads - 1) / threads, threads, 0, stream>>>(p.tile_row, p.carry, row_map, y,
                                                                                 p.ntiles, h->rows);
    G4S_CHECK_LAUNCH("spmv_fixup_kernel");
    return G4S_OK;
}

}  // namespace g4s
