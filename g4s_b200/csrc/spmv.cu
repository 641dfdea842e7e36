// CSR SpMV for sm_100a:  y = A x  (fp64 values, int32 indices).
//
// Replaces, for sparse A, the reference's dense matrix_multiply_dgemv (mv/mv.c:23-27; alpha=1, beta=0) on
// the reference's own CSR container (mm/inc/CSR.h:22-100).
//
// Design:
//  * Inspector (once per matrix, like mkl_sparse_optimize): rows are grouped into CHUNKS of at most 32 rows
//    and at most CAP nonzeros; a row longer than CAP is cut into CAP-sized pieces, one chunk each.  A chunk
//    is the pair (first row, first nnz); chunk c ends where chunk c+1 starts.  For every chunk the inspector
//    also picks L = 2^k lanes per row from the chunk's own row lengths (cost model below), so short regular
//    rows run lane-per-row and skewed (power-law) chunks spread a long row over many lanes.
//  * Executor: persistent CTAs of autonomous warps.  A warp owns a chunk at a time: lane 0 asks the TMA
//    engine for the chunk's colids/values slices (two 1-D bulk async copies, cp.async.bulk + mbarrier
//    complete_tx, L2 evict_first) into the warp's private shared-memory ring, the warp waits on its own
//    mbarrier, multiplies out of shared memory (bank-conflict-free for odd row lengths), and stores y
//    coalesced.  No CTA-wide barrier exists; NBUF chunks per warp are in flight so HBM never waits for
//    arithmetic.  x is gathered with ld.global.nc and is served by L1/L2.
//  * Pieces of a long row leave partial sums in carry[chunk]; one thread per long row adds them in order
//    (deterministic; no atomics anywhere).
#include <algorithm>
#include <chrono>
#include <vector>

#include "common.cuh"

namespace g4s {

constexpr int CHUNK_ROWS = 32;       // rows per chunk for ordinary matrices (one row per lane)
constexpr int MAX_CHUNK_ROWS = 128;  // regular short-row matrices pack up to 4 passes of 32 rows into a chunk

// ------------------------------------------------------------------------------------------------------
// small device-wide exclusive scan (int32), three launches; also used by SpGEMM for C's row pointers
// ------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *warp_sums) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const int base = w ? warp_sums[w - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    return base + incl - v;
}

__global__ void scan_tile_sums_kernel(const int *__restrict__ in, long long n, long long *__restrict__ tile_sums) {
    __shared__ int ws[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        long long k = base + (long long)i * SCAN_THREADS + threadIdx.x;
        if (k < n) s += in[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += ws[w];
        tile_sums[blockIdx.x] = t;
    }
}
// single block: exclusive scan of tile sums in place, grand total to tile_sums[ntiles]
__global__ void scan_tile_offsets_kernel(long long *tile_sums, int ntiles) {
    __shared__ long long carry;
    __shared__ long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long v = i < ntiles ? tile_sums[i] : 0;
        long long incl = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            long long s = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const long long off = carry + (w ? wsum[w - 1] : 0) + incl - v;
        if (i < ntiles) tile_sums[i] = off;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = off + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}
// out[k] = exclusive prefix of in (int32 result; caller guarantees the total fits), out may alias in.
// writes n+1 entries when write_total.
__global__ void scan_apply_kernel(const int *__restrict__ in, int *__restrict__ out, long long n,
                                  const long long *__restrict__ tile_offsets, int write_total) {
    __shared__ int ws[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int off = block_exclusive_scan(s, nullptr, ws) + (int)tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = off;
        off += v[i];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        out[n] = (int)tile_offsets[gridDim.x];
}

// exclusive scan of n int32 counts; total (64-bit) returned through *total_host when non-null (synchronises).
// Scratch for the tile sums, kept per (thread, device, stream): a scan is part of every SpGEMM product and of every radix-sort
// pass, and a cudaMallocAsync / cudaFreeAsync pair per call was two of the ~16 runtime calls of a repeated product.
struct ScanScratch {
    long long *buf = nullptr;
    size_t cap = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
};
static thread_local ScanScratch t_scan;

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host,
                       cudaStream_t stream) {
    const int ntiles = (int)std::max<long long>(1, (n + SCAN_TILE - 1) / SCAN_TILE);
    int dev = 0;
    cudaGetDevice(&dev);
    ScanScratch &sc = t_scan;
    if (sc.device != dev || sc.stream != stream || sc.cap < (size_t)ntiles + 1) {
        // another device / stream / size: the old buffer may still be in use by work queued on its stream, so it is released
        // in that stream's order and a new one is taken in this stream's order
        if (sc.buf && sc.device == dev) cudaFreeAsync(sc.buf, sc.stream);
        sc.buf = nullptr;
        sc.cap = 0;
        const size_t want = std::max<size_t>((size_t)ntiles + 1, 4096);
        G4S_CUDA(cudaMallocAsync(&sc.buf, sizeof(long long) * want, stream));
        sc.cap = want;
        sc.device = dev;
        sc.stream = stream;
    }
    long long *tile_sums = sc.buf;
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, stream>>>(in, n, tile_sums);
    G4S_CHECK_LAUNCH("scan_tile_sums_kernel");
    scan_tile_offsets_kernel<<<1, 1024, 0, stream>>>(tile_sums, ntiles);
    G4S_CHECK_LAUNCH("scan_tile_offsets_kernel");
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, stream>>>(in, out, n, tile_sums, write_total);
    G4S_CHECK_LAUNCH("scan_apply_kernel");
    if (total_host) {
        G4S_CUDA(cudaMemcpyAsync(total_host, tile_sums + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, stream));
        G4S_CUDA(cudaStreamSynchronize(stream));
    }
    return G4S_OK;
}

// ------------------------------------------------------------------------------------------------------
// Inspector.  One thread walks one block of `block_rows` (32, 64 or 128) consecutive rows and cuts it greedily:
//   - rows are appended to the open chunk while its nonzeros stay <= cap;
//   - a row longer than cap closes the open chunk and becomes ceil(len/cap) single-piece chunks.
// Pass 1 counts chunks per block, a scan turns counts into offsets, pass 2 writes the descriptors.
// ------------------------------------------------------------------------------------------------------
struct ChunkWalk {
    int cap;
    // cost model (in "inner-loop iterations") for running a chunk with L = 1<<lg lanes per row
    __device__ static int lanes_for(const int *len, int nrows) {
        int best = 0, best_cost = 0x7fffffff;
        for (int lg = 0; lg <= 5; ++lg) {
            const int L = 1 << lg, per_pass = 32 >> lg;
            int cost = 0;
            for (int b = 0; b < nrows; b += per_pass) {
                int m = 0;
                for (int r = b; r < min(b + per_pass, nrows); ++r) m = max(m, len[r]);
                cost += (m + L - 1) / L + 2 + lg;  // iterations + per-pass overhead (bounds, shuffles, store)
            }
            if (cost < best_cost) {
                best_cost = cost;
                best = lg;
            }
        }
        return best;
    }
};

template <bool FILL>
__global__ void chunk_walk_kernel(const int *__restrict__ rowptr, int rows, int cap, int block_rows, int nblocks,
                                  int *__restrict__ counts,            // !FILL: out, FILL: exclusive offsets in
                                  int2 *__restrict__ desc, unsigned char *__restrict__ lanes_lg,
                                  int4 *__restrict__ long_rows, int *__restrict__ n_long) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int r_begin = b * block_rows, r_end = min(r_begin + block_rows, rows);
    int out = FILL ? counts[b] : 0;
    int n = 0;
    int len[MAX_CHUNK_ROWS];
    int open_rows = 0, open_nnz = 0, open_row0 = r_begin;
    int prev = __ldg(rowptr + r_begin);
    auto close_open = [&]() {
        if (FILL) {
            desc[out] = make_int2(open_row0, __ldg(rowptr + open_row0));
            lanes_lg[out] = (unsigned char)ChunkWalk::lanes_for(len, open_rows);
        }
        ++out;
        ++n;
        open_rows = 0;
        open_nnz = 0;
    };
    for (int r = r_begin; r < r_end; ++r) {
        const int next = __ldg(rowptr + r + 1);
        const int l = next - prev;
        if (l > cap) {
            if (open_rows) close_open();
            const int pieces = (l + cap - 1) / cap;
            if (FILL) {
                for (int p = 0; p < pieces; ++p) {
                    desc[out + p] = make_int2(r, prev + p * cap);
                    lanes_lg[out + p] = 5;
                }
                long_rows[atomicAdd(n_long, 1)] = make_int4(r, out, pieces, 0);
            }
            out += pieces;
            n += pieces;
            open_row0 = r + 1;
        } else {
            if (open_nnz + l > cap) close_open();
            if (open_rows == 0) open_row0 = r;
            len[open_rows++] = l;
            open_nnz += l;
        }
        prev = next;
    }
    if (open_rows) close_open();
    if (!FILL) counts[b] = n;
}

__global__ void count_long_rows_kernel(const int *__restrict__ rowptr, int rows, int cap, int *__restrict__ n_long,
                                       int *__restrict__ max_len) {
    int c = 0, m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x) {
        const int l = __ldg(rowptr + i + 1) - __ldg(rowptr + i);
        c += l > cap;
        m = max(m, l);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (c) atomicAdd(n_long, c);
        if (m) atomicMax(max_len, m);
    }
}

__global__ void write_sentinel_kernel(int2 *desc, int nchunks, int rows, const int *rowptr) {
    desc[nchunks] = make_int2(rows, rowptr[rows]);
}

// ------------------------------------------------------------------------------------------------------
// Executor
// ------------------------------------------------------------------------------------------------------
struct SpmvArgs {
    const int *rowptr;
    const int *colids;
    const double *values;
    const double *x;
    double *y;
    const int2 *desc;
    const unsigned char *lanes_lg;
    double *carry;
    const int *row_map;  // optional: local row r is y[row_map[r]] (row-compressed off-diagonal blocks)
    int rows;
    long long nnz;
    int nchunks;       // one past the last chunk this launch processes
    int chunk_begin;   // first chunk this launch processes (row-block launches of the pipelined host path)
    int force_lg;  // -1: use the inspector's choice
    const unsigned char *part_flags;  // partitioned x: 1 where a chunk references a column another GPU owns
};

// x split over the GPUs of one NVSwitch box: part q holds x[cut[q] .. cut[q+1]) and base[q] is a pointer this GPU
// can dereference (its own memory for q == self, CUDA-IPC-mapped peer memory otherwise).  Loads of remote entries
// go straight over NVLink from inside the SpMV kernel: there is no separate exchange step.
constexpr int MAX_PARTS = 8;
struct XParts {
    const double *base[MAX_PARTS];
    int cut[MAX_PARTS + 1];
    int world;
    int lo, hi;  // this GPU's own column range and slice
    const double *self_base;
    // readiness flags written by the owners (flags[q] >= epoch: rank q's slice for this product is in place)
    const unsigned long long *flags;
    unsigned long long epoch;
    unsigned long long *signal[MAX_PARTS];  // every rank's flag array (peer-mapped), or null: the caller signals
    int self;
    // in-kernel halo pull (localized handles): the first gather_ctas CTAs copy the remote x entries this rank needs
    // (needed[k] = global column of halo[k]) into local memory, then count themselves into halo_done; column ids of
    // remote entries were rewritten to -1 - k, so chunks never touch NVLink themselves
    const int *needed;
    int n_needed;
    double *halo;
    unsigned long long *halo_done;
    unsigned long long halo_target;
    int gather_ctas;
    unsigned owner_mask;  // ranks whose flags this rank waits for (the owners of its remote columns; all if unknown)
    // dynamic tail (localized plans): the last pool_n chunks of the sweep are not dealt out statically but claimed one at
    // a time from pool_counter by whichever warp runs out of work first — the gather CTAs and the warps that had to wait
    // for a peer start their matrix stream late, and with a purely static deal the whole launch ends that much later.
    // The counter only grows: this launch's claims start at pool_base (every warp's last, failing claim is counted too).
    unsigned long long *pool_counter;
    unsigned long long pool_base;
    int pool_n;
    // how long a kernel waits for a peer's flag before it gives up (nanoseconds).  The flag is published by the peer's own
    // product kernel, so the wait covers everything that can delay a peer's launch (plan building on its first product, host
    // I/O between iterations, a debugger): the default is 10 minutes, G4S_PEER_TIMEOUT_S changes it.  Expiry traps — a peer
    // that never launches would otherwise hang every rank silently.
    unsigned long long wait_ns;
};
__device__ __forceinline__ double load_x_part(const XParts &xp, int c) {
    if (c >= xp.lo && c < xp.hi) return __ldg(xp.self_base + (c - xp.lo));  // own slice: the common case
    const double *b = xp.base[0];  // static indices only: the struct stays in the kernel-parameter bank
    int cut = xp.cut[0];
#pragma unroll
    for (int i = 1; i < MAX_PARTS; ++i)
        if (i < xp.world && c >= xp.cut[i]) {
            b = xp.base[i];
            cut = xp.cut[i];
        }
    return *reinterpret_cast<const volatile double *>(b + (c - cut));  // ordered after the flag acquire
}

// same lookup without the early-out, returning the address: lets the compiler issue many of them back to back
__device__ __forceinline__ const double *x_part_ptr(const XParts &xp, int c) {
    const double *b = xp.base[0];
    int cut = xp.cut[0];
#pragma unroll
    for (int i = 1; i < MAX_PARTS; ++i) {
        const bool ge = i < xp.world && c >= xp.cut[i];
        b = ge ? xp.base[i] : b;
        cut = ge ? xp.cut[i] : cut;
    }
    return b + (c - cut);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// warp-collective: lane q waits until rank q has published this product's epoch.  xp.wait_ns (default 10 minutes) without
// progress means a peer died: trap, so that the failure is loud instead of a silent hang.
__device__ __forceinline__ void wait_peers(const XParts &xp, int lane) {
    if (lane < xp.world && ((xp.owner_mask >> lane) & 1u)) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys_u64(xp.flags + lane) < xp.epoch) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > xp.wait_ns) __trap();
            __nanosleep(200);
        }
    }
    __syncwarp();
}

template <int CAP>
struct alignas(32) ChunkBuf {
    double vals[CAP + 8];
    int cols[CAP + 8];
};

template <int L, int CAP, bool ACCUM, bool PART>
__device__ __forceinline__ void chunk_rows(const ChunkBuf<CAP> &buf, const SpmvArgs &a, const XParts &xp, int chunk,
                                           int row0, int row1, int nnz0, int nnz1, int lane, int rp0, int rp1) {
    const int a0 = nnz0 & ~3;
    const int R = row1 - row0;
    const int nloc = R > 0 ? R : 1;  // R == 0: a non-final piece of a long row (its sum goes to carry)
    constexpr int PER_PASS = 32 / L;
    const int g = lane / L, sub = lane % L;
    const double *__restrict__ x = a.x;
    if (PART && xp.halo) {
        // Localized plan: remote entries carry column -1 - k and read halo[k], which the gather CTAs filled from the
        // owners' memory earlier in this launch.
        const double *halo = xp.halo;
        for (int base = 0; base < nloc; base += PER_PASS) {
            const int r = base + g;
            int s = 0, e = 0;
            if (r < nloc) {
                if (base) {
                    rp0 = __ldg(a.rowptr + row0 + r);
                    rp1 = __ldg(a.rowptr + row0 + r + 1);
                }
                s = max(rp0, nnz0);
                e = min(rp1, nnz1);
            }
            double acc = 0.0;
            int k = s + sub - a0;
            const int ke = e - a0;
#pragma unroll 4
            for (; k < ke; k += L) {
                // one load instruction for both kinds of entry (a pointer select, no divergent branch), cached in L1:
                // this SM cannot hold a stale copy of a halo line, because it only reads the halo after the acquire on
                // halo_done and L1 starts every launch empty.  (The first version chose between ld.global.nc and
                // ld.global.cg per element: the boundary chunks ran ~4x slower than interior ones, which at 8 GPUs —
                // 4 % of a rank's chunks — cost 12 % of the step.)
                const int c = buf.cols[k];
                const double *px = c >= 0 ? x + c : halo + (-1 - c);
                double xv;
                asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(xv) : "l"(px));
                acc = fma(buf.vals[k], xv, acc);
            }
#pragma unroll
            for (int o = L >> 1; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (sub == 0 && r < nloc) {
                if (R > 0) a.y[row0 + r] = acc;
                else a.carry[chunk] = acc;
            }
        }
        return;
    }
    // FLAT: gather first, sum afterwards.  Every staged value is multiplied by its x entry in one flat sweep over the
    // chunk — 32 consecutive nonzeros per step, 8 steps in flight per lane, all loads independent of the row structure —
    // and the row sums below read finished products out of shared memory.  Used for chunks that touch another GPU's
    // slice: NVLink loads cost microseconds and must not be chained into the row sums.  (Measured for the 256-nnz chunks of
    // power-law graphs too, where a lane owns one or two nonzeros of a row: R-MAT 24 1.315 -> 1.385 ms.  That kernel does
    // not wait for gathers in flight; it sits on the L2's request rate — 263 M distinct 32-byte sectors per product over
    // ~96 slices at one request per clock is 1.39 ms — profiles/r02_rmat_summary.md.)
    // The products are rounded before they are added (no FMA chain): within the parity bound 1e-12 sum|a||x|.
    constexpr bool FLAT = PART;
    if (FLAT) {
        double *pv = const_cast<double *>(buf.vals);
        const int kb = nnz0 - a0, kend = nnz1 - a0;
        int k = kb + lane;
        for (; k + 7 * 32 < kend; k += 8 * 32) {
            double xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = PART ? *x_part_ptr(xp, buf.cols[k + u * 32]) : __ldg(x + buf.cols[k + u * 32]);
#pragma unroll
            for (int u = 0; u < 8; ++u) pv[k + u * 32] *= xv[u];
        }
        for (; k < kend; k += 32) pv[k] *= PART ? *x_part_ptr(xp, buf.cols[k]) : __ldg(x + buf.cols[k]);
        __syncwarp();
    }
    for (int base = 0; base < nloc; base += PER_PASS) {
        const int r = base + g;
        int s = 0, e = 0;
        if (r < nloc) {
            if (base) {  // the first pass's row pointers were loaded before the wait on the copy
                rp0 = __ldg(a.rowptr + row0 + r);
                rp1 = __ldg(a.rowptr + row0 + r + 1);
            }
            s = max(rp0, nnz0);
            e = min(rp1, nnz1);
        }
        double acc = 0.0;
        int k = s + sub - a0;
        const int ke = e - a0;
#pragma unroll 4
        for (; k < ke; k += L) acc = FLAT ? acc + buf.vals[k] : fma(buf.vals[k], __ldg(x + buf.cols[k]), acc);
#pragma unroll
        for (int o = L >> 1; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0 && r < nloc) {
            if (R > 0) {
                const int row = a.row_map ? __ldg(a.row_map + row0 + r) : row0 + r;
                if (ACCUM) a.y[row] += acc;
                else a.y[row] = acc;
            } else {
                a.carry[chunk] = acc;
            }
        }
    }
}

template <int CAP, int NBUF, int WARPS, bool ACCUM, bool PART>
__global__ void __launch_bounds__(WARPS * 32) spmv_chunk_kernel(const SpmvArgs a, const XParts xp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[WARPS * NBUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ChunkBuf<CAP> *bufs = reinterpret_cast<ChunkBuf<CAP> *>(smem_raw) + warp * NBUF;
    uint64_t *bar = bars + warp * NBUF;
    uint64_t policy = 0;

    // lane 0: stage chunk c into ring slot b
    auto issue = [&](int b, int nnz0, int nnz1) {
        // bulk-copy the 16-byte-aligned cover of [nnz0, nnz1); whatever sticks out past the last aligned
        // element of the arrays (only in the matrix's final chunk) is copied with plain loads
        const int a0 = nnz0 & ~3;
        long long a1 = ((long long)nnz1 + 3) & ~3LL;
        const long long amax = a.nnz & ~3LL;
        if (a1 > amax) a1 = amax;
        if (a1 < a0) a1 = a0;
        const uint32_t ncopy = (uint32_t)(a1 - a0);
        for (long long k = a1; k < nnz1; ++k) {
            bufs[b].cols[k - a0] = a.colids[k];
            bufs[b].vals[k - a0] = a.values[k];
        }
        mbar_arrive_expect_tx(&bar[b], ncopy * 12u);
        if (ncopy) {
            bulk_g2s(bufs[b].vals, a.values + a0, ncopy * 8u, &bar[b], policy);
            bulk_g2s(bufs[b].cols, a.colids + a0, ncopy * 4u, &bar[b], policy);
        }
    };

    const int wglobal = a.chunk_begin + blockIdx.x * WARPS + warp;
    const int stride = gridDim.x * WARPS;
    // this warp's chunks are wglobal, wglobal + stride, ...: nit of them, in ascending order, so that at any moment
    // the whole grid works on one narrow band of rows and x is reused out of L1/L2.  (Starting the CTAs at 8 different
    // phases of the sweep, to de-synchronise the NVLink-bound chunks at the ends of a rank's range, was measured:
    // 8 bands cost more in x re-reads — 0.42 -> 0.52 ms on 8 M rows — than the de-synchronisation gains.)
    const int nit = wglobal < a.nchunks ? (int)(((long long)a.nchunks - wglobal + stride - 1) / stride) : 0;
    // With an in-kernel halo the whole grid sweeps the chunk list ROTATED by half its length: the chunks that need the
    // halo sit at the two ends of a rank's range, so they come up half a kernel after the gather CTAs started.  Position
    // rho of the rotated sweep is chunk (rho + mid) mod n; positions below n_static are dealt out round-robin, the rest
    // (the dynamic tail) is claimed from a counter.
    const bool dyn = PART && NBUF == 1 && xp.halo != nullptr;
    const int n_all = a.nchunks, mid = dyn ? n_all / 2 : 0;
    const int n_static = dyn ? n_all - xp.pool_n : n_all;
    auto chunk_of = [&](long long rho) -> long long {
        long long c = rho + mid;
        return c >= n_all ? c - n_all : c;
    };
    auto chunk_at = [&](int i) -> long long { return (long long)wglobal + (long long)i * stride; };
    // lane 0: the position after rho (static successor, else one claim from the pool; -1: the sweep is over)
    auto next_pos = [&](long long rho) -> long long {
        long long nx = rho + stride;
        if (rho < n_static && nx < n_static) return nx;
        if (xp.pool_n <= 0 && n_static >= n_all) return -1;
        const unsigned long long t = atomicAdd(xp.pool_counter, 1ULL) - xp.pool_base;
        return (long long)t < (long long)xp.pool_n ? (long long)n_static + (long long)t : -1;
    };
    if (PART && xp.flags && xp.signal[0] && blockIdx.x == 0 && warp == 0 && lane < xp.world) {
        // publish this rank's slice for this product (it was written before the launch): a release store of the epoch
        // into every rank's flag array, slot self — the separate signal launch is folded into the product
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(xp.signal[lane] + xp.self), "l"(xp.epoch) : "memory");
    }
    long long cur = -1;  // dyn: this warp's current position of the rotated sweep (warp-uniform)
    if (lane == 0) {
        policy = policy_evict_first();
        for (int b = 0; b < NBUF; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
        if (dyn) {
            cur = wglobal < n_static ? (long long)wglobal : next_pos((long long)n_static);
            if (cur >= 0) {
                const long long c = chunk_of(cur);
                issue(0, __ldg(a.desc + c).y, __ldg(a.desc + c + 1).y);
            }
        } else {
            for (int b = 0; b < NBUF; ++b) {
                if (b < nit) {
                    const long long c = chunk_at(b);
                    issue(b, __ldg(a.desc + c).y, __ldg(a.desc + c + 1).y);
                }
            }
        }
    }
    if (dyn) cur = __shfl_sync(0xffffffffu, cur, 0);
    __syncwarp();

    bool peers_ready = !PART || xp.flags == nullptr;
    bool halo_ready = !(PART && xp.halo);
    if (PART && xp.halo && (int)blockIdx.x < xp.gather_ctas) {
        // halo pull: wait for the owners, then copy the needed remote entries with 8 independent NVLink loads in
        // flight per lane, publish them device-wide and count this warp in
        if (!peers_ready) {
            wait_peers(xp, lane);
            peers_ready = true;
        }
        const int gstride = xp.gather_ctas * WARPS * 32;
        int k = (blockIdx.x * WARPS + warp) * 32 + lane;
        for (; k + 7 * gstride < xp.n_needed; k += 8 * gstride) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = *x_part_ptr(xp, __ldg(xp.needed + k + u * gstride));
#pragma unroll
            for (int u = 0; u < 8; ++u) xp.halo[k + u * gstride] = v[u];
        }
        for (; k < xp.n_needed; k += gstride) xp.halo[k] = *x_part_ptr(xp, __ldg(xp.needed + k));
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(xp.halo_done, 1ULL);
    }
    for (int it = 0; dyn ? cur >= 0 : it < nit; ++it) {
        const long long c = dyn ? chunk_of(cur) : chunk_at(it);
        const int b = it % NBUF;
        const uint32_t parity = (it / NBUF) & 1;
        const int2 d0 = __ldg(a.desc + c), d1 = __ldg(a.desc + c + 1);
        const int code = PART ? (int)__ldg(a.part_flags + c) : (int)__ldg(a.lanes_lg + c);
        const int lg = a.force_lg >= 0 ? a.force_lg : (code & 7);
        // loads that do not depend on the staged data go out before the wait: this chunk's row pointers
        // and the descriptor of the chunk that will reuse this ring slot
        const int nloc = d1.x > d0.x ? d1.x - d0.x : 1;
        int rp0 = 0, rp1 = 0;
        if ((lane >> lg) < nloc) {
            rp0 = __ldg(a.rowptr + d0.x + (lane >> lg));
            rp1 = __ldg(a.rowptr + d0.x + (lane >> lg) + 1);
        }
        bool has_next = it + NBUF < nit;
        long long nxt = -1;
        int2 n0 = make_int2(0, 0), n1 = n0;
        if (lane == 0) {
            if (dyn) {
                nxt = next_pos(cur);
                has_next = nxt >= 0;
            }
            if (has_next) {
                const long long next = dyn ? chunk_of(nxt) : chunk_at(it + NBUF);
                n0 = __ldg(a.desc + next);
                n1 = __ldg(a.desc + next + 1);
            }
        }
        mbar_wait(&bar[b], parity);
        const ChunkBuf<CAP> &buf = bufs[b];
        // partitioned x: only chunks that reference a remote column pay for the owner lookup; every other chunk
        // runs the plain code on the GPU's own slice (a.x = own slice rebased to global column ids)
        const bool remote = PART && (code & 0x80);
        if (PART && remote && xp.halo) {
            if (!halo_ready) {  // the gather CTAs have published the whole halo for this product
                if (lane == 0) {
                    unsigned long long t0, t1, seen;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                    do {
                        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(xp.halo_done) : "memory");
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > xp.wait_ns) __trap();
                    } while (seen < xp.halo_target);
                }
                __syncwarp();
                halo_ready = true;
            }
        } else if (PART && remote && !peers_ready) {  // only chunks that touch another GPU's slice ever wait
            wait_peers(xp, lane);
            peers_ready = true;
        }
#define G4S_CHUNK_CASE(LANES)                                                                                         \
    if (remote) chunk_rows<LANES, CAP, ACCUM, true>(buf, a, xp, (int)c, d0.x, d1.x, d0.y, d1.y, lane, rp0, rp1);       \
    else chunk_rows<LANES, CAP, ACCUM, false>(buf, a, xp, (int)c, d0.x, d1.x, d0.y, d1.y, lane, rp0, rp1);             \
    break;
        switch (lg) {
            case 0: G4S_CHUNK_CASE(1)
            case 1: G4S_CHUNK_CASE(2)
            case 2: G4S_CHUNK_CASE(4)
            case 3: G4S_CHUNK_CASE(8)
            case 4: G4S_CHUNK_CASE(16)
            default: G4S_CHUNK_CASE(32)
        }
#undef G4S_CHUNK_CASE
        __syncwarp();  // every lane is done reading slot b
        if (lane == 0 && has_next) issue(b, n0.y, n1.y);
        if (dyn) cur = __shfl_sync(0xffffffffu, nxt, 0);
    }
    // one warp per GPU always observes every rank's flag before the kernel ends, even when no chunk was remote:
    // a rank can then never run two products ahead of a peer that still reads its double-buffered slice
    if (PART && !peers_ready && blockIdx.x == 0 && warp == 0 && xp.flags) wait_peers(xp, lane);
}

struct FlagPtrs {
    unsigned long long *ptr[MAX_PARTS];
};
// thread q publishes `epoch` into rank q's flag array, slot `self` (a peer store over NVLink for q != self)
__global__ void peer_signal_kernel(FlagPtrs f, int world, int self, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();  // everything this stream wrote before (the x slice) is visible first
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.ptr[q] + self), "l"(epoch) : "memory");
    }
}
int peer_signal(unsigned long long *const *flag_arrays, int world, int self, unsigned long long epoch, cudaStream_t stream) {
    if (world < 1 || world > MAX_PARTS || self < 0 || self >= world) return fail(G4S_ERR_INVALID, "peer signal: 1..8 ranks");
    FlagPtrs f;
    for (int q = 0; q < MAX_PARTS; ++q) f.ptr[q] = q < world ? flag_arrays[q] : nullptr;
    peer_signal_kernel<<<1, 32, 0, stream>>>(f, world, self, epoch);
    G4S_CHECK_LAUNCH("peer_signal_kernel");
    return G4S_OK;
}

// Localization of a rank's row block for the in-kernel halo pull: columns outside [lo, hi) are numbered in ascending
// order (needed[k]) and their ids rewritten to -1 - k.
__global__ void localize_flag_kernel(const int *__restrict__ colids, long long nnz, int lo, int hi, int *__restrict__ flags) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const int c = __ldg(colids + k);
        if (c < lo || c >= hi) flags[c] = 1;
    }
}
__global__ void localize_list_kernel(const int *__restrict__ flags, const int *__restrict__ pos, int cols,
                                     int *__restrict__ needed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cols && flags[c]) needed[pos[c]] = c;
}
// writes the plan's PRIVATE localized copy; the handle's own column ids are never modified
__global__ void localize_rewrite_kernel(const int *__restrict__ colids, int *__restrict__ out, long long nnz, int lo, int hi,
                                        const int *__restrict__ pos) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const int c = __ldg(colids + k);
        out[k] = (c < lo || c >= hi) ? -1 - __ldg(pos + c) : c;
    }
}

// one warp per chunk: does it reference a column outside [lo, hi)?
__global__ void chunk_remote_flags_kernel(const int2 *__restrict__ desc, const int *__restrict__ colids,
                                          const unsigned char *__restrict__ lanes_lg, int nchunks, int lo, int hi,
                                          unsigned char *__restrict__ flags) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    const int s = desc[c].y, e = desc[c + 1].y;
    int any = 0;
    for (int k = s + lane; k < e; k += 32) {
        const int col = __ldg(colids + k);
        any |= (col < lo || col >= hi);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[c] = lanes_lg[c] | (any ? 0x80 : 0);  // one byte per chunk: lanes (low bits) + remote (bit 7)
}

// y[row] += carries of the row's non-final pieces, in piece order
__global__ void spmv_long_fixup_kernel(const int4 *__restrict__ long_rows, int n_long,
                                       const double *__restrict__ carry, const int *__restrict__ row_map,
                                       double *__restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    const int4 lr = long_rows[i];
    double s = 0.0;
    for (int p = 0; p < lr.z - 1; ++p) s += carry[lr.y + p];
    y[row_map ? row_map[lr.x] : lr.x] += s;
}

// Baseline kept for comparison: one warp per row straight from global memory, no staging.
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
// CTAs a launch may use, in percent of the full persistent grid (0: all).  The pipelined host-pointer product throttles its
// row-block launches: a block's product only has to finish before the next block of x has arrived, and a kernel that
// saturates HBM slows the host link's DMA down while it runs (spmv_host_pipelined).
static thread_local int t_grid_percent = 0;
template <int CAP, int NBUF, int WARPS>
static int launch_chunk_kernel(const SpmvArgs &args, bool accum, int ctas_per_sm, cudaStream_t stream,
                               const XParts *parts = nullptr) {
    const size_t smem = sizeof(ChunkBuf<CAP>) * NBUF * WARPS;
    auto k0 = spmv_chunk_kernel<CAP, NBUF, WARPS, false, false>;
    auto k1 = spmv_chunk_kernel<CAP, NBUF, WARPS, true, false>;
    auto kp = spmv_chunk_kernel<CAP, NBUF, WARPS, false, true>;
    static PerDeviceOnce configured;  // keyed by the CTAs per SM: shapes that share an instantiation differ in the carve-out
    if (configured.needs(ctas_per_sm)) {
        G4S_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        G4S_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        G4S_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // carve out exactly what the resident CTAs need: the rest of the 228 KB stays L1 for the x gathers
        const int pct = std::min(100, (int)((ctas_per_sm * (smem + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024)));
        G4S_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        G4S_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        G4S_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        configured.done(ctas_per_sm);
    }
    long long want = ((long long)(args.nchunks - args.chunk_begin) + WARPS - 1) / WARPS;
    int grid = (int)std::min<long long>(want, (long long)sm_count() * ctas_per_sm);
    if (t_grid_percent > 0) grid = (int)(((long long)grid * t_grid_percent + 99) / 100);
    if (grid < 1) grid = 1;
    XParts none;
    none.world = 0;
    none.lo = none.hi = 0;
    none.self_base = nullptr;
    none.flags = nullptr;
    none.epoch = 0;
    none.self = 0;
    for (int q = 0; q < MAX_PARTS; ++q) none.signal[q] = nullptr;
    none.needed = nullptr;
    none.n_needed = 0;
    none.halo = nullptr;
    none.halo_done = nullptr;
    none.halo_target = 0;
    none.gather_ctas = 0;
    none.owner_mask = 0xffu;
    none.pool_counter = nullptr;
    none.pool_base = 0;
    none.pool_n = 0;
    none.wait_ns = 0;
    if (parts) kp<<<grid, WARPS * 32, smem, stream>>>(args, *parts);
    else if (accum) k1<<<grid, WARPS * 32, smem, stream>>>(args, none);
    else k0<<<grid, WARPS * 32, smem, stream>>>(args, none);
    G4S_CHECK_LAUNCH("spmv_chunk_kernel");
    return G4S_OK;
}

constexpr int SPMV_CAP = 1024;
constexpr int G4S_SPMV_AUTO_REGULAR_SHORT = 15;  // short regular rows: 512-nnz chunks of up to 128 rows, 36 warps per SM
constexpr int G4S_SPMV_AUTO_SHORT = 12;  // short-row matrices: 256-nnz chunks, 32 warps per SM, 128 KB left as L1
void spmv_free_plan(g4s_csr *h);
void spmv_free_host_pipe(g4s_csr *h);

// Kernel shapes (tuning variants).  cap = nonzeros per chunk = size of a warp's staging buffer (12 bytes each);
// what shared memory the buffers leave becomes L1, which is where scattered x gathers (power-law graphs) live.
struct SpmvShape {
    int cap, nbuf, warps, ctas;
};
static SpmvShape shape_of(int variant) {
    switch (variant) {
        case 1: return {1024, 1, 16, 1};
        case 2: return {1024, 2, 8, 1};
        case 3: return {1024, 2, 4, 2};
        case 4: return {1024, 1, 8, 2};
        case 5: return {1024, 1, 6, 3};
        case 7: return {512, 1, 16, 1};
        case 8: return {512, 1, 24, 1};
        case 10: return {512, 1, 16, 2};
        case 11: return {512, 1, 32, 1};
        case 12: return {256, 1, 32, 1};
        case 13: return {256, 1, 32, 2};
        case 14: return {256, 1, 24, 2};
        case 15: return {512, 1, 18, 2};
        default: return {1024, 1, 9, 2};
    }
}
// Automatic choice from the row-length statistics (profiles/r01_other_kernels_summary.md):
//  * rows long enough that 32 of them fill a 1024-nnz chunk (27-point stencils): 1024-nnz chunks, 18 warps per SM;
//  * short REGULAR rows (2-D 5-point: 32 rows are only 160 nonzeros, and the per-chunk latency dominates): 512-nnz
//    chunks of up to 128 rows (several 32-row passes per chunk) and 36 warps per SM;
//  * short SKEWED rows (power-law graphs): 256-nnz chunks, 32 warps per SM and 128 KB left as L1 for the x gathers.
struct SpmvConfig {
    int variant, rows_per_chunk;
};
static SpmvConfig choose_config(const g4s_csr *h, int forced_variant, int max_row_len) {
    const double avg = h->rows ? (double)h->nnz / h->rows : 0.0;
    const bool regular = max_row_len <= 4.0 * avg + 8.0;
    int variant = forced_variant;
    if (!variant) variant = avg * CHUNK_ROWS > 512.0 ? 6 : (regular ? G4S_SPMV_AUTO_REGULAR_SHORT : G4S_SPMV_AUTO_SHORT);
    int rpc = CHUNK_ROWS;
    if (regular) {
        const int cap = shape_of(variant).cap;
        while (rpc < MAX_CHUNK_ROWS && avg * (2 * rpc) <= (double)cap) rpc *= 2;
    }
    return {variant, rpc};
}

int spmv_build_plan(g4s_csr *h, cudaStream_t stream) {
    SpmvPlan &p = h->plan;
    if (p.desc && p.built_for_variant == p.variant) return G4S_OK;
    if (p.desc) {  // the tuning variant changed: cut the rows again
        spmv_free_host_pipe(h);
        spmv_free_plan(h);
    }
    int *counts = nullptr, *stats = nullptr;
    G4S_CUDA(cudaMalloc(&stats, sizeof(int) * 4));
    G4S_CUDA(cudaMemsetAsync(stats, 0, sizeof(int) * 4, stream));
    count_long_rows_kernel<<<sm_count() * 8, 256, 0, stream>>>(h->rowptr, h->rows, 0x7fffffff, stats, stats + 1);
    G4S_CHECK_LAUNCH("count_long_rows_kernel");
    int hmax[2] = {0, 0};
    G4S_CUDA(cudaMemcpyAsync(hmax, stats, sizeof(hmax), cudaMemcpyDeviceToHost, stream));
    G4S_CUDA(cudaStreamSynchronize(stream));
    const SpmvConfig cfg = choose_config(h, p.variant, hmax[1]);
    p.built_for_variant = p.variant;
    p.shape_variant = cfg.variant;
    p.rows_per_chunk = cfg.rows_per_chunk;
    const int want_cap = shape_of(cfg.variant).cap;
    const int block_rows = cfg.rows_per_chunk;
    p.cap = want_cap;
    const int nblocks = (h->rows + block_rows - 1) / block_rows;
    G4S_CUDA(cudaMalloc(&counts, sizeof(int) * ((size_t)nblocks + 1)));
    G4S_CUDA(cudaMemsetAsync(stats, 0, sizeof(int) * 4, stream));
    const int threads = 128;
    count_long_rows_kernel<<<sm_count() * 8, 256, 0, stream>>>(h->rowptr, h->rows, p.cap, stats, stats + 1);
    G4S_CHECK_LAUNCH("count_long_rows_kernel");
    chunk_walk_kernel<false><<<(nblocks + threads - 1) / threads, threads, 0, stream>>>(
        h->rowptr, h->rows, p.cap, block_rows, nblocks, counts, nullptr, nullptr, nullptr, nullptr);
    G4S_CHECK_LAUNCH("chunk_walk_kernel<count>");
    long long total = 0;
    int rc = exclusive_scan_i32(counts, counts, nblocks, 0, &total, stream);
    if (rc) return rc;
    int hstats[4];
    G4S_CUDA(cudaMemcpy(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost));
    if (total > 2147483000LL) return fail(G4S_ERR_INVALID, "spmv plan: too many chunks");
    p.nchunks = (int)total;
    p.n_long = hstats[0];
    p.max_row_len = hstats[1];
    G4S_CUDA(cudaMalloc(&p.desc, sizeof(int2) * ((size_t)p.nchunks + 1)));
    G4S_CUDA(cudaMalloc(&p.lanes_lg, (size_t)p.nchunks + 1));
    G4S_CUDA(cudaMalloc(&p.carry, sizeof(double) * (size_t)std::max(p.nchunks, 1)));
    G4S_CUDA(cudaMalloc(&p.long_rows, sizeof(int4) * (size_t)std::max(p.n_long, 1)));
    G4S_CUDA(cudaMemsetAsync(p.carry, 0, sizeof(double) * (size_t)std::max(p.nchunks, 1), stream));
    G4S_CUDA(cudaMemsetAsync(stats + 2, 0, sizeof(int), stream));
    chunk_walk_kernel<true><<<(nblocks + threads - 1) / threads, threads, 0, stream>>>(
        h->rowptr, h->rows, p.cap, block_rows, nblocks, counts, p.desc, p.lanes_lg, p.long_rows, stats + 2);
    G4S_CHECK_LAUNCH("chunk_walk_kernel<fill>");
    write_sentinel_kernel<<<1, 1, 0, stream>>>(p.desc, p.nchunks, h->rows, h->rowptr);
    G4S_CHECK_LAUNCH("write_sentinel_kernel");
    G4S_CUDA(cudaStreamSynchronize(stream));
    cudaFree(counts);
    cudaFree(stats);
    return G4S_OK;
}

void spmv_free_plan(g4s_csr *h) {
    SpmvPlan &p = h->plan;
    if (p.desc) cudaFree(p.desc);
    if (p.lanes_lg) cudaFree(p.lanes_lg);
    if (p.carry) cudaFree(p.carry);
    if (p.long_rows) cudaFree(p.long_rows);
    if (p.part_flags) cudaFree(p.part_flags);
    if (p.part_needed) cudaFree(p.part_needed);
    if (p.part_halo) cudaFree(p.part_halo);
    if (p.part_halo_done) cudaFree(p.part_halo_done);
    if (p.part_colids) cudaFree(p.part_colids);
    const int lanes = p.lanes_per_row, variant = p.variant;
    p = SpmvPlan();
    p.lanes_per_row = lanes;
    p.variant = variant;
}

int spmv_run(g4s_csr *h, const double *x, double *y, const int *row_map, bool accum, cudaStream_t stream,
             const XParts *parts, int chunk_begin, int chunk_end) {
    if (h->rows == 0) return G4S_OK;
    int rc = spmv_build_plan(h, stream);
    if (rc) return rc;
    const SpmvPlan &p = h->plan;
    if (p.variant == 9) {  // comparison baseline (the plan is unused)
        if (accum || row_map || parts) return fail(G4S_ERR_INVALID, "warp-per-row baseline supports plain y = A x only");
        spmv_warp_row_kernel<<<sm_count() * 8, 256, 0, stream>>>(h->rowptr, h->colids, h->values, x, y, h->rows);
        G4S_CHECK_LAUNCH("spmv_warp_row_kernel");
        return G4S_OK;
    }
    SpmvArgs a;
    a.rowptr = h->rowptr;
    // a partitioned product on a localized plan reads the plan's private copy of the column ids (remote columns
    // numbered -1 - k); the handle's own array keeps the global ids for every other entry point
    a.colids = (parts && parts->halo && h->plan.part_colids) ? h->plan.part_colids : h->colids;
    a.values = h->values;
    a.x = x;
    a.y = y;
    a.desc = p.desc;
    a.lanes_lg = p.lanes_lg;
    a.carry = p.carry;
    a.row_map = row_map;
    a.rows = h->rows;
    a.nnz = h->nnz;
    a.nchunks = chunk_end >= 0 ? chunk_end : p.nchunks;
    a.chunk_begin = chunk_begin;
    a.force_lg = -1;
    a.part_flags = nullptr;
    a.part_flags = parts ? p.part_flags : nullptr;
    if (p.lanes_per_row > 0) {
        int lg = 0;
        while ((1 << lg) < p.lanes_per_row && lg < 5) ++lg;
        a.force_lg = lg;
    }
    // kernel shapes: chunk size x ring depth x warps per CTA x CTAs per SM (profiles/r01_spmv_summary.md,
    // r01_other_kernels_summary.md).  27-nnz stencil rows: 1024-nnz chunks, 18 single-buffered warps per SM as 2 CTAs
    // of 9 (deeper per-warp rings with fewer warps lose: resident warps hide the latency).  Short or power-law rows
    // (mean row length <= 16): 256-nnz chunks, 32 warps per SM and the rest of shared memory left to L1.
    const int variant = parts ? 6 : p.shape_variant;
    switch (variant) {
        case 1: rc = launch_chunk_kernel<1024, 1, 16>(a, accum, 1, stream); break;
        case 2: rc = launch_chunk_kernel<1024, 2, 8>(a, accum, 1, stream); break;
        case 3: rc = launch_chunk_kernel<1024, 2, 4>(a, accum, 2, stream); break;
        case 4: rc = launch_chunk_kernel<1024, 1, 8>(a, accum, 2, stream); break;
        case 5: rc = launch_chunk_kernel<1024, 1, 6>(a, accum, 3, stream); break;
        case 7: rc = launch_chunk_kernel<512, 1, 16>(a, accum, 1, stream); break;
        case 8: rc = launch_chunk_kernel<512, 1, 24>(a, accum, 1, stream); break;
        case 10: rc = launch_chunk_kernel<512, 1, 16>(a, accum, 2, stream); break;
        case 11: rc = launch_chunk_kernel<512, 1, 32>(a, accum, 1, stream); break;
        case 12: rc = launch_chunk_kernel<256, 1, 32>(a, accum, 1, stream); break;
        case 13: rc = launch_chunk_kernel<256, 1, 32>(a, accum, 2, stream); break;
        case 14: rc = launch_chunk_kernel<256, 1, 24>(a, accum, 2, stream); break;
        case 15: rc = launch_chunk_kernel<512, 1, 18>(a, accum, 2, stream); break;
        default: rc = launch_chunk_kernel<1024, 1, 9>(a, accum, 2, stream, parts); break;
    }
    if (rc) return rc;
    if (p.n_long > 0 && chunk_end < 0) {  // row-block launches leave the fix-up to the caller
        spmv_long_fixup_kernel<<<(p.n_long + 127) / 128, 128, 0, stream>>>(p.long_rows, p.n_long, p.carry, row_map, y);
        G4S_CHECK_LAUNCH("spmv_long_fixup_kernel");
    }
    return G4S_OK;
}

int spmv_run_partitioned(g4s_csr *h, int world, int self, const double *const *x_parts, const int *cuts, double *y,
                         const unsigned long long *flags, unsigned long long epoch,
                         unsigned long long *const *signal_arrays, cudaStream_t stream) {
    if (world < 1 || world > MAX_PARTS || self < 0 || self >= world)
        return fail(G4S_ERR_INVALID, "partitioned SpMV supports 1..8 parts (one NVSwitch box)");
    XParts xp;
    for (int q = 0; q < MAX_PARTS; ++q) xp.base[q] = q < world ? x_parts[q] : nullptr;
    for (int q = 0; q <= MAX_PARTS; ++q) xp.cut[q] = cuts[q < world ? q : world];
    xp.world = world;
    xp.lo = cuts[self];
    xp.hi = cuts[self + 1];
    xp.self_base = x_parts[self];
    xp.flags = flags;
    xp.epoch = epoch;
    xp.self = self;
    for (int q = 0; q < MAX_PARTS; ++q) xp.signal[q] = (signal_arrays && flags && q < world) ? signal_arrays[q] : nullptr;
    int rc = spmv_build_plan(h, stream);
    if (rc) return rc;
    SpmvPlan &p = h->plan;
    if (p.part_colids && (p.part_lo != xp.lo || p.part_hi != xp.hi)) {  // another owned range: localize again
        cudaFree(p.part_colids);
        cudaFree(p.part_needed);
        cudaFree(p.part_halo);
        cudaFree(p.part_halo_done);
        p.part_colids = nullptr;
        p.part_needed = nullptr;
        p.part_halo = nullptr;
        p.part_halo_done = nullptr;
    }
    if (!p.part_colids && world > 1) {
        // first partitioned product for this owned range: number the remote columns and write a private copy of the
        // column ids in which they read -1 - k (once; the handle's own arrays are left untouched, so the handle
        // stays valid for every other entry point and borrowed arrays are never modified)
        const int cols = h->cols, grid = sm_count() * 8;
        int *flags = nullptr, *pos = nullptr;
        G4S_CUDA(cudaMalloc(&flags, sizeof(int) * ((size_t)cols + 1)));
        G4S_CUDA(cudaMalloc(&pos, sizeof(int) * ((size_t)cols + 1)));
        G4S_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)cols + 1), stream));
        if (h->nnz) {
            localize_flag_kernel<<<grid, 256, 0, stream>>>(h->colids, h->nnz, xp.lo, xp.hi, flags);
            G4S_CHECK_LAUNCH("localize_flag_kernel");
        }
        long long nn = 0;
        if ((rc = exclusive_scan_i32(flags, pos, cols, 0, &nn, stream))) return rc;
        p.part_n_needed = (int)nn;
        G4S_CUDA(cudaMalloc(&p.part_needed, sizeof(int) * (size_t)std::max<long long>(nn, 1)));
        G4S_CUDA(cudaMalloc(&p.part_halo, sizeof(double) * (size_t)std::max<long long>(nn, 1)));
        G4S_CUDA(cudaMalloc(&p.part_halo_done, 2 * sizeof(unsigned long long)));  // [0] gather warps done, [1] tail-pool claims
        G4S_CUDA(cudaMalloc(&p.part_colids, sizeof(int) * (size_t)std::max<long long>(h->nnz, 1)));
        G4S_CUDA(cudaMemsetAsync(p.part_halo_done, 0, 2 * sizeof(unsigned long long), stream));
        if (cols) {
            localize_list_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(flags, pos, cols, p.part_needed);
            G4S_CHECK_LAUNCH("localize_list_kernel");
        }
        if (h->nnz) {
            localize_rewrite_kernel<<<grid, 256, 0, stream>>>(h->colids, p.part_colids, h->nnz, xp.lo, xp.hi, pos);
            G4S_CHECK_LAUNCH("localize_rewrite_kernel");
        }
        G4S_CUDA(cudaStreamSynchronize(stream));
        cudaFree(flags);
        cudaFree(pos);
        p.part_products = 0;
        // which ranks own those columns: a product only has to wait for them (for a stencil: the two neighbours),
        // not for all ranks of the box
        p.part_owner_mask = 0;
        if (nn > 0) {
            std::vector<int> hn((size_t)nn);
            G4S_CUDA(cudaMemcpy(hn.data(), p.part_needed, sizeof(int) * (size_t)nn, cudaMemcpyDeviceToHost));
            int q = 0;
            for (long long k = 0; k < nn; ++k) {  // needed is ascending: walk the cuts once
                while (q + 1 < world && hn[k] >= cuts[q + 1]) ++q;
                p.part_owner_mask |= 1u << q;
            }
        }
    }
    if (!p.part_flags || p.part_lo != xp.lo || p.part_hi != xp.hi) {  // once per (matrix, owned range)
        if (!p.part_flags) G4S_CUDA(cudaMalloc(&p.part_flags, (size_t)p.nchunks + 1));
        if (p.nchunks) {
            chunk_remote_flags_kernel<<<(int)(((long long)p.nchunks * 32 + 255) / 256), 256, 0, stream>>>(
                p.desc, h->colids, p.lanes_lg, p.nchunks, xp.lo, xp.hi, p.part_flags);
            G4S_CHECK_LAUNCH("chunk_remote_flags_kernel");
        }
        p.part_lo = xp.lo;
        p.part_hi = xp.hi;
    }
    xp.needed = nullptr;
    xp.n_needed = 0;
    xp.halo = nullptr;
    xp.halo_done = nullptr;
    xp.halo_target = 0;
    xp.gather_ctas = 0;
    xp.owner_mask = 0xffu;
    xp.pool_counter = nullptr;
    xp.pool_base = 0;
    xp.pool_n = 0;
    static const unsigned long long wait_ns = [] {
        const char *e = getenv("G4S_PEER_TIMEOUT_S");
        const double sec = e ? atof(e) : 600.0;
        return (unsigned long long)((sec > 0 ? sec : 600.0) * 1e9);
    }();
    xp.wait_ns = wait_ns;
    if (p.part_colids) {
        const long long want = ((long long)p.nchunks + 8) / 9;  // grid of the 1 x 9 x 2 shape (launch_chunk_kernel)
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 2));
        xp.needed = p.part_needed;
        xp.n_needed = p.part_n_needed;
        xp.halo = p.part_halo;
        xp.halo_done = p.part_halo_done;
        xp.gather_ctas = std::min(grid, 16);
        xp.owner_mask = p.part_owner_mask;
        p.part_products += 1;  // the counter only grows: product n is complete at n * (gather warps)
        xp.halo_target = p.part_products * (unsigned long long)xp.gather_ctas * 9ULL;
        // dynamic tail: 1/16 of the chunks (G4S_SPMV_POOL_SHIFT: another power-of-two share; 31 or more: none).  Every
        // warp of the grid makes exactly one failing claim, so a launch advances the counter by pool_n + warps.
        static const int shift = [] {
            const char *e = getenv("G4S_SPMV_POOL_SHIFT");
            return e ? atoi(e) : 4;
        }();
        xp.pool_n = shift >= 31 ? 0 : (p.nchunks >> shift);
        xp.pool_counter = p.part_halo_done + 1;
        xp.pool_base = (p.part_products - 1) * ((unsigned long long)xp.pool_n + (unsigned long long)grid * 9ULL);
    }
    // the plain path indexes x by global column id: rebase the own slice (never dereferenced outside [lo, hi))
    return spmv_run(h, x_parts[self] - xp.lo, y, nullptr, false, stream, &xp, 0, -1);
}

// ------------------------------------------------------------------------------------------------------
// Host-pointer product, pipelined: y = A x with x and y in (pinned) host memory.
// The chunk list is cut into NB row blocks of equal chunk count; block b needs x[0 .. colmax_b] to have arrived.
// For banded matrices colmax grows with b, so the upload of x (copy stream), the products of the blocks whose
// x range is complete (compute stream) and the download of finished y blocks (second copy stream) all overlap,
// and PCIe runs in both directions at once.  A matrix whose first block already references the end of x
// degenerates to upload-all / compute / download, which is what a non-pipelined call does.
// ------------------------------------------------------------------------------------------------------
__global__ void block_colmax_kernel(const int2 *__restrict__ desc, const int *__restrict__ colids, int nchunks,
                                    int chunks_per_block, int *__restrict__ colmax) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= nchunks) return;
    int m = -1;
    for (int k = desc[c].y + lane; k < desc[c + 1].y; k += 32) m = max(m, __ldg(colids + k));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m >= 0) atomicMax(&colmax[c / chunks_per_block], m);
}

struct HostPipe {
    // Row-block launches of a banded matrix get a QUARTER of the persistent grid each and go round-robin over four compute
    // streams.  A block's product only has to be done before the next block of x has arrived; launched at full width it
    // saturates HBM in bursts and the host link's DMA — which pays for the whole call — runs at about half speed while it does
    // (trace: 16 MB of x in 0.395 ms instead of 0.34; n = 400 27-point Laplacian 12.2-12.8 ms at full width, 11.1-11.8 ms at
    // a quarter = one CTA on every other SM; the two copies alone, chained block by block, take 10.8 ms).  If a narrow
    // product is not done when the next block of x is, the next product starts beside it on another stream: the width grows
    // by itself up to the whole grid, so a matrix (or a clock state) with slower products never waits for a throttled kernel.
    static constexpr int N_COMP = 4, NARROW_PERCENT = 25;
    bool progressive = false;  // blocks become ready while x is still arriving (banded matrices): narrow launches pay
    int nb = 0, chunks_per_block = 0;
    std::vector<int> colmax, row_end;  // per block: highest column referenced (running max), one past its last row
    cudaStream_t up = nullptr, comp[N_COMP] = {nullptr, nullptr, nullptr, nullptr}, down = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_comp;
    cudaEvent_t ev_start = nullptr;
};

int spmv_host_pipelined(g4s_csr *h, const double *x, double *y) {
    int rc = spmv_build_plan(h, 0);
    if (rc) return rc;
    const SpmvPlan &p = h->plan;
    HostPipe *&hp = h->host_pipe;
    if (!hp) {
        hp = new HostPipe();
        static const int max_blocks = [] {  // G4S_SPMV_HOST_BLOCKS: row blocks of the pipelined host-pointer product
            const char *e = getenv("G4S_SPMV_HOST_BLOCKS");
            return e && atoi(e) > 0 ? atoi(e) : 32;
        }();
        hp->nb = std::max(1, std::min(max_blocks, p.nchunks / 4096));
        hp->chunks_per_block = (p.nchunks + hp->nb - 1) / hp->nb;
        hp->nb = (p.nchunks + hp->chunks_per_block - 1) / hp->chunks_per_block;
        int *dmax = nullptr;
        G4S_CUDA(cudaMalloc(&dmax, sizeof(int) * hp->nb));
        G4S_CUDA(cudaMemset(dmax, 0xff, sizeof(int) * hp->nb));
        block_colmax_kernel<<<(int)(((long long)p.nchunks * 32 + 255) / 256), 256>>>(p.desc, h->colids, p.nchunks,
                                                                                     hp->chunks_per_block, dmax);
        G4S_CHECK_LAUNCH("block_colmax_kernel");
        hp->colmax.resize(hp->nb);
        G4S_CUDA(cudaMemcpy(hp->colmax.data(), dmax, sizeof(int) * hp->nb, cudaMemcpyDeviceToHost));
        cudaFree(dmax);
        std::vector<int2> hdesc(hp->nb + 1);
        for (int b = 0; b <= hp->nb; ++b) {
            const int c = std::min(b * hp->chunks_per_block, p.nchunks);
            G4S_CUDA(cudaMemcpy(&hdesc[b], p.desc + c, sizeof(int2), cudaMemcpyDeviceToHost));
        }
        hp->row_end.resize(hp->nb);
        for (int b = 0; b < hp->nb; ++b) {
            hp->row_end[b] = hdesc[b + 1].x;  // first row of the next block (a long row cut here is handled below)
            if (b) hp->colmax[b] = std::max(hp->colmax[b], hp->colmax[b - 1]);
        }
        hp->row_end[hp->nb - 1] = h->rows;
        hp->progressive = hp->nb >= 4 && (long long)hp->colmax[hp->nb / 2] + 1 < (long long)h->cols * 9 / 10;
        G4S_CUDA(cudaStreamCreateWithFlags(&hp->up, cudaStreamNonBlocking));
        for (auto &c : hp->comp) G4S_CUDA(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
        G4S_CUDA(cudaStreamCreateWithFlags(&hp->down, cudaStreamNonBlocking));
        hp->ev_up.resize(hp->nb);
        hp->ev_comp.resize(hp->nb);
        for (int b = 0; b < hp->nb; ++b) {
            G4S_CUDA(cudaEventCreateWithFlags(&hp->ev_up[b], cudaEventDisableTiming));
            G4S_CUDA(cudaEventCreateWithFlags(&hp->ev_comp[b], cudaEventDisableTiming));
        }
        G4S_CUDA(cudaEventCreateWithFlags(&hp->ev_start, cudaEventDisableTiming));
    }
    if (!h->x_dev) G4S_CUDA(cudaMalloc(&h->x_dev, sizeof(double) * (size_t)std::max(h->cols, 1)));
    if (!h->y_dev) G4S_CUDA(cudaMalloc(&h->y_dev, sizeof(double) * (size_t)std::max(h->rows, 1)));
    // order after whatever the caller queued on the default stream
    G4S_CUDA(cudaEventRecord(hp->ev_start, 0));
    G4S_CUDA(cudaStreamWaitEvent(hp->up, hp->ev_start, 0));
    for (auto c : hp->comp) G4S_CUDA(cudaStreamWaitEvent(c, hp->ev_start, 0));
    G4S_CUDA(cudaStreamWaitEvent(hp->down, hp->ev_start, 0));
    const bool defer_download = p.n_long > 0;  // the long-row fix-up touches y after every block has run
    long long uploaded = 0;
    int row0 = 0;
    // G4S_SPMV_HOST_TRACE=1: per-block completion times of the three streams on stderr (diagnostic; timing events per call)
    static const bool trace = [] { const char *e = getenv("G4S_SPMV_HOST_TRACE"); return e && atoi(e) != 0; }();
    std::vector<cudaEvent_t> tr;
    auto mark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tr.push_back(e);
    };
    mark(hp->up);
    static const int narrow_pct = [] {  // G4S_SPMV_HOST_GRID_PERCENT: share of the grid per row-block launch (100: full width)
        const char *e = getenv("G4S_SPMV_HOST_GRID_PERCENT");
        return e && atoi(e) > 0 ? std::min(100, atoi(e)) : HostPipe::NARROW_PERCENT;
    }();
    // the long-row fix-up runs after every block: those matrices keep one compute stream (and are not banded anyway)
    const bool narrow = hp->progressive && !defer_download && narrow_pct < 100;
    for (int b = 0; b < hp->nb; ++b) {
        const long long need = std::min<long long>((long long)hp->colmax[b] + 1, h->cols);
        if (need > uploaded) {
            G4S_CUDA(cudaMemcpyAsync(h->x_dev + uploaded, x + uploaded, sizeof(double) * (size_t)(need - uploaded),
                                     cudaMemcpyHostToDevice, hp->up));
            uploaded = need;
        }
        G4S_CUDA(cudaEventRecord(hp->ev_up[b], hp->up));
        mark(hp->up);
        cudaStream_t comp = hp->comp[narrow ? b % HostPipe::N_COMP : 0];
        G4S_CUDA(cudaStreamWaitEvent(comp, hp->ev_up[b], 0));
        const int c0 = b * hp->chunks_per_block, c1 = std::min(c0 + hp->chunks_per_block, p.nchunks);
        t_grid_percent = (narrow && b + 1 < hp->nb) ? narrow_pct : 0;  // the last block is on the critical path: full width
        rc = spmv_run(h, h->x_dev, h->y_dev, nullptr, false, comp, nullptr, c0, c1);
        t_grid_percent = 0;
        if (rc) return rc;
        G4S_CUDA(cudaEventRecord(hp->ev_comp[b], comp));
        mark(comp);
        if (!defer_download && hp->row_end[b] > row0) {
            G4S_CUDA(cudaStreamWaitEvent(hp->down, hp->ev_comp[b], 0));
            G4S_CUDA(cudaMemcpyAsync(y + row0, h->y_dev + row0, sizeof(double) * (size_t)(hp->row_end[b] - row0),
                                     cudaMemcpyDeviceToHost, hp->down));
            row0 = hp->row_end[b];
        }
        mark(hp->down);
    }
    if (defer_download) {
        spmv_long_fixup_kernel<<<(p.n_long + 127) / 128, 128, 0, hp->comp[0]>>>(p.long_rows, p.n_long, p.carry, nullptr, h->y_dev);
        G4S_CHECK_LAUNCH("spmv_long_fixup_kernel");
        G4S_CUDA(cudaMemcpyAsync(y, h->y_dev, sizeof(double) * (size_t)h->rows, cudaMemcpyDeviceToHost, hp->comp[0]));
        G4S_CUDA(cudaStreamSynchronize(hp->comp[0]));
    }
    G4S_CUDA(cudaStreamSynchronize(hp->down));
    for (auto c : hp->comp) G4S_CUDA(cudaStreamSynchronize(c));
    if (trace) {
        fprintf(stderr, "g4s_spmv_host trace (ms since the first copy was queued): block  upload  product  download\n");
        for (int b = 0; b < hp->nb; ++b) {
            float t[3] = {0, 0, 0};
            for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&t[k], tr[0], tr[1 + 3 * b + k]);
            fprintf(stderr, "  %3d  %8.3f  %8.3f  %8.3f\n", b, t[0], t[1], t[2]);
        }
        for (auto e : tr) cudaEventDestroy(e);
    }
    return G4S_OK;
}

void spmv_free_host_pipe(g4s_csr *h) {
    HostPipe *hp = h->host_pipe;
    if (!hp) return;
    for (auto e : hp->ev_up) cudaEventDestroy(e);
    for (auto e : hp->ev_comp) cudaEventDestroy(e);
    if (hp->ev_start) cudaEventDestroy(hp->ev_start);
    if (hp->up) cudaStreamDestroy(hp->up);
    for (auto c : hp->comp)
        if (c) cudaStreamDestroy(c);
    if (hp->down) cudaStreamDestroy(hp->down);
    delete hp;
    h->host_pipe = nullptr;
}

}  // namespace g4s
