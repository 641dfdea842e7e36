// BSR SpMM for sm_100a (BASELINE config 5):  C[mb*bs, ncol] = A_bsr B,  bs x bs row-major blocks, row-major
// dense B [kb*bs, ncol] and C.  The reference has no such routine (nearest relatives are
// CitcomS's 3-dof node operator, citcoms/lib/Element_calculations.c:516-571); the oracle's restatement
// (oracle/oracle_spmv.c: oracle_bsr_spmm) is the checker.
//
// Four kernels, one warp per block row:
//   dmma   (bs = 3, ncol = 64): FP64 tensor cores, mma.sync.m8n8k4.f64 (DMMA.8x8x4 in SASS; tcgen05 has no FP64
//          kind).  The product is taken transposed, C_I^T[64 x 3] = Bg^T[64 x 3nb] A_I^T[3nb x 3], so the dense
//          64 columns fill the M = 8 side of eight tiles, the 3 block rows sit in N = 8 (3/8 used) and the
//          blocks of the row are packed back to back along K (3nb, four at a time).
//   fma    (bs = 3, ncol = 64): plain DFMA, each lane owns two columns of the 64 (128-bit loads of B).
//   kpack  (bs = 3, ncol = 64): DFMA with the row's scalar columns dealt to four lane groups (a third of the broadcast
//          loads of block values); the single-GPU default.
//   generic (any bs <= 8, any ncol): lane-per-column DFMA.
// Which of dmma / fma is faster is a measured property of the part (profiles/): the default follows it.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace g4s {

static int g_bsr_variant = 0;  // 0 auto, 1 fma, 2 dmma, 3 generic, 4 K-packed fma

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) bsr3_spmm64_dmma_kernel(int mb, const int *__restrict__ browptr,
                                                               const int *__restrict__ bcolids,
                                                               const double *__restrict__ bvalues,
                                                               const double *__restrict__ B, double *__restrict__ C) {
    const int lane = threadIdx.x & 31;
    const int k = lane & 3, q = lane >> 2;  // k: position along K inside a step; q: row of the A tile / col of the B tile
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < mb; I += (gridDim.x * blockDim.x) >> 5) {
        const int p0 = __ldg(browptr + I), p1 = __ldg(browptr + I + 1);
        const int K = 3 * (p1 - p0);
        double acc[8][2];
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = 0.0;
        for (int kk0 = 0; kk0 < K; kk0 += 4) {
            const int kk = kk0 + k;
            const bool valid = kk < K;
            const int p = p0 + kk / 3, s = kk % 3;
            double bfrag = 0.0;  // A_I^T[kk][q] = block p, entry (q, s)
            const double *brow = B;
            if (valid) {
                const int J = __ldg(bcolids + p);
                brow = B + ((size_t)J * 3 + s) * 64 + q;
                if (q < 3) bfrag = __ldg(bvalues + (size_t)p * 9 + q * 3 + s);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double afrag = valid ? __ldg(brow + 8 * m) : 0.0;  // Bg^T[8m + q][kk]
                dmma_m8n8k4(acc[m][0], acc[m][1], afrag, bfrag);
            }
        }
        // D tile m: lane holds rows (dense column) 8m + q, cols (block row r) 2k and 2k+1
        double *crow = C + (size_t)I * 3 * 64;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (k == 0) {
                crow[0 * 64 + 8 * m + q] = acc[m][0];
                crow[1 * 64 + 8 * m + q] = acc[m][1];
            } else if (k == 1) {
                crow[2 * 64 + 8 * m + q] = acc[m][0];
            }
        }
    }
}

__global__ void __launch_bounds__(256) bsr3_spmm64_fma_kernel(int mb, const int *__restrict__ browptr,
                                                              const int *__restrict__ bcolids,
                                                              const double *__restrict__ bvalues,
                                                              const double *__restrict__ B, double *__restrict__ C) {
    const int lane = threadIdx.x & 31;
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < mb; I += (gridDim.x * blockDim.x) >> 5) {
        const int p0 = __ldg(browptr + I), p1 = __ldg(browptr + I + 1);
        double2 c0 = make_double2(0, 0), c1 = c0, c2 = c0;
        for (int p = p0; p < p1; ++p) {
            const int J = __ldg(bcolids + p);
            const double *blk = bvalues + (size_t)p * 9;
            const double2 *b = reinterpret_cast<const double2 *>(B + (size_t)J * 3 * 64) + lane;
            const double2 b0 = __ldg(b), b1 = __ldg(b + 32), b2 = __ldg(b + 64);
            const double a00 = __ldg(blk), a01 = __ldg(blk + 1), a02 = __ldg(blk + 2), a10 = __ldg(blk + 3),
                         a11 = __ldg(blk + 4), a12 = __ldg(blk + 5), a20 = __ldg(blk + 6), a21 = __ldg(blk + 7),
                         a22 = __ldg(blk + 8);
            c0.x = fma(a00, b0.x, c0.x); c0.y = fma(a00, b0.y, c0.y);
            c0.x = fma(a01, b1.x, c0.x); c0.y = fma(a01, b1.y, c0.y);
            c0.x = fma(a02, b2.x, c0.x); c0.y = fma(a02, b2.y, c0.y);
            c1.x = fma(a10, b0.x, c1.x); c1.y = fma(a10, b0.y, c1.y);
            c1.x = fma(a11, b1.x, c1.x); c1.y = fma(a11, b1.y, c1.y);
            c1.x = fma(a12, b2.x, c1.x); c1.y = fma(a12, b2.y, c1.y);
            c2.x = fma(a20, b0.x, c2.x); c2.y = fma(a20, b0.y, c2.y);
            c2.x = fma(a21, b1.x, c2.x); c2.y = fma(a21, b1.y, c2.y);
            c2.x = fma(a22, b2.x, c2.x); c2.y = fma(a22, b2.y, c2.y);
        }
        double2 *c = reinterpret_cast<double2 *>(C + (size_t)I * 3 * 64) + lane;
        c[0] = c0;
        c[32] = c1;
        c[64] = c2;
    }
}

// Multi-GPU form of the DFMA kernel (one NVSwitch box, <= 8 GPUs): the dense operand B is row-partitioned like the block
// rows, every rank keeps its slice in CUDA-IPC-shared memory, and the kernel reads the rows of B that other GPUs own
// straight over NVLink — no halo exchange, no replicated B.  base[q] points at block row cut[q] of B.
__global__ void __launch_bounds__(256) bsr3_spmm64_fma_parts_kernel(int mb, const int *__restrict__ browptr,
                                                                    const int *__restrict__ bcolids,
                                                                    const double *__restrict__ bvalues,
                                                                    const BParts bp, double *__restrict__ C) {
    const int lane = threadIdx.x & 31;
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < mb; I += (gridDim.x * blockDim.x) >> 5) {
        const int p0 = __ldg(browptr + I), p1 = __ldg(browptr + I + 1);
        double2 c0 = make_double2(0, 0), c1 = c0, c2 = c0;
        for (int p = p0; p < p1; ++p) {
            const int J = __ldg(bcolids + p);
            const double *blk = bvalues + (size_t)p * 9;
            const double2 *b = reinterpret_cast<const double2 *>(b_part_row(bp, J)) + lane;
            const double2 b0 = __ldg(b), b1 = __ldg(b + 32), b2 = __ldg(b + 64);
            const double a00 = __ldg(blk), a01 = __ldg(blk + 1), a02 = __ldg(blk + 2), a10 = __ldg(blk + 3),
                         a11 = __ldg(blk + 4), a12 = __ldg(blk + 5), a20 = __ldg(blk + 6), a21 = __ldg(blk + 7),
                         a22 = __ldg(blk + 8);
            c0.x = fma(a00, b0.x, c0.x); c0.y = fma(a00, b0.y, c0.y);
            c0.x = fma(a01, b1.x, c0.x); c0.y = fma(a01, b1.y, c0.y);
            c0.x = fma(a02, b2.x, c0.x); c0.y = fma(a02, b2.y, c0.y);
            c1.x = fma(a10, b0.x, c1.x); c1.y = fma(a10, b0.y, c1.y);
            c1.x = fma(a11, b1.x, c1.x); c1.y = fma(a11, b1.y, c1.y);
            c1.x = fma(a12, b2.x, c1.x); c1.y = fma(a12, b2.y, c1.y);
            c2.x = fma(a20, b0.x, c2.x); c2.y = fma(a20, b0.y, c2.y);
            c2.x = fma(a21, b1.x, c2.x); c2.y = fma(a21, b1.y, c2.y);
            c2.x = fma(a22, b2.x, c2.x); c2.y = fma(a22, b2.y, c2.y);
        }
        double2 *c = reinterpret_cast<double2 *>(C + (size_t)I * 3 * 64) + lane;
        c[0] = c0;
        c[32] = c1;
        c[64] = c2;
    }
}

// K-packed DFMA kernel (bs = 3, ncol = 64), the single-GPU default.  ncu on the kernel above shows the LSU register
// WRITEBACK (128 B per cycle per SM) 75-80 % busy, and most of it is the block values: a warp-uniform LDG.64 still writes
// 8 bytes into each of 32 lanes, so the 9 broadcast values of a block cost 18 writeback cycles against 12 for the 1536
// bytes of B the block really needs (32 data-pipe wavefronts per block measured; 64-bit instead of 128-bit loads of B
// change nothing, profiles/r01_other_kernels_summary.md).  Here the row's blocks are read as one flat list of scalar
// columns kk = 3 p + s (K of the product C_I = A_I[3 x K] Bg[K x 64]); lane group g = lane / 8 takes kk = g (mod 4),
// so a lane needs only the THREE values A[0..2][kk] of its own column (3 LDG.64 per step of four columns instead of
// 9 per block) and the 64 entries of row kk of Bg are split over the group's 8 lanes (columns 16 i + 2 j, +1 for
// i = 0..3: every LDG.128 of a group is one 128-byte line).  Each lane accumulates 3 x 8 partial sums over its share
// of K; the four groups are summed once per block row with a halving exchange (36 shuffles), after which lane
// (g, j) holds C_I[0..2][16 g + 2 j, +1] and the warp stores three contiguous 512-byte rows.  Two steps are kept in
// flight (120 registers, 16 warps per SM).  Measured: 20.3 data-pipe wavefronts per block instead of 32.2.
// loads that do not allocate in L1 (block values and column ids are used once; L1 is kept for the rows of B)
template <bool STREAM>
__device__ __forceinline__ double kpack_ld_a(const double *p) {
    if (!STREAM) return __ldg(p);
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <bool STREAM>
__device__ __forceinline__ int kpack_ld_j(const int *p) {
    if (!STREAM) return __ldg(p);
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Asks for the block values and column ids of block row `In` (the row this warp takes next) ahead of time: they are a
// DRAM stream that is used exactly once, so without this every step of the row waits a full DRAM latency for them.
// One prefetch instruction covers 32 lines of 128 bytes.  level 1: into L2, 2: into L1.
__device__ __forceinline__ void kpack_prefetch_row(int level, int pn0, int pn1, const int *bcolids, const double *bvalues) {
    const int lane = threadIdx.x & 31;
    const char *v0 = reinterpret_cast<const char *>(bvalues + (size_t)pn0 * 9);
    const char *v1 = reinterpret_cast<const char *>(bvalues + (size_t)pn1 * 9);
    const char *c0 = reinterpret_cast<const char *>(bcolids + pn0);
    const char *c1 = reinterpret_cast<const char *>(bcolids + pn1);
    for (const char *q = v0 + 128 * lane; q < v1; q += 128 * 32) {
        if (level == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
    for (const char *q = c0 + 128 * lane; q < c1; q += 128 * 32) {
        if (level == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
}

// one block row I by one warp; In = the row the warp takes next (or -1), prefetched at level pf
template <bool PARTS, bool STREAM>
__device__ __forceinline__ void kpack_row(int I, int In, int pf, const int *__restrict__ browptr,
                                          const int *__restrict__ bcolids, const double *__restrict__ bvalues,
                                          const double *__restrict__ B, const BParts &bp, double *__restrict__ C) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    {
        const int p0 = __ldg(browptr + I), p1 = __ldg(browptr + I + 1);
        int pn0 = 0, pn1 = 0;
        if (pf && In >= 0) {
            pn0 = __ldg(browptr + In);
            pn1 = __ldg(browptr + In + 1);
        }
        double2 acc[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = make_double2(0.0, 0.0);
        // this lane's scalar column: block p, column s of the block; a step of 4 columns is one block and one column on
        int p = p0 + (g == 3 ? 1 : 0), s = g == 3 ? 0 : g;
        int Jn = p < p1 ? kpack_ld_j<STREAM>(bcolids + p) : 0;
        struct Step {
            double a0, a1, a2;
            double2 b[4];
        };
        // issues the loads of this lane's current column, moves on by four columns and fetches the next column id;
        // returns whether any lane of the warp had a column left (warp-uniform)
        auto load = [&](Step &st) -> bool {
            const bool valid = p < p1;
            const double *blk = bvalues + (size_t)p * 9 + s;
            const double *brow = (PARTS ? b_part_row_lane(bp, Jn) : B + (size_t)Jn * 192) + s * 64 + 2 * j;
            st.a0 = st.a1 = st.a2 = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) st.b[i] = make_double2(0.0, 0.0);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 4; ++i) st.b[i] = __ldg(reinterpret_cast<const double2 *>(brow + 16 * i));
                st.a0 = kpack_ld_a<STREAM>(blk);
                st.a1 = kpack_ld_a<STREAM>(blk + 3);
                st.a2 = kpack_ld_a<STREAM>(blk + 6);
            }
            ++p;
            ++s;
            if (s >= 3) {
                s -= 3;
                ++p;
            }
            Jn = p < p1 ? kpack_ld_j<STREAM>(bcolids + p) : 0;
            return __any_sync(FULL, valid);
        };
        auto mac = [&](const Step &st) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[0][i].x = fma(st.a0, st.b[i].x, acc[0][i].x); acc[0][i].y = fma(st.a0, st.b[i].y, acc[0][i].y);
                acc[1][i].x = fma(st.a1, st.b[i].x, acc[1][i].x); acc[1][i].y = fma(st.a1, st.b[i].y, acc[1][i].y);
                acc[2][i].x = fma(st.a2, st.b[i].x, acc[2][i].x); acc[2][i].y = fma(st.a2, st.b[i].y, acc[2][i].y);
            }
        };
        // two steps in flight: the loads of the next step are issued before the multiply-adds of the current one
        Step sa, sb;
        bool more = load(sa);
        if (pf && In >= 0) kpack_prefetch_row(pf, pn0, pn1, bcolids, bvalues);
        while (more) {
            const bool more_b = load(sb);
            mac(sa);
            if (!more_b) break;
            more = load(sa);
            mac(sb);
        }
        // sum the four K groups: exchange halves with lane ^ 16, then with lane ^ 8; lane (g, j) ends with chunk i = g
        const bool hi = (g & 2) != 0, lo = (g & 1) != 0;
        double2 f[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double2 t[2];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const double2 send = hi ? acc[r][ii] : acc[r][ii + 2];
                const double2 keep = hi ? acc[r][ii + 2] : acc[r][ii];
                t[ii].x = keep.x + __shfl_xor_sync(FULL, send.x, 16);
                t[ii].y = keep.y + __shfl_xor_sync(FULL, send.y, 16);
            }
            const double2 send = lo ? t[0] : t[1];
            const double2 keep = lo ? t[1] : t[0];
            f[r].x = keep.x + __shfl_xor_sync(FULL, send.x, 8);
            f[r].y = keep.y + __shfl_xor_sync(FULL, send.y, 8);
        }
        double2 *c = reinterpret_cast<double2 *>(C + (size_t)I * 192) + lane;
        c[0] = f[0];
        c[32] = f[1];
        c[64] = f[2];
    }
}

template <bool PARTS>
__global__ void __launch_bounds__(256, 2) bsr3_spmm64_kpack_kernel(int mb, const int *__restrict__ browptr,
                                                                   const int *__restrict__ bcolids,
                                                                   const double *__restrict__ bvalues,
                                                                   const double *__restrict__ B,
                                                                   const __grid_constant__ BParts bp,
                                                                   double *__restrict__ C, int pf) {
    const int stride = (gridDim.x * blockDim.x) >> 5;
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < mb; I += stride)
        kpack_row<PARTS, false>(I, I + stride < mb ? I + stride : -1, pf, browptr, bcolids, bvalues, B, bp, C);
}

// Ordered form: block rows are taken in the order the caller gives (row_order, a permutation of 0..mb-1), cut into TILES
// (tile_ptr); one CTA of 16 warps per SM takes tiles blockIdx, blockIdx + grid, ... and works through a tile 16 entries
// (one per warp) at a time.  With the tile-major order of g4s_grid_pencil_order (a 4 x 4 patch of mesh nodes swept along
// the third axis = one tile) the 16 rows in flight on an SM share their rows of B, which then live in that SM's L1 (no
// shared memory is used: the carve-out is all L1) instead of being fetched from L2 once per mesh line; neighbouring
// pencils are swept by other SMs at the same time, so the halo rows they share are in L2 when the second SM asks.
// Block values and column ids bypass L1 and are prefetched one row ahead.
template <bool PARTS>
__global__ void __launch_bounds__(512, 1) bsr3_spmm64_kpack_ordered_kernel(int mb, const int *__restrict__ row_order,
                                                                           const int *__restrict__ tile_ptr, int ntiles,
                                                                           const int *__restrict__ browptr,
                                                                           const int *__restrict__ bcolids,
                                                                           const double *__restrict__ bvalues,
                                                                           const double *__restrict__ B,
                                                                           const __grid_constant__ BParts bp,
                                                                           double *__restrict__ C, int pf) {
    const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        long long r0, r1;
        if (tile_ptr) {
            r0 = __ldg(tile_ptr + t);
            r1 = __ldg(tile_ptr + t + 1);
        } else {  // no tiles given: one contiguous stretch of the order per CTA
            const long long per = ((long long)mb + ntiles - 1) / ntiles;
            r0 = per * t;
            r1 = r0 + per < mb ? r0 + per : mb;
        }
        for (long long idx = r0 + warp; idx < r1; idx += wpc) {
            const int In = idx + wpc < r1 ? __ldg(row_order + idx + wpc) : -1;
            kpack_row<PARTS, true>(__ldg(row_order + idx), In, pf, browptr, bcolids, bvalues, B, bp, C);
        }
    }
}

template <int BS>
__global__ void __launch_bounds__(256) bsr_spmm_generic_kernel(int mb, const int *__restrict__ browptr,
                                                               const int *__restrict__ bcolids,
                                                               const double *__restrict__ bvalues, int ncol,
                                                               const double *__restrict__ B, double *__restrict__ C) {
    const int lane = threadIdx.x & 31;
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < mb; I += (gridDim.x * blockDim.x) >> 5) {
        const int p0 = __ldg(browptr + I), p1 = __ldg(browptr + I + 1);
        for (int c = lane; c < ncol; c += 32) {
            double acc[BS];
#pragma unroll
            for (int r = 0; r < BS; ++r) acc[r] = 0.0;
            for (int p = p0; p < p1; ++p) {
                const double *blk = bvalues + (size_t)p * BS * BS;
                const double *b = B + (size_t)__ldg(bcolids + p) * BS * ncol + c;
#pragma unroll
                for (int s = 0; s < BS; ++s) {
                    const double bv = __ldg(b + (size_t)s * ncol);
#pragma unroll
                    for (int r = 0; r < BS; ++r) acc[r] = fma(__ldg(blk + r * BS + s), bv, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < BS; ++r) C[((size_t)I * BS + r) * ncol + c] = acc[r];
        }
    }
}

// prefetch level of the K-packed kernels (G4S_BSR_PF overrides: 0 none, 1 L2, 2 L1)
static int kpack_pf() {
    static const int v = [] {
        const char *e = getenv("G4S_BSR_PF");
        return e ? atoi(e) : 2;
    }();
    return v;
}

template <int BS>
static void launch_generic(int grid, cudaStream_t st, int mb, const int *rp, const int *ci, const double *va, int ncol,
                           const double *B, double *C) {
    bsr_spmm_generic_kernel<BS><<<grid, 256, 0, st>>>(mb, rp, ci, va, ncol, B, C);
}

}  // namespace g4s

using namespace g4s;

extern "C" {

int g4s_bsr_spmm_set_variant(int variant) {
    if (variant < 0 || variant > 4)
        return fail(G4S_ERR_INVALID, "variant must be 0 (auto), 1 (fma), 2 (dmma), 3 (generic) or 4 (K-packed fma)");
    g_bsr_variant = variant;
    return G4S_OK;
}

static int partitioned_args(int mb_local, const int *browptr_dev, const double *const *B_parts, const int *cuts,
                            double *C_dev, int world, BParts &bp) {
    if (mb_local < 0 || !browptr_dev || !B_parts || !cuts || !C_dev || world < 1 || world > 8)
        return fail(G4S_ERR_INVALID, "g4s_bsr3_spmm64_partitioned: bad arguments (1 <= world <= 8)");
    for (int q = 0; q < 8; ++q) bp.base[q] = q < world ? B_parts[q] : nullptr;
    for (int q = 0; q <= 8; ++q) bp.cut[q] = cuts[q < world ? q : world];
    bp.world = world;
    return G4S_OK;
}

int g4s_bsr3_spmm64_partitioned_device(int mb_local, const int *browptr_dev, const int *bcolids_dev,
                                       const double *bvalues_dev, int world, const double *const *B_parts,
                                       const int *cuts, double *C_dev, void *stream) {
    BParts bp;
    int rc = partitioned_args(mb_local, browptr_dev, B_parts, cuts, C_dev, world, bp);
    if (rc) return rc;
    if (mb_local == 0) return G4S_OK;
    if ((rc = ensure_device())) return rc;
    const int grid = (int)std::min<long long>(((long long)mb_local * 32 + 255) / 256, (long long)sm_count() * 16);
    if (g_bsr_variant == 4)  // K-packed kernel: opt-in here (g4s_bsr_spmm_set_variant(4)); plain DFMA is the measured default
        bsr3_spmm64_kpack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(mb_local, browptr_dev, bcolids_dev,
                                                                             bvalues_dev, nullptr, bp, C_dev, kpack_pf());
    else
        bsr3_spmm64_fma_parts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mb_local, browptr_dev, bcolids_dev, bvalues_dev,
                                                                           bp, C_dev);
    G4S_CHECK_LAUNCH("bsr3_spmm64 partitioned kernel");
    return G4S_OK;
}

int g4s_bsr3_spmm64_partitioned_ordered_device(int mb_local, const int *browptr_dev, const int *bcolids_dev,
                                               const double *bvalues_dev, int world, const double *const *B_parts,
                                               const int *cuts, double *C_dev, const int *row_order_dev, void *stream) {
    BParts bp;
    int rc = partitioned_args(mb_local, browptr_dev, B_parts, cuts, C_dev, world, bp);
    if (rc) return rc;
    if (!row_order_dev) return fail(G4S_ERR_INVALID, "g4s_bsr3_spmm64_partitioned_ordered_device: null row order");
    if (mb_local == 0) return G4S_OK;
    if ((rc = ensure_device())) return rc;
    auto k = bsr3_spmm64_kpack_ordered_kernel<true>;
    static PerDeviceOnce configured;
    if (configured.needs()) {
        G4S_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
        configured.done();
    }
    const int grid = (int)std::min<long long>(((long long)mb_local + 15) / 16, (long long)sm_count());
    k<<<grid, 512, 0, (cudaStream_t)stream>>>(mb_local, row_order_dev, nullptr, grid, browptr_dev, bcolids_dev, bvalues_dev,
                                              nullptr, bp, C_dev, kpack_pf());
    G4S_CHECK_LAUNCH("bsr3_spmm64_kpack_ordered_kernel");
    return G4S_OK;
}

int g4s_bsr3_spmm64_ordered_device(int mb, int kb, const int *browptr_dev, const int *bcolids_dev, const double *bvalues_dev,
                                   const double *B_dev, double *C_dev, const int *row_order_dev, const int *tile_ptr_dev,
                                   int ntiles, void *stream) {
    if (mb < 0 || kb < 0 || !browptr_dev || !B_dev || !C_dev || !row_order_dev ||
        ((reinterpret_cast<uintptr_t>(B_dev) | reinterpret_cast<uintptr_t>(C_dev)) & 15))
        return fail(G4S_ERR_INVALID, "g4s_bsr3_spmm64_ordered_device: bad arguments (B and C 16-byte aligned)");
    if (mb == 0) return G4S_OK;
    int rc = ensure_device();
    if (rc) return rc;
    auto k = bsr3_spmm64_kpack_ordered_kernel<false>;
    static PerDeviceOnce configured;
    if (configured.needs()) {  // no shared memory: the whole carve-out is L1
        G4S_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
        configured.done();
    }
    if (tile_ptr_dev && ntiles < 1) return fail(G4S_ERR_INVALID, "g4s_bsr3_spmm64_ordered_device: tile_ptr without tiles");
    const int grid = (int)std::min<long long>(tile_ptr_dev ? ntiles : ((long long)mb + 15) / 16, (long long)sm_count());
    BParts none = {};
    k<<<grid, 512, 0, (cudaStream_t)stream>>>(mb, row_order_dev, tile_ptr_dev, tile_ptr_dev ? ntiles : grid, browptr_dev,
                                              bcolids_dev, bvalues_dev, B_dev, none, C_dev, kpack_pf());
    G4S_CHECK_LAUNCH("bsr3_spmm64_kpack_ordered_kernel");
    return G4S_OK;
}

int g4s_grid_pencil_order(int n0, int n1, int n2, int p0, int p1, int *order, int *tile_ptr, int *ntiles) {
    if (n0 < 1 || n1 < 1 || n2 < 1 || p0 < 1 || p1 < 1 || !order || (long long)n0 * n1 * n2 > 2147483647LL)
        return fail(G4S_ERR_INVALID, "g4s_grid_pencil_order: bad arguments");
    long long at = 0;
    int t = 0;
    for (int b1 = 0; b1 < n1; b1 += p1)
        for (int b0 = 0; b0 < n0; b0 += p0) {
            if (tile_ptr) tile_ptr[t] = (int)at;
            ++t;
            for (int k = 0; k < n2; ++k)
                for (int j = b1; j < std::min(n1, b1 + p1); ++j)
                    for (int i = b0; i < std::min(n0, b0 + p0); ++i) order[at++] = (int)(((long long)k * n1 + j) * n0 + i);
        }
    if (tile_ptr) tile_ptr[t] = (int)at;
    if (ntiles) *ntiles = t;
    return G4S_OK;
}

int g4s_bsr_spmm_device(int mb, int kb, int bs, const int *browptr_dev, const int *bcolids_dev,
                        const double *bvalues_dev, int ncol, const double *B_dev, double *C_dev, void *stream) {
    if (mb < 0 || kb < 0 || bs < 1 || bs > 8 || ncol < 1 || !browptr_dev || !B_dev || !C_dev)
        return fail(G4S_ERR_INVALID, "g4s_bsr_spmm_device: bad arguments (1 <= bs <= 8)");
    if (mb == 0) return G4S_OK;
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::min<long long>(((long long)mb * 32 + 255) / 256, (long long)sm_count() * 16);
    int variant = g_bsr_variant;
    const bool fast_ok = bs == 3 && ncol == 64 && ((reinterpret_cast<uintptr_t>(B_dev) | reinterpret_cast<uintptr_t>(C_dev)) & 15) == 0;
    if (variant == 0) variant = fast_ok ? 4 : 3;
    if (variant != 3 && !fast_ok) variant = 3;
    if (variant == 4) {
        BParts none = {};
        bsr3_spmm64_kpack_kernel<false><<<grid, 256, 0, st>>>(mb, browptr_dev, bcolids_dev, bvalues_dev, B_dev, none, C_dev,
                                                              kpack_pf());
    } else if (variant == 2) {
        bsr3_spmm64_dmma_kernel<<<grid, 256, 0, st>>>(mb, browptr_dev, bcolids_dev, bvalues_dev, B_dev, C_dev);
    } else if (variant == 1) {
        bsr3_spmm64_fma_kernel<<<grid, 256, 0, st>>>(mb, browptr_dev, bcolids_dev, bvalues_dev, B_dev, C_dev);
    } else {
        switch (bs) {
            case 1: launch_generic<1>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 2: launch_generic<2>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 3: launch_generic<3>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 4: launch_generic<4>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 5: launch_generic<5>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 6: launch_generic<6>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            case 7: launch_generic<7>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
            default: launch_generic<8>(grid, st, mb, browptr_dev, bcolids_dev, bvalues_dev, ncol, B_dev, C_dev); break;
        }
    }
    G4S_CHECK_LAUNCH("bsr_spmm kernel");
    return G4S_OK;
}

}  // extern "C"
