// BSR 3x3 SpMM x 64 columns as a SLIDING-WINDOW SWEEP (inspector / executor), sm_100a.
//
// Why: the row-wise kernels of bsr_spmm.cu end at 20 % of the HBM roofline although DRAM is not saturated
// (profiles/r01_other_kernels_summary.md): what saturates is the SM's 128 B/cycle data path into registers.  A block
// contributes 9 x 64 multiply-adds = 9 cycles of the FP64 pipe per SM, but feeding it row by row costs 12 cycles for the
// block's 1536 bytes of B plus 4.5-18 cycles for its 9 values (a warp-uniform value is still written into every lane).
// Registers are the only place where reuse is free, so this kernel keeps THREE block rows of C in registers at once:
//
//   * a STRIP is a list of block rows swept in order by one group of 8 lanes (lane l owns columns 16 i + 2 l, +1 of the
//     64; the 4 groups of a warp sweep 4 strips, a CTA of 4 consumer warps sweeps a TILE of 16 strips);
//   * at phase q the group holds rows q-1, q, q+1 of its strip in three accumulator slots (slot f = row q-1+f) and
//     walks the block columns J that these rows reference: row J of B is loaded ONCE (3 x 4 LDS.128) and multiplied
//     into all three slots — for a 27-point mesh operator swept along a grid line every loaded row of B feeds three
//     blocks, so B costs 4 cycles per block instead of 12; after the phase row q-1 is complete, stored and its slot is
//     reused for row q+2;
//   * the inspector (host, once per sparsity pattern) turns every (phase, column) into a STEP: the position u of row J of
//     B in the CTA's staged list and a 3 x 3 x 3 record of values (sub-step-major, 240 bytes; absent blocks are explicit
//     zeros), and cuts the sweep of a tile into STAGES sized for shared memory.  Stage = [header | u lists | records],
//     one contiguous chunk of the plan stream, so the block values are read from DRAM strictly sequentially;
//   * a producer warp streams stages with the TMA engine (cp.async.bulk, SASS UBLKCP): one bulk copy for the chunk
//     (L2 evict_first) and one per run of consecutive rows of B in the stage's union list (1536 B per row; for mesh
//     operators a run is a 9 KB line of 6 nodes), double-buffered on full/empty mbarriers; its copy descriptors are one
//     coalesced int2 per lane from the plan's table, fetched a stage early.  The consumers never touch global memory
//     except to store finished rows of C.
//   * the plan fixes the grid (one CTA per SM of the device it was made on) and stores every CTA's stages consecutively,
//     tiles dealt round-robin so that neighbouring patches are swept at the same time and share their halo rows of B
//     through L2 (ncu: DRAM traffic 1.08x the algorithmic bytes).
//   Measured alternatives (profiles/r02_bsr_sweep_summary.md): two 4-warp CTAs per SM with the loader folded into warp 0
//   need stages of a third of a phase to fit shared memory and lose more to per-stage overhead and shallow prefetch
//   (3.6-4.6 ms) than two warps per scheduler gain; a single warp per scheduler already issues a DFMA every 2.64 cycles
//   against 2.46 with two (tools/micro/dfma_rate.cu), i.e. the FP64 pipe of this part tops out near 30 TFLOP/s.
//
// Per block the data path now carries 4 (B) + 5 (values, 8-lane broadcast) cycles against 9 cycles of DFMA.
// Arithmetic: plain FP64 FMAs in a fixed, schedule-defined order (deterministic; differs from the row-order sum of the
// oracle by rounding only).  Absent blocks are stored as zeros, as a BSR matrix with explicit zero blocks would hold
// them: B must be finite (0 x inf would leak a NaN into a row that does not reference that row of B).
//
// Reference semantics: C = A B for the node operator of citcoms/lib/Element_calculations.c:516-571;
// the reference has no BSR routine, the oracle's restatement (oracle/oracle_spmv.c: oracle_bsr_spmm) is the checker.
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace g4s {

constexpr int SW_WARPS = 4;                 // consumer warps per CTA
constexpr int SW_STRIPS = SW_WARPS * 4;     // strips per tile
constexpr int SW_HDR_INTS = 64;             // stage header: [0] runs [1] steps [2] slide the window [3] byte offset of the B area
constexpr int SW_HDR_ROWS = 4;              //               [4..19] row that leaves the window, per strip (-1: none)
constexpr int SW_HDR_RUNS = 20;             //               [20 + 2 i], [21 + 2 i]: first block column and length of run i
constexpr int SW_MAXRUNS = 22;
constexpr int SW_REC = 30;                  // doubles per step record: 3 sub-steps x (3 slots x 3 block rows + 1 spare)
constexpr int SW_NODE = 192;                // doubles per block row of B (3 x 64)

struct SweepStage {
    long long off;    // byte offset of the stage's chunk in the plan stream (16-byte aligned)
    int chunk_bytes;  // header + u lists + records
    int tx_bytes;     // chunk + staged rows of B: what the stage's mbarrier expects
};

struct SweepArgs {
    const unsigned char *stream;
    const int2 *prod;     // per stage 32 entries: lane i < 30: run i (first block column, length | first staged row << 8);
                          // lane 30: byte offset of the chunk (lo, hi); lane 31: (chunk bytes, bytes the mbarrier expects)
    const int *cta_ptr;   // stages of CTA b: [cta_ptr[b], cta_ptr[b+1])
    int stage_smem;       // bytes of one stage buffer (two are allocated)
};

__device__ __forceinline__ void sweep_flush(double2 (&acc)[3][4], int row, double *__restrict__ C, int l) {
    if (row >= 0) {
        double2 *c = reinterpret_cast<double2 *>(C + (size_t)row * SW_NODE) + l;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[r * 32 + i * 8] = acc[r][i];
    }
}

// warp 0: start the copies of one stage into `buf` (pd = this lane's entry of the stage's descriptor row)
template <bool PARTS>
__device__ __forceinline__ void sweep_issue(unsigned char *buf, uint64_t *full, const int2 pd, const SweepArgs &a,
                                            const double *__restrict__ B, const BParts &bp, int lane) {
    const int chunk_bytes = __shfl_sync(0xffffffffu, pd.x, 31), tx_bytes = __shfl_sync(0xffffffffu, pd.y, 31);
    const unsigned off_lo = (unsigned)__shfl_sync(0xffffffffu, pd.x, 30), off_hi = (unsigned)__shfl_sync(0xffffffffu, pd.y, 30);
    if (lane == 0) {
        const unsigned long long off = ((unsigned long long)off_hi << 32) | off_lo;
        mbar_arrive_expect_tx(full, (uint32_t)tx_bytes);
        bulk_g2s(buf, a.stream + off, (uint32_t)chunk_bytes, full, policy_evict_first());
    }
    __syncwarp();
    const int len = lane < 30 ? (pd.y & 0xff) : 0;
    if (len > 0) {
        const int J0 = pd.x, pre = pd.y >> 8;
        const double *src = PARTS ? b_part_row_lane(bp, J0) : B + (size_t)J0 * SW_NODE;
        bulk_g2s(buf + ((chunk_bytes + 127) & ~127) + (size_t)pre * (SW_NODE * 8), src, (uint32_t)len * (SW_NODE * 8), full,
                 policy_evict_last());
    }
}

// one stage by one consumer warp; ROT = how often the window has slid so far (mod 3): relative slot f of the records lives
// in accumulator set (f + ROT) % 3, so sliding the window renames register sets at compile time instead of moving 48
// register pairs (ncu: the moves were 7 % of the consumers' issue slots)
template <int ROT>
__device__ __forceinline__ void sweep_stage(double2 (&acc)[3][3][4], const unsigned char *sb, int strip, int l,
                                            double *__restrict__ C) {
    const int *hdr = reinterpret_cast<const int *>(sb);
    const int S = hdr[1], rotate = hdr[2], b_off = hdr[3];
    const int Spad = (S + 3) & ~3;
    const int *ul = reinterpret_cast<const int *>(sb + SW_HDR_INTS * 4) + strip * Spad;
    const double *rec = reinterpret_cast<const double *>(sb + SW_HDR_INTS * 4 + SW_STRIPS * 4 * Spad) + (size_t)strip * S * SW_REC;
    const double *bb = reinterpret_cast<const double *>(sb + b_off) + 2 * l;
    int u = S > 0 ? ul[0] : 0;
    for (int t = 0; t < S; ++t) {
        const double *br = bb + u * SW_NODE;
        if (t + 1 < S) u = ul[t + 1];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            double av[10];
#pragma unroll
            for (int h = 0; h < 5; ++h) {
                const double2 v = *reinterpret_cast<const double2 *>(rec + s * 10 + 2 * h);
                av[2 * h] = v.x;
                av[2 * h + 1] = v.y;
            }
            double2 bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) bv[i] = *reinterpret_cast<const double2 *>(br + s * 64 + 16 * i);
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[(f + ROT) % 3][r][i].x = fma(av[f * 3 + r], bv[i].x, acc[(f + ROT) % 3][r][i].x);
                        acc[(f + ROT) % 3][r][i].y = fma(av[f * 3 + r], bv[i].y, acc[(f + ROT) % 3][r][i].y);
                    }
        }
        rec += SW_REC;
    }
    if (rotate) {  // end of a phase: the oldest row of the window is complete — store it; its set becomes the newest slot
        sweep_flush(acc[ROT % 3], hdr[SW_HDR_ROWS + strip], C, l);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[ROT % 3][r][i] = make_double2(0.0, 0.0);
    }
}

template <bool PARTS>
__global__ void __launch_bounds__((SW_WARPS + 1) * 32, 1)
    bsr3_sweep_kernel(const SweepArgs a, const double *__restrict__ B, const __grid_constant__ BParts bp,
                      double *__restrict__ C) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    __shared__ uint64_t bar_full[2], bar_empty[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = __ldg(a.cta_ptr + blockIdx.x), q1 = __ldg(a.cta_ptr + blockIdx.x + 1);
    if (threadIdx.x == 0) {
        mbar_init(&bar_full[0], 1);
        mbar_init(&bar_full[1], 1);
        mbar_init(&bar_empty[0], SW_WARPS);
        mbar_init(&bar_empty[1], SW_WARPS);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == SW_WARPS) {
        // ---- producer: stage q into buffer (q - q0) & 1 as soon as the consumers have released it ------------------------
        int2 pd = q0 < q1 ? __ldg(a.prod + (size_t)q0 * 32 + lane) : make_int2(0, 0);
        for (int q = q0; q < q1; ++q) {
            const unsigned it = (unsigned)(q - q0);
            const int st = it & 1;
            const int2 cur = pd;
            if (q + 1 < q1) pd = __ldg(a.prod + (size_t)(q + 1) * 32 + lane);  // the next stage's descriptors, a stage early
            if (it >= 2) mbar_wait(&bar_empty[st], ((it >> 1) - 1) & 1);
            sweep_issue<PARTS>(sw_smem + (size_t)st * a.stage_smem, &bar_full[st], cur, a, B, bp, lane);
        }
        return;
    }
    // ---- consumers ------------------------------------------------------------------------------------------------------
    const int g = lane >> 3, l = lane & 7, strip = warp * 4 + g;
    double2 acc[3][3][4];
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[f][r][i] = make_double2(0.0, 0.0);
    int rot = 0;
    for (int q = q0; q < q1; ++q) {
        const unsigned it = (unsigned)(q - q0);
        const int st = it & 1;
        const unsigned char *sb = sw_smem + (size_t)st * a.stage_smem;
        mbar_wait(&bar_full[st], (it >> 1) & 1);
        const int rotate = reinterpret_cast<const int *>(sb)[2];
        if (rot == 0) sweep_stage<0>(acc, sb, strip, l, C);
        else if (rot == 1) sweep_stage<1>(acc, sb, strip, l, C);
        else sweep_stage<2>(acc, sb, strip, l, C);
        if (rotate) rot = rot == 2 ? 0 : rot + 1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);
    }
}

// values of block p go to stream[base[p] + 10 s + r] for its entry (r, s): sub-step-major inside the step record
__global__ void sweep_set_values_kernel(const double *__restrict__ bvalues, const long long *__restrict__ base, long long nb,
                                        double *__restrict__ stream) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < nb; p += (long long)gridDim.x * blockDim.x) {
        const long long b = __ldg(base + p);
        const double *v = bvalues + p * 9;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) stream[b + 10 * s + r] = __ldg(v + r * 3 + s);
    }
}

// stage headers and u lists: meta[meta_off[q] ...) -> stream[stage q] (one CTA per stage)
__global__ void sweep_scatter_meta_kernel(const int *__restrict__ meta, const long long *__restrict__ meta_off,
                                          const SweepStage *__restrict__ stages, int nstages, unsigned char *__restrict__ stream) {
    const int q = blockIdx.x;
    if (q >= nstages) return;
    const long long m0 = meta_off[q], m1 = meta_off[q + 1];
    int *dst = reinterpret_cast<int *>(stream + stages[q].off);
    for (long long k = threadIdx.x; k < m1 - m0; k += blockDim.x) dst[k] = meta[m0 + k];
}

}  // namespace g4s

using namespace g4s;

struct g4s_bsr_plan {
    int mb = 0, kb = 0, world = 1;
    long long nb = 0;
    std::vector<int> cuts;
    unsigned char *stream = nullptr;
    size_t stream_bytes = 0;
    int2 *prod = nullptr;    // copy descriptors, 32 per stage
    int nstages = 0;
    int *cta_ptr = nullptr;  // [grid + 1]
    int grid = 0, ntiles = 0;
    long long *base = nullptr;
    int stage_smem = 0;
    long long total_steps = 0;  // strip-steps including padding (each moves 3 block slots)
    bool values_set = false;
};

namespace {

struct StepRec {
    int J;
    int p[3];
};
struct StageOut {
    int S = 0, U = 0;
    int hdr[SW_HDR_INTS];
    std::vector<int> ulist;                        // SW_STRIPS x Spad
    std::vector<std::pair<int, int>> blocks;       // (block p, double offset inside the stage's record area)
};

inline int stage_chunk_bytes(int S) { return SW_HDR_INTS * 4 + SW_STRIPS * 4 * ((S + 3) & ~3) + SW_STRIPS * S * SW_REC * 8; }
inline int round128(int v) { return (v + 127) & ~127; }

// The sweep of one tile (<= 16 strips): steps per (strip, phase), then stages.
void plan_tile(const int *browptr, const int *bcolids, const int *strip_ptr, const int *strip_rows, int s_begin, int s_end,
               int world, const int *cuts, int smem_budget, int max_steps, std::vector<StageOut> &out) {
    const int ns = s_end - s_begin;
    int P = 0;
    for (int s = s_begin; s < s_end; ++s) P = std::max(P, strip_ptr[s + 1] - strip_ptr[s]);
    if (P == 0) return;
    std::vector<std::vector<std::vector<StepRec>>> steps(ns, std::vector<std::vector<StepRec>>(P));
    std::vector<std::tuple<int, int, int>> e;
    for (int si = 0; si < ns; ++si) {
        const int *rows = strip_rows + strip_ptr[s_begin + si];
        const int m = strip_ptr[s_begin + si + 1] - strip_ptr[s_begin + si];
        e.clear();
        for (int a = 0; a < m; ++a)
            for (int p = browptr[rows[a]]; p < browptr[rows[a] + 1]; ++p) e.emplace_back(bcolids[p], a, p);
        std::sort(e.begin(), e.end());
        size_t i = 0;
        while (i < e.size()) {
            const int J = std::get<0>(e[i]), a0 = std::get<1>(e[i]);
            const int q = std::min(a0 + 1, m - 1);  // the phase whose window {q-1, q, q+1} starts at the earliest open row
            StepRec r{J, {-1, -1, -1}};
            while (i < e.size() && std::get<0>(e[i]) == J && std::get<1>(e[i]) <= q + 1) {
                const int slot = std::get<1>(e[i]) - (q - 1);  // relative to the window {q-1, q, q+1}
                if (r.p[slot] != -1) {  // a second block with the same (row, column): BSR built from duplicates
                    steps[si][q].push_back(r);
                    r = StepRec{J, {-1, -1, -1}};
                }
                r.p[slot] = std::get<2>(e[i]);
                ++i;
            }
            steps[si][q].push_back(r);
        }
    }
    auto owner_of = [&](int J) {
        int o = 0;
        for (int k = 1; k < world; ++k) o += J >= cuts[k];
        return o;
    };
    std::vector<int> uni;
    for (int q = 0; q <= P; ++q) {  // q == P: the closing stage that stores the last row of every strip
        int Smax = 0;
        if (q < P)
            for (int si = 0; si < ns; ++si) Smax = std::max(Smax, (int)steps[si][q].size());
        int start = 0;
        do {
            int cnt = std::min(max_steps, Smax - start);
            StageOut so;
            int nruns = 0;
            for (;;) {  // shrink the piece until its union of rows of B fits the stage buffer and the run table
                uni.clear();
                for (int si = 0; si < ns; ++si)
                    for (int t = start; t < std::min(start + cnt, q < P ? (int)steps[si][q].size() : 0); ++t)
                        uni.push_back(steps[si][q][t].J);
                std::sort(uni.begin(), uni.end());
                uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
                nruns = 0;
                for (size_t k = 0; k < uni.size();) {
                    size_t k2 = k + 1;
                    while (k2 < uni.size() && uni[k2] == uni[k2 - 1] + 1 && k2 - k < 64 &&
                           (world <= 1 || owner_of(uni[k2]) == owner_of(uni[k])))
                        ++k2;
                    if (nruns < SW_MAXRUNS) {
                        so.hdr[SW_HDR_RUNS + 2 * nruns] = uni[k];
                        so.hdr[SW_HDR_RUNS + 2 * nruns + 1] = (int)(k2 - k);
                    }
                    ++nruns;
                    k = k2;
                }
                const int need = round128(stage_chunk_bytes(cnt)) + (int)uni.size() * SW_NODE * 8;
                if ((nruns <= SW_MAXRUNS && need <= smem_budget) || cnt <= 1) break;
                cnt = std::max(1, cnt / 2);
            }
            for (int k = nruns; k < SW_MAXRUNS; ++k) so.hdr[SW_HDR_RUNS + 2 * k] = so.hdr[SW_HDR_RUNS + 2 * k + 1] = 0;
            so.S = cnt;
            so.U = (int)uni.size();
            const int Spad = (cnt + 3) & ~3;
            so.hdr[0] = nruns;
            so.hdr[1] = cnt;
            const bool last_of_phase = start + cnt >= Smax;
            so.hdr[2] = last_of_phase ? 1 : 0;  // slide the window after this stage
            so.hdr[3] = round128(stage_chunk_bytes(cnt));
            for (int si = 0; si < SW_STRIPS; ++si) {
                int row = -1;
                if (last_of_phase && q >= 1 && si < ns) {
                    const int m = strip_ptr[s_begin + si + 1] - strip_ptr[s_begin + si];
                    if (q - 1 < m) row = strip_rows[strip_ptr[s_begin + si] + q - 1];
                }
                so.hdr[SW_HDR_ROWS + si] = row;
            }
            so.ulist.assign((size_t)SW_STRIPS * Spad, 0);
            for (int si = 0; si < ns; ++si)
                for (int t = start; t < std::min(start + cnt, q < P ? (int)steps[si][q].size() : 0); ++t) {
                    const StepRec &r = steps[si][q][t];
                    so.ulist[(size_t)si * Spad + (t - start)] = (int)(std::lower_bound(uni.begin(), uni.end(), r.J) - uni.begin());
                    for (int f = 0; f < 3; ++f)
                        if (r.p[f] >= 0) so.blocks.emplace_back(r.p[f], (si * cnt + (t - start)) * SW_REC + f * 3);
                }
            out.push_back(std::move(so));
            start += cnt;
        } while (start < Smax);
    }
}

struct HostPlan {
    std::vector<SweepStage> stages;   // in execution order: CTA by CTA
    std::vector<int> cta_ptr;         // [grid + 1]
    std::vector<int> prod;            // 64 ints (32 x int2) per stage: what warp 0 needs to start the stage's copies
    int ntiles = 0;
    std::vector<int> meta;            // per stage: header (64 ints) followed by the u lists
    std::vector<long long> meta_off;  // [nstages + 1]
    std::vector<long long> base;      // per block: index (in doubles) of its slot in the plan stream
    long long stream_bytes = 0, total_steps = 0;
    int stage_smem = 128;
};

int build_host_plan(int mb, int kb, const int *rp, const int *ci, int nstrips, const int *strip_ptr, const int *strip_rows,
                    int world, const int *cuts, int grid, int ctas_per_sm, HostPlan &hp) {
    if (grid < 1 || ctas_per_sm < 1 || ctas_per_sm > 2) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: bad grid");
    if (mb < 0 || kb < 0 || !rp || world < 1 || world > 8 || (world > 1 && !cuts) || (nstrips > 0 && (!strip_ptr || !strip_rows)))
        return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: bad arguments");
    const long long nb = mb ? rp[mb] : 0;
    if (nb < 0 || (nb && !ci) || (mb && rp[0] != 0)) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: bad row pointers");
    for (int i = 0; i < mb; ++i)
        if (rp[i + 1] < rp[i]) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: row pointers not monotone");
    // strips: the caller's, or runs of 64 consecutive block rows (for a banded matrix in its natural order neighbouring
    // rows share their columns, which is all the window needs)
    std::vector<int> dsp, dsr;
    if (nstrips <= 0) {
        const int len = 64;
        nstrips = (mb + len - 1) / len;
        dsp.resize((size_t)nstrips + 1);
        dsr.resize((size_t)mb);
        for (int s = 0; s <= nstrips; ++s) dsp[s] = (int)std::min<long long>(mb, (long long)s * len);
        for (int i = 0; i < mb; ++i) dsr[i] = i;
        strip_ptr = dsp.data();
        strip_rows = dsr.data();
    }
    if (strip_ptr[0] != 0 || strip_ptr[nstrips] != mb)
        return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: strips must cover every block row once");
    {
        std::vector<unsigned char> seen((size_t)mb, 0);
        for (int s = 0; s < nstrips; ++s) {
            if (strip_ptr[s + 1] < strip_ptr[s]) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: strip_ptr not monotone");
            for (int k = strip_ptr[s]; k < strip_ptr[s + 1]; ++k) {
                const int r = strip_rows[k];
                if (r < 0 || r >= mb || seen[r]) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: strips must cover every block row once");
                seen[r] = 1;
            }
        }
    }
    for (long long p = 0; p < nb; ++p)
        if (ci[p] < 0 || ci[p] >= kb) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: block column out of range");
    // two stage buffers per CTA; with two CTAs per SM a stage holds a third of a phase of a mesh operator (3 of the 9
    // columns of each strip: one line of the halo patch per strip row), with one CTA a whole phase
    const int smem_budget = ((227 * 1024 / ctas_per_sm - 1024) / 2) & ~127;
    static const int steps_env = [] {
        const char *e = getenv("G4S_BSR_SWEEP_STEPS");
        return e && atoi(e) > 0 ? atoi(e) : 0;
    }();
    const int max_steps = steps_env ? steps_env : (ctas_per_sm == 2 ? 3 : 9);
    const int ntiles = (nstrips + SW_STRIPS - 1) / SW_STRIPS;
    std::vector<std::vector<StageOut>> per_tile((size_t)ntiles);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < ntiles; ++t)
        plan_tile(rp, ci, strip_ptr, strip_rows, t * SW_STRIPS, std::min(nstrips, (t + 1) * SW_STRIPS), world, cuts, smem_budget,
                  max_steps, per_tile[t]);
    // execution order: CTA b sweeps tiles b, b + grid, ... (neighbouring tiles run at the same time on other SMs)
    hp.ntiles = ntiles;
    std::vector<const StageOut *> order;
    hp.cta_ptr.assign((size_t)grid + 1, 0);
    for (int b = 0; b < grid; ++b) {
        for (int t = b; t < ntiles; t += grid)
            for (const StageOut &so : per_tile[t]) order.push_back(&so);
        hp.cta_ptr[b + 1] = (int)order.size();
    }
    const int nstages = (int)order.size();
    hp.stages.resize((size_t)nstages);
    hp.meta_off.assign((size_t)nstages + 1, 0);
    hp.prod.assign((size_t)nstages * 64, 0);
    long long off = 0;
    for (int q = 0; q < nstages; ++q) {
        const StageOut &so = *order[q];
        const int cb = stage_chunk_bytes(so.S);
        hp.stages[q].off = off;
        hp.stages[q].chunk_bytes = cb;
        hp.stages[q].tx_bytes = cb + so.U * SW_NODE * 8;
        hp.meta_off[q + 1] = hp.meta_off[q] + SW_HDR_INTS + (long long)so.ulist.size();
        hp.stage_smem = std::max(hp.stage_smem, round128(cb) + so.U * SW_NODE * 8);
        int *pd = hp.prod.data() + (size_t)q * 64;
        int pre = 0;
        for (int k = 0; k < so.hdr[0]; ++k) {
            const int len = so.hdr[SW_HDR_RUNS + 2 * k + 1];
            pd[2 * k] = so.hdr[SW_HDR_RUNS + 2 * k];
            pd[2 * k + 1] = len | (pre << 8);
            pre += len;
        }
        pd[60] = (int)(unsigned)(off & 0xffffffffLL);
        pd[61] = (int)(off >> 32);
        pd[62] = cb;
        pd[63] = hp.stages[q].tx_bytes;
        off += cb;
        hp.total_steps += (long long)SW_STRIPS * so.S;
    }
    hp.stream_bytes = off;
    if (hp.stage_smem > smem_budget) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: a stage does not fit shared memory");
    hp.meta.resize((size_t)hp.meta_off[nstages]);
    hp.base.assign((size_t)nb, -1);
#pragma omp parallel for schedule(static)
    for (int q = 0; q < nstages; ++q) {
        const StageOut &so = *order[q];
        int *m = hp.meta.data() + hp.meta_off[q];
        std::memcpy(m, so.hdr, sizeof(int) * SW_HDR_INTS);
        if (!so.ulist.empty()) std::memcpy(m + SW_HDR_INTS, so.ulist.data(), sizeof(int) * so.ulist.size());
        const long long rec0 = (hp.stages[q].off + SW_HDR_INTS * 4 + (long long)so.ulist.size() * 4) / 8;
        for (const auto &b : so.blocks) hp.base[b.first] = rec0 + b.second;
    }
    for (long long p = 0; p < nb; ++p)
        if (hp.base[p] < 0) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan: internal error (unplaced block)");
    return G4S_OK;
}

}  // namespace

extern "C" {

int g4s_bsr3_plan_destroy(g4s_bsr_plan_t plan) {
    if (!plan) return G4S_OK;
    if (plan->stream) cudaFree(plan->stream);
    if (plan->prod) cudaFree(plan->prod);
    if (plan->cta_ptr) cudaFree(plan->cta_ptr);
    if (plan->base) cudaFree(plan->base);
    delete plan;
    return G4S_OK;
}

int g4s_bsr3_plan_create(g4s_bsr_plan_t *out, int mb, int kb, const int *browptr_dev, const int *bcolids_dev, int nstrips,
                         const int *strip_ptr, const int *strip_rows, int world, const int *cuts) {
    if (!out || mb < 0 || !browptr_dev) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_create: bad arguments");
    *out = nullptr;
    int rc = ensure_device();
    if (rc) return rc;
    std::vector<int> rp((size_t)mb + 1, 0);
    G4S_CUDA(cudaMemcpy(rp.data(), browptr_dev, sizeof(int) * ((size_t)mb + 1), cudaMemcpyDeviceToHost));
    const long long nb = mb ? rp[mb] : 0;
    if (nb < 0 || (nb && !bcolids_dev)) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_create: bad row pointers");
    std::vector<int> ci((size_t)nb);
    if (nb) G4S_CUDA(cudaMemcpy(ci.data(), bcolids_dev, sizeof(int) * (size_t)nb, cudaMemcpyDeviceToHost));
    HostPlan hp;
    static const int ctas_env = [] {
        const char *e = getenv("G4S_BSR_SWEEP_CTAS");
        return e ? atoi(e) : 1;
    }();
    const int ctas_per_sm = ctas_env == 2 ? 2 : 1, grid = ctas_per_sm * sm_count();
    if ((rc = build_host_plan(mb, kb, rp.data(), ci.data(), nstrips, strip_ptr, strip_rows, world, cuts, grid, ctas_per_sm, hp)))
        return rc;
    const int nstages = (int)hp.stages.size(), ntiles = hp.ntiles;
    g4s_bsr_plan *pl = new g4s_bsr_plan();
    pl->mb = mb;
    pl->kb = kb;
    pl->world = world;
    pl->nb = nb;
    if (cuts) pl->cuts.assign(cuts, cuts + world + 1);
    pl->stream_bytes = (size_t)hp.stream_bytes;
    pl->nstages = nstages;
    pl->ntiles = ntiles;
    pl->grid = grid;
    pl->stage_smem = hp.stage_smem;
    pl->total_steps = hp.total_steps;
    auto bail = [&](const char *what) {
        g4s_bsr3_plan_destroy(pl);
        return fail(G4S_ERR_CUDA, std::string("g4s_bsr3_plan_create: ") + what);
    };
    int *meta_dev = nullptr;
    long long *meta_off_dev = nullptr;
    SweepStage *stages_dev = nullptr;
    if (cudaMalloc(&pl->stream, std::max<size_t>(pl->stream_bytes, 16)) != cudaSuccess) return bail("stream allocation");
    if (cudaMalloc(&pl->prod, sizeof(int2) * 32 * (size_t)std::max(nstages, 1)) != cudaSuccess) return bail("descriptor table");
    if (cudaMalloc(&pl->cta_ptr, sizeof(int) * ((size_t)grid + 1)) != cudaSuccess) return bail("CTA table");
    if (cudaMalloc(&stages_dev, sizeof(SweepStage) * std::max(nstages, 1)) != cudaSuccess) return bail("stage table");
    if (cudaMalloc(&pl->base, sizeof(long long) * std::max<long long>(nb, 1)) != cudaSuccess) return bail("block map");
    if (cudaMalloc(&meta_dev, sizeof(int) * std::max<size_t>(hp.meta.size(), 1)) != cudaSuccess) return bail("meta");
    if (cudaMalloc(&meta_off_dev, sizeof(long long) * ((size_t)nstages + 1)) != cudaSuccess) {
        cudaFree(meta_dev);
        cudaFree(stages_dev);
        return bail("meta offsets");
    }
    cudaMemset(pl->stream, 0, std::max<size_t>(pl->stream_bytes, 16));
    if (nstages) cudaMemcpy(stages_dev, hp.stages.data(), sizeof(SweepStage) * nstages, cudaMemcpyHostToDevice);
    if (nstages) cudaMemcpy(pl->prod, hp.prod.data(), sizeof(int) * hp.prod.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->cta_ptr, hp.cta_ptr.data(), sizeof(int) * ((size_t)grid + 1), cudaMemcpyHostToDevice);
    if (nb) cudaMemcpy(pl->base, hp.base.data(), sizeof(long long) * (size_t)nb, cudaMemcpyHostToDevice);
    if (!hp.meta.empty()) cudaMemcpy(meta_dev, hp.meta.data(), sizeof(int) * hp.meta.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(meta_off_dev, hp.meta_off.data(), sizeof(long long) * ((size_t)nstages + 1), cudaMemcpyHostToDevice);
    if (nstages) {
        sweep_scatter_meta_kernel<<<nstages, 128>>>(meta_dev, meta_off_dev, stages_dev, nstages, pl->stream);
        count_launch();
    }
    const cudaError_t e = cudaDeviceSynchronize();
    cudaFree(meta_dev);
    cudaFree(meta_off_dev);
    cudaFree(stages_dev);
    if (e != cudaSuccess) return bail(cudaGetErrorString(e));
    *out = pl;
    return G4S_OK;
}

// The schedule alone, from HOST arrays, without touching a device: what the CPU tests replay step by step
// (tests/test_bsr_plan_cpu.py).  Arrays are malloc'd (g4s_free).
int g4s_bsr3_plan_inspect_host(int mb, int kb, const int *browptr, const int *bcolids, int nstrips, const int *strip_ptr,
                               const int *strip_rows, int world, const int *cuts, int grid, int ctas_per_sm, int *nstages,
                               int *ntiles, long long *stream_bytes, int *stage_smem_bytes, double *slot_fill,
                               long long **stage_table, int **cta_ptr, int **prod, int **meta, long long **meta_off,
                               long long **base) {
    if (!browptr || mb < 0) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_inspect_host: bad arguments");
    HostPlan hp;
    int rc = build_host_plan(mb, kb, browptr, bcolids, nstrips, strip_ptr, strip_rows, world, cuts, grid, ctas_per_sm, hp);
    if (rc) return rc;
    const long long nb = mb ? browptr[mb] : 0;
    if (nstages) *nstages = (int)hp.stages.size();
    if (ntiles) *ntiles = hp.ntiles;
    if (stream_bytes) *stream_bytes = hp.stream_bytes;
    if (stage_smem_bytes) *stage_smem_bytes = hp.stage_smem;
    if (slot_fill) *slot_fill = hp.total_steps ? (double)nb / (3.0 * (double)hp.total_steps) : 1.0;
    auto dup = [](const void *src, size_t bytes) {
        void *p = malloc(std::max<size_t>(bytes, 8));
        if (p && bytes) std::memcpy(p, src, bytes);
        return p;
    };
    if (stage_table) {  // (offset, chunk bytes | tx bytes << 32) per stage
        static_assert(sizeof(SweepStage) == 16, "stage table entries are two 64-bit words");
        *stage_table = (long long *)dup(hp.stages.data(), sizeof(SweepStage) * hp.stages.size());
    }
    if (cta_ptr) *cta_ptr = (int *)dup(hp.cta_ptr.data(), sizeof(int) * hp.cta_ptr.size());
    if (prod) *prod = (int *)dup(hp.prod.data(), sizeof(int) * hp.prod.size());
    if (meta) *meta = (int *)dup(hp.meta.data(), sizeof(int) * hp.meta.size());
    if (meta_off) *meta_off = (long long *)dup(hp.meta_off.data(), sizeof(long long) * hp.meta_off.size());
    if (base) *base = (long long *)dup(hp.base.data(), sizeof(long long) * hp.base.size());
    return G4S_OK;
}

int g4s_bsr3_plan_set_values(g4s_bsr_plan_t plan, const double *bvalues_dev, void *stream) {
    if (!plan || (plan->nb && !bvalues_dev)) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_set_values: null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (plan->nb) {
        const int grid = (int)std::min<long long>((plan->nb + 255) / 256, (long long)sm_count() * 16);
        sweep_set_values_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bvalues_dev, plan->base, plan->nb,
                                                                        reinterpret_cast<double *>(plan->stream));
        G4S_CHECK_LAUNCH("sweep_set_values_kernel");
    }
    plan->values_set = true;
    return G4S_OK;
}

int g4s_bsr3_plan_info(g4s_bsr_plan_t plan, double *slot_fill, long long *stream_bytes, int *nstages, int *ntiles,
                       int *stage_smem_bytes) {
    if (!plan) return fail(G4S_ERR_INVALID, "null plan");
    if (slot_fill) *slot_fill = plan->total_steps ? (double)plan->nb / (3.0 * (double)plan->total_steps) : 1.0;
    if (stream_bytes) *stream_bytes = (long long)plan->stream_bytes;
    if (nstages) *nstages = plan->nstages;
    if (ntiles) *ntiles = plan->ntiles;
    if (stage_smem_bytes) *stage_smem_bytes = plan->stage_smem;
    return G4S_OK;
}

static int sweep_launch(g4s_bsr_plan_t plan, const double *B_dev, const BParts &bp, bool parts, double *C_dev, cudaStream_t st) {
    if (!plan->values_set) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_spmm64: call g4s_bsr3_plan_set_values first");
    if (plan->mb == 0 || plan->ntiles == 0) return G4S_OK;
    int rc = ensure_device();
    if (rc) return rc;
    SweepArgs a;
    a.stream = plan->stream;
    a.prod = plan->prod;
    a.cta_ptr = plan->cta_ptr;
    a.stage_smem = plan->stage_smem;
    const int smem = 2 * plan->stage_smem;
    auto k0 = bsr3_sweep_kernel<false>;
    auto k1 = bsr3_sweep_kernel<true>;
    static PerDeviceOnce configured;
    if (configured.needs(smem)) {
        G4S_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        G4S_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.done(smem);
    }
    const int grid = plan->grid;  // fixed by the plan: every CTA's stages are stored consecutively
    if (parts) k1<<<grid, (SW_WARPS + 1) * 32, smem, st>>>(a, nullptr, bp, C_dev);
    else k0<<<grid, (SW_WARPS + 1) * 32, smem, st>>>(a, B_dev, bp, C_dev);
    G4S_CHECK_LAUNCH("bsr3_sweep_kernel");
    return G4S_OK;
}

int g4s_bsr3_plan_spmm64_device(g4s_bsr_plan_t plan, const double *B_dev, double *C_dev, void *stream) {
    if (!plan || !B_dev || !C_dev || ((reinterpret_cast<uintptr_t>(B_dev) | reinterpret_cast<uintptr_t>(C_dev)) & 15))
        return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_spmm64_device: bad arguments (B and C 16-byte aligned)");
    if (plan->world != 1) return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_spmm64_device: the plan was made for a partitioned B");
    BParts none = {};
    return sweep_launch(plan, B_dev, none, false, C_dev, (cudaStream_t)stream);
}

int g4s_bsr3_plan_spmm64_partitioned_device(g4s_bsr_plan_t plan, int world, const double *const *B_parts, const int *cuts,
                                            double *C_dev, void *stream) {
    if (!plan || !B_parts || !cuts || !C_dev || world != plan->world)
        return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_spmm64_partitioned_device: bad arguments (world must match the plan)");
    for (int q = 0; q <= world; ++q)
        if (plan->cuts.empty() || cuts[q] != plan->cuts[q])
            return fail(G4S_ERR_INVALID, "g4s_bsr3_plan_spmm64_partitioned_device: cuts differ from the plan's");
    BParts bp;
    for (int q = 0; q < 8; ++q) bp.base[q] = q < world ? B_parts[q] : nullptr;
    for (int q = 0; q <= 8; ++q) bp.cut[q] = cuts[q < world ? q : world];
    bp.world = world;
    return sweep_launch(plan, nullptr, bp, true, C_dev, (cudaStream_t)stream);
}

int g4s_grid_pencil_strips(int n0, int n1, int k_begin, int k_end, int p0, int p1, int *strip_ptr, int *strip_rows) {
    if (n0 < 1 || n1 < 1 || k_begin < 0 || k_end < k_begin || p0 < 1 || p1 < 1 || !strip_ptr || !strip_rows ||
        (long long)n0 * n1 * (k_end - k_begin) > 2147483647LL)
        return fail(G4S_ERR_INVALID, "g4s_grid_pencil_strips: bad arguments");
    long long at = 0;
    int s = 0;
    strip_ptr[0] = 0;
    for (int b1 = 0; b1 < n1; b1 += p1)
        for (int b0 = 0; b0 < n0; b0 += p0)
            for (int j = b1; j < std::min(n1, b1 + p1); ++j)
                for (int i = b0; i < std::min(n0, b0 + p0); ++i) {
                    for (int k = k_begin; k < k_end; ++k) strip_rows[at++] = (int)(((long long)(k - k_begin) * n1 + j) * n0 + i);
                    strip_ptr[++s] = (int)at;
                }
    return G4S_OK;
}

}  // extern "C"
