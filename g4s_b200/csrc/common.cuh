// Shared device/host helpers for the g4s_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/g4s_b200.h"

namespace g4s {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int status, const std::string &msg);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define G4S_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::g4s::fail(G4S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

#define G4S_CHECK_LAUNCH(name)                                                                           \
    do {                                                                                                 \
        cudaError_t _e = cudaGetLastError();                                                             \
        if (_e != cudaSuccess)                                                                           \
            return ::g4s::fail(G4S_ERR_CUDA, std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
        ::g4s::count_launch();                                                                           \
    } while (0)

int ensure_device();  // G4S_OK when the current device is usable (sm_100), else G4S_ERR_CUDA
// stable LSD radix sort of (key, 8-byte value) pairs by the low end_bit bits; result in (k1, v1) (radix_sort.cu)
int radix_sort_pairs_u64(unsigned long long *k0, unsigned long long *k1, unsigned long long *v0, unsigned long long *v1,
                         long long n, int end_bit, cudaStream_t stream);
int sm_count();

// cudaFuncSetAttribute state is per device: a once-per-process flag would leave every kernel that needs > 48 KB of dynamic
// shared memory unconfigured on the second device a process uses.  One slot per device, holding the value the attribute
// was last set for (key) — a launch with another key (e.g. another shared-memory carve-out for the same instantiation)
// configures again.  Thread-safe: concurrent first launches at worst both set the same attributes.
struct PerDeviceOnce {
    std::atomic<int> key[64];
    PerDeviceOnce() {
        for (auto &k : key) k.store(-1, std::memory_order_relaxed);
    }
    // true when the caller must (re)configure for `want` on the current device
    bool needs(int want = 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        return key[dev].load(std::memory_order_acquire) != want;
    }
    void done(int want = 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) key[dev].store(want, std::memory_order_release);
    }
};

// ---- the CSR handle ---------------------------------------------------------------------------------
struct SpmvPlan {
    int cap = 0;                        // max nonzeros per chunk
    int nchunks = 0;
    int2 *desc = nullptr;               // [nchunks+1] (first row, first nnz) of each chunk
    unsigned char *lanes_lg = nullptr;  // [nchunks]   log2(lanes per row) chosen by the inspector
    double *carry = nullptr;            // [nchunks]   partial sums of non-final pieces of long rows
    int4 *long_rows = nullptr;          // [n_long]    (row, first chunk, pieces, -)
    unsigned char *part_flags = nullptr;  // [nchunks] partitioned x: chunk references a remote column
    int *part_colids = nullptr;           // [nnz] private copy of the column ids with remote columns rewritten to -1 - k
    int *part_needed = nullptr;           // localized plan: global column of halo entry k (ascending)
    int part_n_needed = 0;
    double *part_halo = nullptr;          // [part_n_needed] remote x entries, refilled by every partitioned product
    unsigned long long *part_halo_done = nullptr;  // gather warps that have published their share, over all products
    unsigned long long part_products = 0;
    unsigned part_owner_mask = 0xffu;     // ranks that own this rank's remote columns
    int part_lo = 0, part_hi = 0;       // owned column range the flags were computed for
    int n_long = 0;                     // rows longer than cap
    int max_row_len = 0;
    int lanes_per_row = 0;              // tuning override, 0 = inspector's choice
    int variant = 0;                    // tuning override, 0 = default kernel shape
    int built_for_variant = -1;         // the override the plan was cut for
    int shape_variant = 6;              // kernel shape the plan was cut for (the automatic choice resolved)
    int rows_per_chunk = 32;            // 32, or 64 / 128 for regular short-row matrices
};

struct HostPipe;  // spmv.cu: streams / events / row blocks of the pipelined host-pointer product

}  // namespace g4s

struct g4s_csr {
    int rows = 0, cols = 0;
    long long nnz = 0;
    int *rowptr = nullptr;
    int *colids = nullptr;
    double *values = nullptr;
    bool owns = false;
    int sorted_cols = -1;  // -1 unknown, 1 every row's column ids strictly ascending, 0 not (SpGEMM merge class)
    bool pooled = false;  // arrays came from cudaMallocAsync (stream-ordered pool) rather than cudaMalloc
    void *pool_base = nullptr;  // pooled: the three arrays are slices of this one allocation (a repeated SpGEMM's result)
    g4s::SpmvPlan plan;
    // row-compressed blocks (off-diagonal part of a multi-GPU row block): local row r is row row_map[r] of a
    // block with full_rows rows
    int *row_map = nullptr;
    int full_rows = 0;
    // scratch for the host-pointer entry points
    double *x_dev = nullptr, *y_dev = nullptr;
    g4s::HostPipe *host_pipe = nullptr;
};

namespace g4s {

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// L2 eviction policy for data that is streamed exactly once (matrix arrays): keep x / y resident instead.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16, completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ double ldg_f64(const double *p) { return __ldg(p); }

// ---- dense operand of the multi-GPU BSR SpMM: row-partitioned, every part in CUDA-IPC-shared memory ------------
struct BParts {
    const double *base[8];
    int cut[9];
    int world;
};
__device__ __forceinline__ const double *b_part_row(const BParts &bp, int J) {
    const double *b = bp.base[0];
    int cut = bp.cut[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const bool ge = i < bp.world && J >= bp.cut[i];
        b = ge ? bp.base[i] : b;
        cut = ge ? bp.cut[i] : cut;
    }
    return b + (size_t)(J - cut) * 3 * 64;
}
// the same lookup for a block column that differs from lane to lane (K-packed kernels): count the cuts at or below J and
// index the parameter block with the result (two constant-bank loads) instead of carrying a pointer through seven selects
__device__ __forceinline__ const double *b_part_row_lane(const BParts &bp, int J) {
    int owner = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) owner += (i < bp.world && J >= bp.cut[i]) ? 1 : 0;
    return bp.base[owner] + (size_t)(J - bp.cut[owner]) * 3 * 64;
}

}  // namespace g4s
