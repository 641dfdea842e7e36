// Stable LSD radix sort of (64-bit key, 8-byte value) pairs on the device: the library's own sort, replacing
// cub::DeviceRadixSort in the expand-sort-compress SpGEMM (spgemm_esc.cu; the reference sorts its (row, col, value)
// tuples with radix_sort(begin, end, buf, key), mm/inc/radix_sort.h:701-705, called at mm/inc/outer_mult.h:427-442) and in
// the device graph-to-CSR loader (rmat.cu; the reference uses std::sort per row, mm/inc/CSR.h:273-301).
//
// 8 bits per pass, three launches per pass:
//   histogram : every CTA counts the digits of its tile of 4096 keys into hist[digit][cta]
//   scan      : exclusive prefix sum over hist in (digit, cta) order = where each CTA's keys of each digit go
//   scatter   : every warp owns a contiguous 512-key piece of the tile and walks it 32 keys at a time; lanes with equal
//               digits find each other with match.any, rank themselves by lane order and advance the warp's running
//               offset of that digit — keys of equal digit keep their input order (stable), which the LSD scheme and
//               the callers (values of equal (row, col) keys are summed left to right) rely on.
#include <algorithm>

#include "common.cuh"

namespace g4s {

int exclusive_scan_i32(const int *in, int *out, long long n, int write_total, long long *total_host, cudaStream_t stream);

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_TILE = 4096, RS_PIECE = RS_TILE / RS_WARPS;

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const unsigned long long *__restrict__ keys, long long n,
                                                                int shift, unsigned mask, int *__restrict__ hist, int nblocks) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const long long k = base + i;
        if (k < n) atomicAdd(&h[(unsigned)(__ldg(keys + k) >> shift) & mask], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
    radix_scatter_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vals, long long n,
                         int shift, unsigned mask, const int *__restrict__ offs, int nblocks,
                         unsigned long long *__restrict__ keys_out, unsigned long long *__restrict__ vals_out) {
    __shared__ int wh[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = lane; d < 256; d += 32) wh[warp][d] = 0;
    __syncwarp();
    const long long piece = (long long)blockIdx.x * RS_TILE + (long long)warp * RS_PIECE;
    const unsigned lt = (1u << lane) - 1;
    // pass A: digits of this warp's piece
    for (int r = 0; r < RS_PIECE; r += 32) {
        const long long k = piece + r + lane;
        const bool on = k < n;
        const unsigned act = __ballot_sync(0xffffffffu, on);
        if (on) {
            const unsigned d = (unsigned)(__ldg(keys + k) >> shift) & mask;
            const unsigned peers = __match_any_sync(act, d);
            if ((peers & lt) == 0) wh[warp][d] += __popc(peers);  // the lowest lane of every digit group
        }
        __syncwarp();
    }
    __syncthreads();
    {  // thread d: where the CTA's keys of digit d start, then warp by warp
        const int d = threadIdx.x;
        int run = __ldg(offs + (size_t)d * nblocks + blockIdx.x);
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const int c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // pass B: the same walk, now writing
    for (int r = 0; r < RS_PIECE; r += 32) {
        const long long k = piece + r + lane;
        const bool on = k < n;
        const unsigned act = __ballot_sync(0xffffffffu, on);
        if (on) {
            const unsigned long long key = __ldg(keys + k);
            const unsigned d = (unsigned)(key >> shift) & mask;
            const unsigned peers = __match_any_sync(act, d);
            const int pos = wh[warp][d] + __popc(peers & lt);
            keys_out[pos] = key;
            vals_out[pos] = __ldg(vals + k);
            __syncwarp(act);
            if ((peers & lt) == 0) wh[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
}

// Sorts n pairs by the low `end_bit` bits of the key; the result is in (k1, v1), (k0, v0) are scratch afterwards.
int radix_sort_pairs_u64(unsigned long long *k0, unsigned long long *k1, unsigned long long *v0, unsigned long long *v1,
                         long long n, int end_bit, cudaStream_t stream) {
    if (n <= 0) return G4S_OK;
    if (n > 2147483647LL) return fail(G4S_ERR_INVALID, "radix sort: more than 2^31-1 pairs");
    end_bit = std::max(1, std::min(64, end_bit));
    const int passes = (end_bit + 7) / 8;
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    int *hist = nullptr;
    G4S_CUDA(cudaMallocAsync(&hist, sizeof(int) * 256 * (size_t)nblocks, stream));
    unsigned long long *ks = k0, *kd = k1, *vs = v0, *vd = v1;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p, bits = std::min(8, end_bit - shift);
        const unsigned mask = (1u << bits) - 1;
        radix_hist_kernel<<<nblocks, RS_THREADS, 0, stream>>>(ks, n, shift, mask, hist, nblocks);
        G4S_CHECK_LAUNCH("radix_hist_kernel");
        int rc = exclusive_scan_i32(hist, hist, 256LL * nblocks, 0, nullptr, stream);
        if (rc) {
            cudaFreeAsync(hist, stream);
            return rc;
        }
        radix_scatter_kernel<<<nblocks, RS_THREADS, 0, stream>>>(ks, vs, n, shift, mask, hist, nblocks, kd, vd);
        G4S_CHECK_LAUNCH("radix_scatter_kernel");
        std::swap(ks, kd);
        std::swap(vs, vd);
    }
    if (ks != k1) {  // an even number of passes ends in the first pair of buffers
        G4S_CUDA(cudaMemcpyAsync(k1, ks, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
        G4S_CUDA(cudaMemcpyAsync(v1, vs, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    }
    G4S_CUDA(cudaFreeAsync(hist, stream));
    return G4S_OK;
}

}  // namespace g4s

extern "C" int g4s_radix_sort_pairs_device(unsigned long long *keys_dev, unsigned long long *keys_tmp_dev, void *values_dev,
                                           void *values_tmp_dev, long long n, int key_bits, void *stream) {
    if (n < 0 || (n && (!keys_dev || !keys_tmp_dev || !values_dev || !values_tmp_dev)))
        return g4s::fail(G4S_ERR_INVALID, "g4s_radix_sort_pairs_device: bad arguments");
    if (n == 0) return G4S_OK;
    int rc = g4s::ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = g4s::radix_sort_pairs_u64(keys_dev, keys_tmp_dev, (unsigned long long *)values_dev, (unsigned long long *)values_tmp_dev, n,
                                   key_bits, st);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(keys_dev, keys_tmp_dev, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(values_dev, values_tmp_dev, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return g4s::fail(G4S_ERR_CUDA, cudaGetErrorString(e));
    return G4S_OK;
}
