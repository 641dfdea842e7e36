// OptMatmul: res[M,K] = xx[M,N] w[N,K], fp64, row-major — the dense product DeePMD-kit routes through the
// G4S engine (deepmd/source/op/opt_matmul.cc:24-62: vertices = the M rows, degree = K, gather = one dot product of
// length N; engine loop deepmd/source/op/graph.h:21-32), and the two products of its registered gradient
// (deepmd/source/op/_opt_matmul_grad.py: dxx = grad w^T, dw = xx^T grad).
//
// One kernel for all three: C[m,n] = sum_k A(m,k) B(k,n) with A and B addressed through (row, column) strides, so a
// transposed operand is a stride swap.  FP64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no FP64
// kind).  On this part DMMA retires FMAs at the DFMA rate, but a fragment is ONE double per lane for 256 FMAs, so the
// operand traffic into registers (what limits DFMA kernels, see bsr_sweep.cu) is an eighth of the scalar form.
// CTA = 8 warps, tile 128 x 64 x 16: warp tile 32 x 32 = 4 x 4 DMMA tiles (32 accumulator registers); operands staged
// in shared memory through registers (plain loads: N and K of the embedding / fitting nets are arbitrary, often odd, so
// 16-byte copies cannot be assumed), double-buffered; row strides of 20 and 68 doubles make both fragment loads
// conflict-free per half-warp.  Sums run over k in ascending groups of four: rounding-level differences from the
// reference's sequential loop only.
#include <algorithm>

#include "common.cuh"

namespace g4s {

constexpr int GM = 128, GN = 64, GK = 16, GA_LD = GK + 4, GB_LD = GN + 4;

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

struct GemmArgs {
    const double *A, *B;
    double *C;
    int M, N, K;                    // C is M x N, the sum runs over K
    long long a_rs, a_cs, b_rs, b_cs;  // element (r, c) of an operand sits at r * rs + c * cs
    long long ldc;
};

__global__ void __launch_bounds__(256) dgemm_dmma_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) double gemm_sm[];
    double(*As)[GM * GA_LD] = reinterpret_cast<double(*)[GM * GA_LD]>(gemm_sm);
    double(*Bs)[GK * GB_LD] = reinterpret_cast<double(*)[GK * GB_LD]>(gemm_sm + 2 * GM * GA_LD);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    // loader roles: A tile 128 x 16 = 2048 doubles, 8 per thread; B tile 16 x 64 = 1024 doubles, 4 per thread.  The faster
    // index of a thread follows the operand's contiguous direction (row-major xx / w: columns; a transposed view: rows).
    const bool a_kfast = g.a_cs <= g.a_rs, b_nfast = g.b_cs <= g.b_rs;
    double ra[8], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = t + u * 256;
            const int m = a_kfast ? idx >> 4 : idx & 127, k = a_kfast ? idx & 15 : idx >> 7;
            const int gm = m0 + m, gk = k0 + k;
            ra[u] = (gm < g.M && gk < g.K) ? __ldg(g.A + gm * g.a_rs + gk * g.a_cs) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = t + u * 256;
            const int k = b_nfast ? idx >> 6 : idx & 15, n = b_nfast ? idx & 63 : idx >> 4;
            const int gk = k0 + k, gn = n0 + n;
            rb[u] = (gk < g.K && gn < g.N) ? __ldg(g.B + gk * g.b_rs + gn * g.b_cs) : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = t + u * 256;
            const int m = a_kfast ? idx >> 4 : idx & 127, k = a_kfast ? idx & 15 : idx >> 7;
            As[buf][m * GA_LD + k] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = t + u * 256;
            const int k = b_nfast ? idx >> 6 : idx & 15, n = b_nfast ? idx & 63 : idx >> 4;
            Bs[buf][k * GB_LD + n] = rb[u];
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nk = (g.K + GK - 1) / GK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * GK);  // in flight while this tile is multiplied
        const double *as = As[buf] + (wm + (lane >> 2)) * GA_LD + (lane & 3);
        const double *bs = Bs[buf] + (lane & 3) * GB_LD + wn + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < GK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = as[i * 8 * GA_LD + k4 * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[k4 * 4 * GB_LD + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (kt + 1 < nk) {
            stash(buf ^ 1);  // the other buffer: its last readers passed the barrier below one iteration ago
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm + i * 8 + (lane >> 2);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn + j * 8 + 2 * (lane & 3);
            double *c = g.C + (long long)m * g.ldc + n;
            if (n < g.N) c[0] = acc[i][j][0];
            if (n + 1 < g.N) c[1] = acc[i][j][1];
        }
    }
}

static int dgemm_launch(const double *A, long long a_rs, long long a_cs, const double *B, long long b_rs, long long b_cs,
                        double *C, long long ldc, int M, int N, int K, cudaStream_t stream) {
    if (M == 0 || N == 0) return G4S_OK;
    GemmArgs g;
    g.A = A;
    g.B = B;
    g.C = C;
    g.M = M;
    g.N = N;
    g.K = K;
    g.a_rs = a_rs;
    g.a_cs = a_cs;
    g.b_rs = b_rs;
    g.b_cs = b_cs;
    g.ldc = ldc;
    const dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN);
    if (grid.y > 65535) return fail(G4S_ERR_INVALID, "g4s_opt_matmul: output too wide");
    constexpr int smem = sizeof(double) * 2 * (GM * GA_LD + GK * GB_LD);  // 58 KB: above the 48 KB static limit
    static PerDeviceOnce configured;
    if (configured.needs()) {
        G4S_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.done();
    }
    dgemm_dmma_kernel<<<grid, 256, smem, stream>>>(g);
    G4S_CHECK_LAUNCH("dgemm_dmma_kernel");
    return G4S_OK;
}

}  // namespace g4s

using namespace g4s;

extern "C" {

int g4s_opt_matmul_device(int M, int N, int K, const double *xx_dev, const double *w_dev, double *res_dev, void *stream) {
    if (M < 0 || N < 0 || K < 0 || (M && N && !xx_dev) || (N && K && !w_dev) || (M && K && !res_dev))
        return fail(G4S_ERR_INVALID, "g4s_opt_matmul_device: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    // res[M,K] = xx[M,N] w[N,K]: the sum runs over N
    return dgemm_launch(xx_dev, N, 1, w_dev, K, 1, res_dev, K, M, K, N, (cudaStream_t)stream);
}

int g4s_opt_matmul_grad_device(int M, int N, int K, const double *xx_dev, const double *w_dev, const double *grad_dev,
                               double *dxx_dev, double *dw_dev, void *stream) {
    if (M < 0 || N < 0 || K < 0 || !grad_dev) return fail(G4S_ERR_INVALID, "g4s_opt_matmul_grad_device: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    // dxx[M,N] = grad[M,K] w^T : B(k,n) = w[n*K + k]
    if (dxx_dev) {
        if (!w_dev) return fail(G4S_ERR_INVALID, "g4s_opt_matmul_grad_device: dxx needs w");
        if ((rc = dgemm_launch(grad_dev, K, 1, w_dev, 1, K, dxx_dev, N, M, N, K, (cudaStream_t)stream))) return rc;
    }
    // dw[N,K] = xx^T grad : A(n,m) = xx[m*N + n]
    if (dw_dev) {
        if (!xx_dev) return fail(G4S_ERR_INVALID, "g4s_opt_matmul_grad_device: dw needs xx");
        if ((rc = dgemm_launch(xx_dev, 1, N, grad_dev, K, 1, dw_dev, K, N, K, M, (cudaStream_t)stream))) return rc;
    }
    return G4S_OK;
}

int g4s_opt_matmul(int M, int N, int K, const double *xx, const double *w, double *res) {
    if (M < 0 || N < 0 || K < 0 || (M && N && !xx) || (N && K && !w) || (M && K && !res))
        return fail(G4S_ERR_INVALID, "g4s_opt_matmul: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    if (M == 0 || K == 0) return G4S_OK;
    double *dx = nullptr, *dw = nullptr, *dr = nullptr;
    auto done = [&](int code) {
        cudaFree(dx);
        cudaFree(dw);
        cudaFree(dr);
        return code;
    };
    if (cudaMalloc(&dx, sizeof(double) * std::max<size_t>((size_t)M * N, 1)) != cudaSuccess ||
        cudaMalloc(&dw, sizeof(double) * std::max<size_t>((size_t)N * K, 1)) != cudaSuccess ||
        cudaMalloc(&dr, sizeof(double) * (size_t)M * K) != cudaSuccess)
        return done(fail(G4S_ERR_ALLOC, "g4s_opt_matmul: device allocation failed"));
    if ((size_t)M * N && cudaMemcpy(dx, xx, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice) != cudaSuccess)
        return done(fail(G4S_ERR_CUDA, "g4s_opt_matmul: upload failed"));
    if ((size_t)N * K && cudaMemcpy(dw, w, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice) != cudaSuccess)
        return done(fail(G4S_ERR_CUDA, "g4s_opt_matmul: upload failed"));
    if ((rc = g4s_opt_matmul_device(M, N, K, dx, dw, dr, nullptr))) return done(rc);
    if (cudaMemcpy(res, dr, sizeof(double) * (size_t)M * K, cudaMemcpyDeviceToHost) != cudaSuccess)
        return done(fail(G4S_ERR_CUDA, "g4s_opt_matmul: download failed"));
    return done(G4S_OK);
}

}  // extern "C"
