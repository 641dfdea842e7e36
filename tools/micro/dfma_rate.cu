// Micro-benchmark: how fast can ONE warp per scheduler issue independent DFMAs (72 accumulators, 9 x 8 outer product per
// sub-step, operands in registers), against two warps per scheduler?  Prints cycles per warp-DFMA per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dfma_rate tools/micro/dfma_rate.cu  (measured on B200: 2.639 cycles per
// warp-DFMA per scheduler with one warp per scheduler, 2.458 with two; 2.0 would be the nominal FP64 rate)
#include <cstdio>
#include <cuda_runtime.h>
template <int ITER>
__global__ void __launch_bounds__(256, 1) k(double *out, const double *in, long long *cyc, int reps) {
    double acc[9][8], a[9], b[8];
    for (int i = 0; i < 9; ++i) a[i] = in[threadIdx.x + i];
    for (int j = 0; j < 8; ++j) b[j] = in[threadIdx.x + 9 + j];
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < ITER; ++u)
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        // perturb operands a little so the loop is not hoisted
        a[r % 9] += 1e-30;
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out, *in;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 256 * 8);
    cudaMalloc(&in, 4096 * 8);
    cudaMemset(in, 0, 4096 * 8);
    cudaMalloc(&cyc, 8);
    const int reps = 2000;
    for (int warps = 4; warps <= 8; warps += 4) {
        k<3><<<148, warps * 32>>>(out, in, cyc, reps);
        k<3><<<148, warps * 32>>>(out, in, cyc, reps);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double dfma_per_sched = (double)reps * 3 * 72 * (warps / 4);
        printf("warps/CTA %d: %lld cycles, %.3f cycles per warp-DFMA per scheduler (2.0 = FP64 pipe peak)\n", warps, h,
               h / dfma_per_sched);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
