// DFMA latency / throughput on one SM as a function of warps and independent chains per thread:
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_probe dfma_probe.cu && ./dfma_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, int iters, double x, double y) {
    double a[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = (double)(threadIdx.x + c) * 1e-6;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) a[c] = fma(a[c], x, y);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
void run(int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 256;
    k<CH><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
    k<CH><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_dfma_warp = (double)h / (iters * 8.0 * CH);
    printf("warps %2d chains %d: %8lld cycles, %.2f cycles per DFMA per warp, %.1f lanes/clk/SM\n", warps, CH, h, per_dfma_warp,
           warps * 32.0 / per_dfma_warp);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 13, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
