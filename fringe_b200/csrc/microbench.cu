// FP32 FMA peak microbenchmark: the measured denominator of the covariance + eigen roofline.
// Two register-resident variants are timed (scalar FFMA chains and packed fma.rn.f32x2, the
// sm_100 two-wide FP32 FMA); the better one is reported.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace fringe {

template <int CHAINS>
__global__ void __launch_bounds__(256) k_fma_scalar(float* out, int iters, float x, float y) {
    float a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = (float)(threadIdx.x + c) * 1e-6f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) a[c] = fmaf(a[c], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_fma_packed(float* out, int iters, float x, float y) {
    unsigned long long a[CHAINS];
    unsigned long long xx, yy;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        const float v = (float)(threadIdx.x + c) * 1e-6f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a[c]) : "f"(v));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[c]) : "l"(xx), "l"(yy));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[c]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fp32_peak(cudaStream_t st, double* tflops) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = nsm * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 4; ++rep) {          // rep 0 = warm-up
            cudaEventRecord(e0, st);
            double flops;
            if (variant == 0) {
                k_fma_scalar<16><<<blocks, threads, 0, st>>>(out, iters, 0.999f, 1e-3f);
                flops = 2.0 * blocks * threads * (double)iters * 8 * 16;
            } else {
                k_fma_packed<8><<<blocks, threads, 0, st>>>(out, iters, 0.999f, 1e-3f);
                flops = 2.0 * blocks * threads * (double)iters * 8 * 8 * 2;
            }
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double tf = (ms > 0.f) ? flops / (ms * 1e-3) * 1e-12 : 0.0;
            if (rep > 0 && getenv("FRINGE_PEAK_VERBOSE")) printf("fp32 peak variant %d rep %d: %.2f TFLOP/s\n", variant, rep, tf);
            if (rep > 0 && tf > best) best = tf;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return cudaGetLastError();
}

}  // namespace fringe
