// Datum adjustment of the sequential estimator: the wrapped time series is
//     adjusted(date) = ministack phasor(date) * datum phasor(ministack of that date)
// python/adjustMiniStacks.py:180-199 leaves that product to GDAL's "mul" VRT pixel function (complex
// sources are multiplied in double and written back as CFloat32); here it is one streaming kernel.
// HBM-bound: 16 bytes in, 8 bytes out per pixel.
#include "common.cuh"

namespace fringe {

// (a * b) with both products of each component exact in double (24 x 24 bit mantissas), one
// rounding to double for the sum / difference, one to float: what a double-precision complex
// multiply followed by a cast gives, whichever order or contraction the host compiler chose.
__device__ __forceinline__ float2 cmul_via_double(float2 a, float2 b) {
    const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
    const double re = __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi));
    const double im = __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br));
    return make_float2((float)re, (float)im);
}

__global__ void __launch_bounds__(256) k_cmul(const float4* __restrict__ a, const float4* __restrict__ b,
                                              float4* __restrict__ out, long npairs, const float2* __restrict__ a1,
                                              const float2* __restrict__ b1, float2* __restrict__ out1, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride) {
        const float4 x = __ldg(a + i), y = __ldg(b + i);              // two pixels per thread
        const float2 p = cmul_via_double(make_float2(x.x, x.y), make_float2(y.x, y.y));
        const float2 q = cmul_via_double(make_float2(x.z, x.w), make_float2(y.z, y.w));
        out[i] = make_float4(p.x, p.y, q.x, q.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out1[n - 1] = cmul_via_double(a1[n - 1], b1[n - 1]);
}

cudaError_t launch_cmul(const float2* a, const float2* b, float2* out, long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long npairs = n >> 1;
    long blocks = (npairs + 255) / 256;
    const long cap = (long)nsm * 8 * 4;                               // grid-stride above 32 CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_cmul<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                             reinterpret_cast<float4*>(out), npairs, a, b, out, n);
    return cudaGetLastError();
}

}  // namespace fringe
