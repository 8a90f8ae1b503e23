// Datum adjustment of the sequential estimator: the wrapped time series is
//     adjusted(date) = ministack phasor(date) * datum phasor(ministack of that date)
// python/adjustMiniStacks.py:180-199 leaves that product to GDAL's "mul" VRT pixel function (complex
// sources are multiplied in double and written back as CFloat32); here it is one streaming kernel.
// HBM-bound: 16 bytes in, 8 bytes out per pixel.
#include <algorithm>

#include <math_constants.h>

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

// one pixel per thread, for operands that are only 8-byte aligned (odd row offsets inside a larger raster)
__global__ void __launch_bounds__(256) k_cmul_scalar(const float2* __restrict__ a, const float2* __restrict__ b,
                                                     float2* __restrict__ out, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = cmul_via_double(__ldg(a + i), __ldg(b + i));
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
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) != 0) {
        long sb = (n + 255) / 256;
        if (sb > cap) sb = cap;
        k_cmul_scalar<<<(unsigned)sb, 256, 0, st>>>(a, b, out, n);
        return cudaGetLastError();
    }
    k_cmul<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                             reinterpret_cast<float4*>(out), npairs, a, b, out, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// despeck: SHP-weighted average of one band's amplitude or of an interferogram,
// optionally normalised to a coherence.  src/despeck/despeck.cpp:321-361 (per-block preparation)
// and :387-432 (pixel loop).  Two kernels: the preparation is elementwise, the average is a gather
// of up to (2Nx+1)(2Ny+1) prepared samples per pixel (L1/L2-resident: neighbouring pixels share
// almost all of them).  All sums are float additions in window raster order and every product /
// quotient is a single correctly rounded float operation, as in the reference, so the result is
// bit-identical to the CPU loop.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float hypotf_exact(float2 z) {          // glibc hypotf, see nmap_kernels.cu
    if (isinf(z.x) || isinf(z.y)) return __int_as_float(0x7f800000);
    return (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)z.x, (double)z.x), __dmul_rn((double)z.y, (double)z.y)));
}

// mode 0: one band, d1 = |z1| ; mode 1: d1 = z1 * conj(z2) ; mode 2: mode 1 plus d2 = (|z1|^2, |z2|^2)
__global__ void __launch_bounds__(256) k_despeck_prep(const float2* __restrict__ z1, const float2* __restrict__ z2,
                                                      long n, int mode, float2* __restrict__ d1,
                                                      float2* __restrict__ d2) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 a = __ldg(z1 + i);
        if (mode == 0) { d1[i] = make_float2(hypotf_exact(a), 0.f); continue; }
        const float2 b = __ldg(z2 + i);
        // complex<float> a *= conj(b): four float products, one float add / subtract each (libgcc __mulsc3)
        const float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
        const float im = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
        d1[i] = make_float2(re, im);
        if (mode == 2) {
            const float p = hypotf_exact(a), q = hypotf_exact(b);
            d2[i] = make_float2(__fmul_rn(p, p), __fmul_rn(q, q));
        }
    }
}

struct DespeckArgs {
    const float2* d1;
    const float2* d2;
    const uint32_t* wts;
    int cols, lines, Nx, Ny, nulong, first_line, n_lines, mode;
    float2* out;
};

__global__ void __launch_bounds__(256) k_despeck(const DespeckArgs a) {
    const long total = (long)a.n_lines * a.cols;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long pp = (long)a.first_line * a.cols + i;
    const int ci = (int)(pp / a.cols), cj = (int)(pp - (long)ci * a.cols);
    const uint32_t* cen = a.wts + pp * a.nulong;
    const int WX = 2 * a.Nx + 1, center = a.Ny * WX + a.Nx;
    float2 res = make_float2(0.f, 0.f);
    if ((__ldg(cen + (center >> 5)) >> (center & 31)) & 1u) {
        const int xmin = max(cj - a.Nx, 0), xmax = min(a.cols - 1, cj + a.Nx);
        const int ymin = max(ci - a.Ny, 0), ymax = min(a.lines - 1, ci + a.Ny);
        float vr = 0.f, vi = 0.f, sr = 0.f, si = 0.f;
        for (int ii = ymin; ii <= ymax; ++ii) {
            int f = (ii - ci + a.Ny) * WX + (xmin - cj + a.Nx);
            uint32_t word = __ldg(cen + (f >> 5));
            for (int jj = xmin; jj <= xmax; ++jj, ++f) {
                if ((f & 31) == 0) word = __ldg(cen + (f >> 5));
                if ((word >> (f & 31)) & 1u) {
                    const long q = (long)ii * a.cols + jj;
                    const float2 v = __ldg(a.d1 + q);
                    vr = __fadd_rn(vr, v.x); vi = __fadd_rn(vi, v.y);
                    if (a.mode == 2) {
                        const float2 w = __ldg(a.d2 + q);
                        sr = __fadd_rn(sr, w.x); si = __fadd_rn(si, w.y);
                    } else {
                        sr = __fadd_rn(sr, 1.0f);                   // weights 1 + 0i
                    }
                }
            }
        }
        if (sr > 0.f) {
            if (a.mode == 2) {
                if (si > 0.f) {
                    const float den = __fmul_rn(__fsqrt_rn(sr), __fsqrt_rn(si));
                    res = make_float2(__fdiv_rn(vr, den), __fdiv_rn(vi, den));
                }
            } else if (a.mode >= 0) {
                res = make_float2(__fdiv_rn(vr, sr), __fdiv_rn(vi, sr));
            }
        }
    }
    a.out[pp] = res;
}

// The same average with the window's samples staged in shared memory: a CTA of TW x TH = 32 x 8 pixels loads its tile plus the
// window halo once (zero outside the block: such positions are never added, see `in`), the mask words of a pixel sit in
// registers, and the walk over the window is branch-free -- the sum is formed with and without each sample and the mask bit
// selects (a select, not an addition of zero: -0 stays -0).  Additions and their order are those of k_despeck, so the
// result is bit-identical; the gather through L1 with a divergent branch per window position is what made that one slow.
constexpr int DS_TW = 32, DS_TH = 8, DS_MAXWORDS = 8;
template <bool MODE2>
__global__ void __launch_bounds__(DS_TW * DS_TH) k_despeck_tile(const DespeckArgs a) {
    extern __shared__ __align__(16) unsigned char ds_raw[];
    const int WX = 2 * a.Nx + 1, WY = 2 * a.Ny + 1, center = a.Ny * WX + a.Nx;
    const int SW = DS_TW + 2 * a.Nx, SH = DS_TH + 2 * a.Ny;           // staged tile
    float2* s1 = reinterpret_cast<float2*>(ds_raw);
    float2* s2 = s1 + (MODE2 ? SW * SH : 0);
    const int tx = threadIdx.x & (DS_TW - 1), ty = threadIdx.x / DS_TW;
    const int x0 = blockIdx.x * DS_TW, y0 = a.first_line + blockIdx.y * DS_TH;
    for (int k = threadIdx.x; k < SW * SH; k += DS_TW * DS_TH) {
        const int sy = k / SW, sx = k - sy * SW;
        const int yy = y0 + sy - a.Ny, xx = x0 + sx - a.Nx;
        const bool in = yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
        const long q = (long)yy * a.cols + xx;
        s1[k] = in ? __ldg(a.d1 + q) : make_float2(0.f, 0.f);
        if (MODE2) s2[k] = in ? __ldg(a.d2 + q) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int cj = x0 + tx, ci = y0 + ty;
    if (cj >= a.cols || ci >= a.first_line + a.n_lines) return;
    const long pp = (long)ci * a.cols + cj;
    uint32_t wd[DS_MAXWORDS];
#pragma unroll
    for (int w = 0; w < DS_MAXWORDS; ++w) wd[w] = (w < a.nulong) ? __ldg(a.wts + pp * a.nulong + w) : 0u;
    auto word = [&](int w) -> uint32_t {                 // register select, no dynamic indexing
        uint32_t v = wd[0];
#pragma unroll
        for (int k = 1; k < DS_MAXWORDS; ++k) v = (w == k) ? wd[k] : v;
        return v;
    };
    float2 res = make_float2(0.f, 0.f);
    if ((word(center >> 5) >> (center & 31)) & 1u) {
        // window positions outside the block are never added (the reference clamps its loops): their mask bits are cleared
        // up front, for the pixels near the border only, so that the walk tests nothing but the bit (measured against a
        // range test inside the walk: 0.26 vs 0.31 ms per 6 M pixels)
        if (ci < a.Ny || ci >= a.lines - a.Ny || cj < a.Nx || cj >= a.cols - a.Nx) {
            int f = 0;
            for (int dy = -a.Ny; dy <= a.Ny; ++dy)
                for (int dx = -a.Nx; dx <= a.Nx; ++dx, ++f) {
                    const int yy = ci + dy, xx = cj + dx;
                    if (yy < 0 || yy >= a.lines || xx < 0 || xx >= a.cols) {
#pragma unroll
                        for (int k = 0; k < DS_MAXWORDS; ++k) if ((f >> 5) == k) wd[k] &= ~(1u << (f & 31));
                    }
                }
        }
        float vr = 0.f, vi = 0.f, sr = 0.f, si = 0.f;
        int f = 0;
        uint32_t cur = wd[0];
        for (int dy = 0; dy < WY; ++dy) {
            const float2* r1 = s1 + (ty + dy) * SW + tx;
            const float2* r2 = s2 + (ty + dy) * SW + tx;
#pragma unroll 4
            for (int dx = 0; dx < WX; ++dx, ++f) {
                if ((f & 31) == 0) cur = word(f >> 5);                       // f is the same in every thread
                const bool on = (cur >> (f & 31)) & 1u;
                const float2 v = r1[dx];
                const float nvr = __fadd_rn(vr, v.x), nvi = __fadd_rn(vi, v.y);
                vr = on ? nvr : vr; vi = on ? nvi : vi;
                if (MODE2) {
                    const float2 w = r2[dx];
                    const float nsr = __fadd_rn(sr, w.x), nsi = __fadd_rn(si, w.y);
                    sr = on ? nsr : sr; si = on ? nsi : si;
                } else {
                    const float nsr = __fadd_rn(sr, 1.0f);                  // weights 1 + 0i
                    sr = on ? nsr : sr;
                }
            }
        }
        if (sr > 0.f) {
            if (MODE2) {
                if (si > 0.f) {
                    const float den = __fmul_rn(__fsqrt_rn(sr), __fsqrt_rn(si));
                    res = make_float2(__fdiv_rn(vr, den), __fdiv_rn(vi, den));
                }
            } else if (a.mode >= 0) {
                res = make_float2(__fdiv_rn(vr, sr), __fdiv_rn(vi, sr));
            }
        }
    }
    a.out[pp] = res;
}

// mode: 0 one band, 1 interferogram, 2 interferogram coherence, 3 one band + coherence flag (the
// reference then divides by sqrt(sum 1) * sqrt(0) only if the imaginary weight sum is positive: never)
cudaError_t launch_despeck(const float2* z1, const float2* z2, const uint32_t* wts, int cols, int lines, int Nx,
                           int Ny, int first_line, int n_lines, int mode, float2* d1, float2* d2, float2* out,
                           cudaStream_t st) {
    const long npix = (long)cols * lines;
    if (npix <= 0 || n_lines <= 0) return cudaSuccess;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long blocks = (npix + 255) / 256;
    if (blocks > (long)nsm * 32) blocks = (long)nsm * 32;
    const int prep_mode = (mode == 3) ? 0 : mode;
    k_despeck_prep<<<(unsigned)blocks, 256, 0, st>>>(z1, z2, npix, prep_mode, d1, d2);
    DespeckArgs a;
    a.d1 = d1; a.d2 = d2; a.wts = wts; a.cols = cols; a.lines = lines; a.Nx = Nx; a.Ny = Ny;
    a.nulong = ((2 * Ny + 1) * (2 * Nx + 1) + 31) / 32;
    a.first_line = first_line; a.n_lines = n_lines;
    a.mode = (mode == 3) ? -1 : mode;                            // -1: numerator summed, result stays zero
    a.out = out;
    // staged-tile kernel when the window's mask fits eight registers and tile + halo fits shared memory
    const size_t tile_bytes = (size_t)(DS_TW + 2 * Nx) * (DS_TH + 2 * Ny) * sizeof(float2) * (a.mode == 2 ? 2 : 1);
    if (a.nulong <= DS_MAXWORDS && tile_bytes <= 96 * 1024) {
        const dim3 grid((unsigned)((cols + DS_TW - 1) / DS_TW), (unsigned)((n_lines + DS_TH - 1) / DS_TH));
        cudaError_t e;
        if (a.mode == 2) {
            e = cudaFuncSetAttribute(k_despeck_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
            if (e != cudaSuccess) return e;
            k_despeck_tile<true><<<grid, DS_TW * DS_TH, tile_bytes, st>>>(a);
        } else {
            e = cudaFuncSetAttribute(k_despeck_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
            if (e != cudaSuccess) return e;
            k_despeck_tile<false><<<grid, DS_TW * DS_TH, tile_bytes, st>>>(a);
        }
        return cudaGetLastError();
    }
    const long total = (long)n_lines * cols;
    k_despeck<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// ampdispersion: mean calibrated amplitude and amplitude dispersion of every
// pixel over the stack, src/ampdispersion/ampdispersion.cpp:207-247.  One thread per pixel streams
// the band-major planes (coalesced 8-byte loads), sums in double in band order with the reference's
// operation sequence, no contraction: HBM-bound, 8 N bytes in, 8 bytes out per pixel.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ampdispersion(const float2* __restrict__ slc, const double* __restrict__ alpha,
                                                       long npix, int bands, float* __restrict__ da,
                                                       float* __restrict__ meanamp) {
    // valid / alpha[b] only takes two values per band (valid is 0 or 1): both quotients once per block instead of a
    // double-precision division per sample, the same correctly rounded numbers
    extern __shared__ double s_q[];                         // [bands][2]: 0 / alpha, 1 / alpha
    for (int b = threadIdx.x; b < bands; b += blockDim.x) {
        const double al = alpha ? alpha[b] : 1.0;
        s_q[2 * b] = __ddiv_rn(0.0, al);
        s_q[2 * b + 1] = __ddiv_rn(1.0, al);
    }
    __syncthreads();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    double mean = 0.0, meansq = 0.0, norms = 0.0;
#pragma unroll 4
    for (int b = 0; b < bands; ++b) {
        double absval = (double)hypotf_exact(__ldg(slc + (long)b * npix + i));
        const int valid = (absval != 0.0);
        absval = __dmul_rn(absval, s_q[2 * b + valid]);
        mean = __dadd_rn(mean, absval);
        meansq = __dadd_rn(meansq, __dmul_rn(absval, absval));
        norms += (double)valid;
    }
    float m = 0.f, d = -1.f;
    if (norms > 1.0) {
        const double avg = __ddiv_rn(mean, norms), avg2 = __ddiv_rn(meansq, norms);
        const double sdev = __dsqrt_rn(__dsub_rn(avg2, __dmul_rn(avg, avg)));
        m = (float)avg;
        d = (float)((!isnan(sdev) && sdev > 0.0) ? __ddiv_rn(sdev, avg) : -1.0);
    }
    meanamp[i] = m;
    da[i] = d;
}

cudaError_t launch_ampdispersion(const float2* slc, const double* alpha, long npix, int bands, float* da, float* meanamp,
                                 cudaStream_t st) {
    if (npix <= 0) return cudaSuccess;
    k_ampdispersion<<<(unsigned)((npix + 255) / 256), 256, (size_t)bands * 2 * sizeof(double), st>>>(slc, alpha, npix, bands, da, meanamp);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// calamp: per-band amplitude calibration constant = mean |z| over the valid pixels
// (src/calamp/calamp.cpp:207-226: float hypot, NaN -> 0, valid = amplitude != 0 and mask > 0; double sums).
// The reference's OpenMP reduction has no fixed summation order, so its own result is only defined to
// rounding; here: per-thread double partials, warp and block reduction, one atomicAdd per block.
// acc: [bands][2] doubles (sum, count), accumulated across calls (the driver walks the image in blocks).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_calamp(const float2* __restrict__ slc, const uint8_t* __restrict__ mask, long npix,
                                                double* __restrict__ acc) {
    const int b = blockIdx.y;
    const float2* z = slc + (long)b * npix;
    double sum = 0.0, cnt = 0.0;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        float h = hypotf_exact(__ldg(z + i));
        if (isnan(h)) h = 0.f;
        const bool valid = (h != 0.f) && (mask ? mask[i] > 0 : true);
        if (valid) { sum += (double)h; cnt += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    __shared__ double s_sum[8], s_cnt[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_sum[w] = sum; s_cnt[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < 8; ++k) { a += s_sum[k]; c += s_cnt[k]; }
        atomicAdd(&acc[2 * b], a);
        atomicAdd(&acc[2 * b + 1], c);
    }
}

cudaError_t launch_calamp(const float2* slc, const uint8_t* mask, long npix, int bands, double* acc, cudaStream_t st) {
    if (npix <= 0 || bands <= 0) return cudaSuccess;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long bx = (npix + 255) / 256;
    const long cap = std::max<long>(1, (long)nsm * 16 / bands);           // ~16 CTAs per SM over all bands
    if (bx > cap) bx = cap;
    k_calamp<<<dim3((unsigned)bx, (unsigned)bands), 256, 0, st>>>(slc, mask, npix, acc);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// PS / DS integration (python/integratePS.py:97-130 and :134-159): the wrapped interferogram of a
// pair (i, j) is the product of the adjusted DS phasors, ds_j * conj(ds_i), except at PS pixels, where it is the
// full-resolution interferogram reduced to unit modulus, exp(1j * angle(slc_j * conj(slc_i))); the coherence raster
// gets a fixed value at PS pixels.  All arithmetic in complex64 / float32 like the numpy expressions; the unit phasor
// is formed as z / |z| (a zero product gives +-1 by the sign of its real part, as atan2 does).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_integrate_ps(const float2* __restrict__ ds_i, const float2* __restrict__ ds_j,
                                                      const float2* __restrict__ slc_i, const float2* __restrict__ slc_j,
                                                      const uint8_t* __restrict__ ps, long n, float2* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const bool is_ps = ps[k] == 1;
        const float2 a = __ldg((is_ps ? slc_j : ds_j) + k), b = __ldg((is_ps ? slc_i : ds_i) + k);
        // complex64 product a * conj(b), each operation rounded to float (numpy's complex64 multiply)
        float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
        float im = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
        if (is_ps) {
            const float m = hypotf_exact(make_float2(re, im));
            if (m > 0.f && !isinf(m)) { re = __fdiv_rn(re, m); im = __fdiv_rn(im, m); }
            else if (m == 0.f) {
                // angle of a zero product: atan2(+-0, +0) = +-0 -> 1 + 0j, but atan2(+-0, -0) = +-pi -> exp(1j * (float)pi)
                // = -1 -+ 8.74e-8j (the float32 pi is not pi).  A zero SLC sample times a negative one gives -0.
                if (signbit(re)) { im = signbit(im) ? 8.742278e-8f : -8.742278e-8f; re = -1.f; }
                else { re = 1.f; }
            }
            else { re = CUDART_NAN_F; im = CUDART_NAN_F; }
        }
        out[k] = make_float2(re, im);
    }
}

__global__ void __launch_bounds__(256) k_ps_coherence(const float* __restrict__ tcorr, const uint8_t* __restrict__ ps, long n,
                                                      float value, float* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = (ps[k] == 1) ? value : tcorr[k];
}

static unsigned stream_grid(long n) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long b = (n + 255) / 256;
    if (b > (long)nsm * 32) b = (long)nsm * 32;
    return (unsigned)std::max<long>(b, 1);
}

cudaError_t launch_integrate_ps(const float2* ds_i, const float2* ds_j, const float2* slc_i, const float2* slc_j, const uint8_t* ps,
                                long n, float2* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_integrate_ps<<<stream_grid(n), 256, 0, st>>>(ds_i, ds_j, slc_i, slc_j, ps, n, out);
    return cudaGetLastError();
}

cudaError_t launch_ps_coherence(const float* tcorr, const uint8_t* ps, long n, float value, float* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_ps_coherence<<<stream_grid(n), 256, 0, st>>>(tcorr, ps, n, value, out);
    return cudaGetLastError();
}

}  // namespace fringe
