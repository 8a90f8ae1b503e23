// Tensor-pipe covariance (two-term FP16 split on mma.sync.m16n8k16) + in-register power iteration (sm_100a), bands <= 32.
//
// EVD / STBAS with evd.cpp's control flow.  The masked Gram product C = sum_k z_k z_k^H runs on the warp-level tensor
// path, the eigen solve and the epilogue on FP32 FMAs:
//
//   layout      k_layout_f16 multiplies every band by a power of two (k_band_scale: the band's typical magnitude lands
//               near 2^6, so FP16's range covers +-57 dB around it; the coherence is invariant to a per-band factor) and
//               splits every sample into an FP16 "hi" part and an FP16 "lo" remainder (x = hi + lo to ~2^-22, the same
//               accuracy as the 3xTF32 split this replaces, at half the tensor work and half the bytes).  A pixel's 32
//               (zero padded) bands are 64 words: word 4 g + q = half2(re, im) of band g + 8 q, first all hi (128 B),
//               then all lo.  Lane (g, t) of a warp fetches its share of one SHP with four 8-byte loads (bands g, g+8 and
//               g+16, g+24, hi and lo), and the eight lanes that share t read the SHP's 128-byte hi (lo) row contiguously.
//               Pixels whose scaled samples leave the FP16 range (|x| >= 65504, NaN, or everything below 2^-8) get an
//               all-NaN hi row; the NaN surfaces in the band powers of every pixel that has such an SHP, and that pixel
//               recomputes its Gram product from the original planes on FP32 FMAs.
//   covariance  One warp per pixel.  Eight SHPs form one k-chunk of m16n8k16: k = (2t, 2t+1) are (re, im) of SHP t,
//               k = (2t+8, 2t+9) those of SHP t+4, so a half2 sample *is* an A-fragment register (an A quad = the 8-byte
//               loads of SHP t and SHP t+4 side by side, no assembly) and
//                   Re C = A B,  B[(re,im) of q][j] = ( re_j,  im_j)  -- the same registers again
//                   Im C = A B', B'[(re,im) of q][j] = (-im_j,  re_j)  -- halves swapped, one sign flipped
//               Only the 6 of 8 16x8 tiles that touch the upper triangle are computed: 12 accumulator tiles
//               (48 registers) x 3 products (hi*hi + hi*lo + lo*hi) = 36 mma.sync per 8 SHPs, issued product-major so
//               that no instruction waits for the one before it.
//   hand-off    accumulator fragments -> coherence -> planar (re | im) Hermitian matrix in shared
//               memory.
//   eigen       lane = row; the row lives in registers as pairs of consecutive columns and the
//               matrix-vector product runs on fma.rn.f32x2; heavy-ball momentum switched on from the
//               first observed residual decay; Rayleigh quotient and residual are single-round warp
//               reductions in consecutive iterations, scheduled where convergence is predicted.
//   epilogue    phase reference / compressed SLC / temporal coherence from the register-resident row.
//   schedule    one 16-warp CTA per SM; a CTA owns a band of 4 rows x a column segment and its warps
//               draw pixels from a shared counter in column-major order, so the union of their
//               windows stays in L1.
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace fringe {

#define FULLMASK 0xffffffffu

#ifdef FRINGE_PHASE_CLOCKS
#define PHASE_DECL long long ph_t = clock64(); unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
#define PHASE_MARK(k) { const long long ph_n = clock64(); ph[k] += (unsigned long long)(ph_n - ph_t); ph_t = ph_n; }
#define PHASE_FLUSH if (a.stats && lane == 0) { for (int k = 0; k < 6; ++k) atomicAdd(&a.stats[8 + k], ph[k]); }
#else
#define PHASE_DECL
#define PHASE_MARK(k)
#define PHASE_FLUSH
#endif

namespace {

__device__ __forceinline__ float fast_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// SHPs per k-chunk of m16n8k16.  (Every mma.sync shape costs the same tensor-pipe time on this part -- m16n8k8 TF32,
// m16n8k8 FP16 and m16n8k16 FP16 all measure one instruction per 8.5 cycles and SM sub-partition, fringe_prof_mma_* -- so
// the k = 16 shape halves the pipe time of the product; an m16n8k8 FP16 version, whose operands need no assembly at all,
// measured 210 ms per 30 M pixels against 199 ms for this one.)
constexpr int CHUNK = 8;

// warps per CTA (one CTA per SM).  16 leaves 128 registers per thread; 20 would leave 96 and spills ~100 bytes per thread
#ifndef FRINGE_MMA_WARPS
#define FRINGE_MMA_WARPS 16
#endif

// D(16x8) += A(16x16, row) * B(16x8, col), FP16 inputs, FP32 accumulate
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                        uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// m16n8k16: the A quad of row block I is (bands 16 I + g, 16 I + g + 8 of SHP t | the same of SHP t + 4): two 8-byte loads
// whose destinations can sit side by side, so the quads need no assembly
struct MmaOperands {
    uint2 h[2][2], l[2][2];            // [row block I][SHP slot]
    __device__ __forceinline__ void load(const uint32_t* __restrict__ zh, const int (&idx)[2], int g) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint2* p = reinterpret_cast<const uint2*>(zh + (long)idx[k] * 64) + 2 * g;
            h[0][k] = __ldg(p); h[1][k] = __ldg(p + 1); l[0][k] = __ldg(p + 16); l[1][k] = __ldg(p + 17);
        }
    }
    // band 8 J + g of SHP slot k
    __device__ __forceinline__ uint32_t bh(int J, int k) const { return (J & 1) ? h[J >> 1][k].y : h[J >> 1][k].x; }
    __device__ __forceinline__ uint32_t bl(int J, int k) const { return (J & 1) ? l[J >> 1][k].y : l[J >> 1][k].x; }
};
// (re, im) -> (-im, re)
__device__ __forceinline__ uint32_t rot90(uint32_t w) { return __byte_perm(w, 0, 0x1032) ^ 0x00008000u; }

// tiles (I, J) on or above the diagonal of the 2 x 4 grid of 16 x 8 tiles
__device__ __forceinline__ constexpr int tile_i(int tl) { return tl < 4 ? 0 : 1; }
__device__ __forceinline__ constexpr int tile_j(int tl) { return tl < 4 ? tl : tl - 2; }

// CHUNK SHPs into the 6 real and 6 imaginary accumulator tiles: 36 mma.sync.
// Row block I of A = bands 16 I + g and 16 I + g + 8; column block J of B = band 8 J + g of the lane's own samples; the
// imaginary product takes B rotated by 90 degrees.
__device__ __forceinline__ void mma_chunk(float (&cre)[6][4], float (&cim)[6][4], const MmaOperands& o) {
    uint32_t bh[4][2], bl[4][2], nh[4][2], nl[4][2];
#pragma unroll
    for (int J = 0; J < 4; ++J) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            bh[J][k] = o.bh(J, k); bl[J][k] = o.bl(J, k);
            nh[J][k] = rot90(o.bh(J, k)); nl[J][k] = rot90(o.bl(J, k));
        }
    }
    // products outermost: the 12 accumulator tiles are independent, so consecutive mma.sync never
    // wait for each other (with the tile loop outermost every instruction would depend on the one
    // issued two slots earlier and the tensor pipe would idle for the mma latency)
#pragma unroll
    for (int prod = 0; prod < 3; ++prod) {                               // hi * hi, hi * lo, lo * hi
#pragma unroll
        for (int tl = 0; tl < 6; ++tl) {
            const int I = tile_i(tl), J = tile_j(tl);
            const uint2 a0 = (prod == 2) ? o.l[I][0] : o.h[I][0], a1 = (prod == 2) ? o.l[I][1] : o.h[I][1];
            const uint32_t* bv = (prod == 1) ? bl[J] : bh[J];
            const uint32_t* nv = (prod == 1) ? nl[J] : nh[J];
            mma_f16(cre[tl], a0.x, a0.y, a1.x, a1.y, bv[0], bv[1]);
            mma_f16(cim[tl], a0.x, a0.y, a1.x, a1.y, nv[0], nv[1]);
        }
    }
}

template <int NE>
struct MmaCfg {
    static constexpr int WARPS = FRINGE_MMA_WARPS;      // one CTA per SM: its warps share one moving window in L1
    static constexpr int BAND = 4;                      // rows of a CTA's pixel band
    // coherence matrix in shared memory: real and imaginary planes, rows NS floats apart;
    // NS = 4 (mod 8) keeps rows 16-byte aligned and spreads the hand-off stores over the banks
    static constexpr int NS = ((NE + 3) / 4 * 4) % 8 == 4 ? (NE + 3) / 4 * 4 : (NE + 3) / 4 * 4 + 4;
    // floats per plane; orders below 32 keep one extra, permanently zero row for the lanes beyond the matrix
    static constexpr int PLANE = (NE < 32 ? NE + 1 : NE) * NS;
    // per warp: two planes, two broadcast vectors (double buffered, re | im planes of 32),
    // 32 powers, 64 list slots
    static constexpr int SMEM_PER_WARP =
        (((2 * PLANE + 2 * 64 + 32) * (int)sizeof(float) + 64 * (int)sizeof(int)) + 15) & ~15;
};

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {       // two FP32 FMAs in one instruction
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 pack2(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// per-band power-of-two scale: 2^(5 - E), E = rounded mean binary exponent of the finite non-zero components of a
// sample of the first rows of the block call that hold any data (scale[b] = 0 means "still open") (a geometric mean: one absurd value cannot move it).  For Gaussian components of standard
// deviation s the mean exponent is log2(s) - 1.4, so s lands near 2^6.4.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_band_scale(const float2* __restrict__ slc, long npix, long first, long count,
                                                    float* __restrict__ scale) {
    const int b = blockIdx.x;
    if (scale[b] != 0.f) return;                               // fixed by an earlier row chunk of this block call
    __shared__ long s_sum[256];
    __shared__ int s_n[256];
    const long groups = (count + 3) / 4;                       // 4 consecutive pixels = one 32-byte sector
    const long want = 16384;                                   // sectors sampled per band
    // pass 0 samples; if it finds nothing but zeros (or non-finite values), pass 1 looks at every pixel: the scale must
    // be fixed by the first rows that hold any data at all, because rows laid out earlier keep the factor 1
    for (int pass = 0; pass < 2; ++pass) {
        const long step = (pass == 0 && groups > want) ? groups / want : 1;
        long esum = 0; int n = 0;
        for (long gi = threadIdx.x; gi * step < groups; gi += blockDim.x) {
            const long p0 = first + gi * step * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (p0 + k >= first + count) break;
                const float2 v = __ldg(&slc[(long)b * npix + p0 + k]);
                const uint32_t ex = (__float_as_uint(v.x) >> 23) & 0xffu, ey = (__float_as_uint(v.y) >> 23) & 0xffu;
                if (ex > 0 && ex < 255) { esum += (int)ex - 127; ++n; }
                if (ey > 0 && ey < 255) { esum += (int)ey - 127; ++n; }
            }
        }
        s_sum[threadIdx.x] = esum; s_n[threadIdx.x] = n;
        __syncthreads();
        for (int s2 = 128; s2 > 0; s2 >>= 1) {
            if (threadIdx.x < s2) { s_sum[threadIdx.x] += s_sum[threadIdx.x + s2]; s_n[threadIdx.x] += s_n[threadIdx.x + s2]; }
            __syncthreads();
        }
        const int total = s_n[0];
        const long sum = s_sum[0];
        __syncthreads();
        if (total > 0) {
            if (threadIdx.x == 0) {
                int e = 5 - (int)floor((double)sum / (double)total + 0.5);
                e = max(-100, min(100, e));
                scale[b] = exp2f((float)e);
            }
            return;
        }
        if (step == 1) return;                                 // the sample was already exhaustive: the scale stays open
    }
}

// ---------------------------------------------------------------------------------------
// re-layout: [bands][npix] -> [npix][hi 32 | lo 32] half2 words (see the header of this file); out-of-range pixels get
// an all-NaN hi row, which marks every Gram product they enter
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_layout_f16(const float2* __restrict__ slc, long npix, long first, long pend, int bands,
                                                    const float* __restrict__ scale, uint32_t* __restrict__ zh) {
    // lane = 8 * (pixel & 3) + g: a warp reads eight 32-byte sectors per band group and writes four 128-byte rows
    const int g = threadIdx.x & 7, sub = threadIdx.x >> 3;
    const long p = first + (long)blockIdx.x * 32 + sub;
    const bool live = p < pend;
    uint32_t h[4], l[4];
    float top = 0.f;
    bool hot = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b = g + 8 * q;
        float2 v = make_float2(0.f, 0.f);
        if (live && b < bands) {
            v = __ldg(&slc[(long)b * npix + p]);
            float sc = __ldg(&scale[b]);
            if (sc == 0.f) sc = 1.f;                           // still open: nothing but zeros so far
            v.x *= sc; v.y *= sc;
        }
        const float ax = fabsf(v.x), ay = fabsf(v.y);
        if (!(ax < 65504.f) || !(ay < 65504.f)) { hot = true; v = make_float2(0.f, 0.f); }      // also NaN
        top = fmaxf(top, fmaxf(ax, ay));
        const __half2 hi = __floats2half2_rn(v.x, v.y);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(v.x - hf.x, v.y - hf.y);                           // exact differences
        h[q] = *reinterpret_cast<const uint32_t*>(&hi);
        l[q] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    // a pixel = 8 consecutive lanes
#pragma unroll
    for (int s2 = 1; s2 < 8; s2 <<= 1) {
        top = fmaxf(top, __shfl_xor_sync(FULLMASK, top, s2));
        hot = hot | (__shfl_xor_sync(FULLMASK, (int)hot, s2) != 0);
    }
    if (!live) return;
    const bool cold = top > 0.f && top < 0.00390625f;          // everything below 2^-8: the lo parts are gone
    if (hot || cold) { h[0] = h[1] = h[2] = h[3] = 0x7e007e00u; l[0] = l[1] = l[2] = l[3] = 0u; }     // FP16 NaNs
    uint4* o = reinterpret_cast<uint4*>(zh + p * 64) + g;
    o[0] = make_uint4(h[0], h[1], h[2], h[3]);
    o[8] = make_uint4(l[0], l[1], l[2], l[3]);
}

cudaError_t launch_band_scale(const float2* slc, long npix, long first, long count, int bands, float* scale, cudaStream_t st) {
    if (count <= 0 || bands <= 0) return cudaSuccess;
    k_band_scale<<<(unsigned)bands, 256, 0, st>>>(slc, npix, first, count, scale);
    return cudaGetLastError();
}

cudaError_t launch_transpose_mma(const float2* slc, long npix, long first, long count, int bands, const float* scale,
                                 float2* zpix, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    k_layout_f16<<<(unsigned)((count + 31) / 32), 256, 0, st>>>(slc, npix, first, first + count, bands, scale,
                                                               reinterpret_cast<uint32_t*>(zpix));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
template <int NE>
__global__ void __launch_bounds__(FRINGE_MMA_WARPS * 32, 1) k_evd_mma(const EvdArgs a) {
    typedef MmaCfg<NE> Cfg;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int N = a.bands;                      // <= NE <= 32

    // CTA-wide table: window bit index -> (dy, dx)
    short2* s_off = reinterpret_cast<short2*>(s_raw);
    const int WX = 2 * a.Nx + 1, W = WX * (2 * a.Ny + 1), center = a.Ny * WX + a.Nx;
    for (int f = threadIdx.x; f < a.nulong * 32; f += blockDim.x) {
        const int fy = f / WX;
        s_off[f] = (f < W) ? make_short2((short)(fy - a.Ny), (short)(f - fy * WX - a.Nx))
                           : make_short2((short)-30000, (short)-30000);   // never in bounds
    }
    __syncthreads();
    const int lut_bytes = ((a.nulong * 32 * (int)sizeof(short2)) + 15) & ~15;

    unsigned char* base = s_raw + lut_bytes + (size_t)warp * Cfg::SMEM_PER_WARP;
    constexpr int NS = Cfg::NS;
    float* s_re = reinterpret_cast<float*>(base);                          // [NE][NS]
    float* s_im = s_re + Cfg::PLANE;                                       // [NE][NS]
    float* s_vec = s_im + Cfg::PLANE;                                      // [2][re 32 | im 32]
    float* s_pw = s_vec + 128;                                             // [32]
    int* s_list = reinterpret_cast<int*>(s_pw + 32);                       // [64]

    if (NE < 32) {                               // the zero row (never written again)
        for (int e = lane; e < NS; e += 32) { s_re[NE * NS + e] = 0.f; s_im[NE * NS + e] = 0.f; }
        __syncwarp();
    }
    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2);
    const int BW = a.bandwidth;
    // bit j set: the pair (lane, j) enters the temporal-coherence sum (j > lane, inside the
    // matrix and, for STBAS, inside the band)
    uint32_t usemask = 0u;
    for (int j = 0; j < N; ++j)
        if (j > lane && (!isstbas || (j - lane) <= BW)) usemask |= (1u << j);
    float inv_pairs;                             // 1 / number of (i<j) pairs (evd.cpp:773-784)
    {
        int cnt = 0;
        for (int i = 0; i < N; ++i) cnt += isstbas ? min(BW, N - 1 - i) : (N - 1 - i);
        inv_pairs = 1.0f / (float)cnt;
    }
    const long npix_block = (long)a.cols * a.lines;
    const uint32_t* zh = reinterpret_cast<const uint32_t*>(a.zpix);

    // Work: the CTA owns a band of BAND rows x a segment of columns and its warps draw pixels from
    // a shared counter in column-major order inside the band (k -> row k % rows, column k / rows).
    // All 16 warps therefore sit inside a ~4 x 4 pixel patch that slides along the band: the union
    // of their windows (~14 x 8 samples, 57 kB) stays in L1 and every new column of samples is
    // fetched once per band instead of once per row.
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int nbands = (a.n_lines + Cfg::BAND - 1) / Cfg::BAND;
    const int band = blockIdx.x % nbands, seg = blockIdx.x / nbands;
    const int seglen = a.tile_pairs;                          // columns per segment (set by the launcher)
    const int c0 = seg * seglen, c1 = min(a.cols, c0 + seglen);
    const int row0 = a.first_line + band * Cfg::BAND;
    const int rows = min(Cfg::BAND, a.first_line + a.n_lines - row0);
    const int total = (c1 - c0) * rows;
    unsigned long long st_pix = 0, st_it = 0, st_cap = 0, st_hot = 0;
    PHASE_DECL

    auto draw = [&]() -> int {
        int k = 0;
        if (lane == 0) k = atomicAdd(&s_next, 1);
        return __shfl_sync(FULLMASK, k, 0);
    };
    const bool in_regs = a.nulong <= 32;
    auto load_mask_word = [&](int k) -> uint32_t {            // lane w keeps word w of the pixel's mask
        if (k >= total || lane >= a.nulong) return 0u;
        const long pp = (long)(row0 + k % rows) * a.cols + c0 + k / rows;
        return __ldg(&a.wts[pp * a.nulong + lane]);
    };
    int k_next = draw();
    uint32_t mw_next = load_mask_word(k_next);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int zero_idx = (int)npix_block;         // index of the all-zero sample vector

#pragma unroll 1
    while (k_next < total) {
        const int k = k_next;
        const int row = row0 + k % rows;
        const int col = c0 + k / rows;
        const long pg = (long)row * a.cols + col;
        const uint32_t mw = mw_next;
        k_next = draw();
        mw_next = load_mask_word(k_next);
        auto mask_word = [&](int w) -> uint32_t {               // w warp-uniform
            if (w >= a.nulong) return 0u;
            return in_regs ? __shfl_sync(FULLMASK, mw, w) : __ldg(&a.wts[pg * a.nulong + w]);
        };
        const bool center_on = (mask_word(center >> 5) >> (center & 31)) & 1u;

        // ------------------------- covariance (evd.cpp:537-564) -------------------------
        float cre[6][4], cim[6][4];
#pragma unroll
        for (int tl = 0; tl < 6; ++tl)
#pragma unroll
            for (int e = 0; e < 4; ++e) { cre[tl][e] = 0.f; cim[tl][e] = 0.f; }
        int npix = 0;
#pragma unroll 1
        for (int w0 = 0; w0 < a.nulong; w0 += 2) {
            int n = 0;
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int w = w0 + c;
                const uint32_t word = center_on ? mask_word(w) : 0u;
                const short2 d = s_off[min(w, a.nulong - 1) * 32 + lane];
                const int yy = row + d.x, xx = col + d.y;
                const bool ok = ((word >> lane) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                const uint32_t V = __ballot_sync(FULLMASK, ok);
                if (ok) s_list[n + __popc(V & lt_mask)] = yy * a.cols + xx;
                n += __popc(V);
            }
            npix += n;
            const int chunks = (n + CHUNK - 1) / CHUNK;
            if (lane < CHUNK && n + lane < CHUNK * chunks) s_list[n + lane] = zero_idx;
            __syncwarp();
            PHASE_MARK(0)
            // no register double buffering of the operands: the other 15 warps cover the load latency (measured equal
            // to a prefetching version with fewer resident warps); only the next list entries are fetched ahead
            MmaOperands op;
            int idx[CHUNK / 4];
#pragma unroll
            for (int k = 0; k < CHUNK / 4; ++k) idx[k] = s_list[t + 4 * k];
#pragma unroll 1
            for (int c = 0; c < chunks; ++c) {
                op.load(zh, idx, g);
                const int cn = CHUNK * min(c + 1, chunks - 1) + t;
#pragma unroll
                for (int k = 0; k < CHUNK / 4; ++k) idx[k] = s_list[cn + 4 * k];
                mma_chunk(cre, cim, op);
            }
            PHASE_MARK(1)
        }
        const bool solve = center_on && (npix >= 2);         // evd.cpp:566 hard-codes 2

        // ------------------------- coherence (evd.cpp:569-582) --------------------------
        // accumulator fragment of tile (I, J): element e is row 16 I + g + 8 (e >> 1), column 8 J + 2 t + (e & 1)
        __syncwarp();
        bool zero_band = false;
        // band powers = diagonal of Re C: lanes with g in {2t, 2t+1} hold them.  A NaN power means an SHP of this pixel
        // carries the out-of-range marker of k_layout_f16 (its hi row is all NaN): FP32 recomputation below.
        float p0 = 1.f, p1 = 1.f, p2 = 1.f, p3 = 1.f;
        const bool diag = (g == 2 * t || g == 2 * t + 1);
        if (diag) {
            const bool odd = (g != 2 * t);              // selects, not a runtime index: keeps the tiles in registers
            p0 = odd ? cre[0][1] : cre[0][0]; p1 = odd ? cre[1][3] : cre[1][2];
            p2 = odd ? cre[4][1] : cre[4][0]; p3 = odd ? cre[5][3] : cre[5][2];
        }
        const bool hot = __any_sync(FULLMASK, (p0 != p0) || (p1 != p1) || (p2 != p2) || (p3 != p3));
        if (!hot) {
            if (diag) {
                // padded bands get +inf so that their scaled entries come out as exact zeros
                s_pw[g] = (g < N) ? sqrtf(p0) : CUDART_INF_F;
                s_pw[g + 8] = (g + 8 < N) ? sqrtf(p1) : CUDART_INF_F;
                s_pw[g + 16] = (g + 16 < N) ? sqrtf(p2) : CUDART_INF_F;
                s_pw[g + 24] = (g + 24 < N) ? sqrtf(p3) : CUDART_INF_F;
                zero_band = (g < N && !(p0 > 0.f)) || (g + 8 < N && !(p1 > 0.f)) || (g + 16 < N && !(p2 > 0.f)) || (g + 24 < N && !(p3 > 0.f));
            }
            // a band that is zero in every SHP puts NaNs into C.  The reference hands that matrix to zheevr, which (OpenBLAS)
            // reports success with an undefined vector; its temporal coherence then comes out as NaN (arg of NaN entries,
            // evd.cpp:770-786).  Here: temporal coherence NaN as well, phasors and compressed SLC 0 instead of LAPACK's
            // undefined values.  Arises only when a whole window is zero in one date.
            zero_band = __any_sync(FULLMASK, zero_band);
            __syncwarp();
            {
                float ir[4], ic[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) ir[q] = fast_rcp(s_pw[g + 8 * q]);
#pragma unroll
                for (int J = 0; J < 4; ++J) { ic[2 * J] = fast_rcp(s_pw[8 * J + 2 * t]); ic[2 * J + 1] = fast_rcp(s_pw[8 * J + 2 * t + 1]); }
#pragma unroll
                for (int tl = 0; tl < 6; ++tl) {
                    const int I = tile_i(tl), J = tile_j(tl);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = 16 * I + g + 8 * (e >> 1), c = 8 * J + 2 * t + (e & 1);
                        const float s = ir[2 * I + (e >> 1)] * ic[2 * J + (e & 1)];
                        if (c > r && c < NE) {                   // strict upper entry and its mirror
                            const float vr = cre[tl][e] * s, vi = cim[tl][e] * s;
                            s_re[r * NS + c] = vr; s_im[r * NS + c] = vi;
                            s_re[c * NS + r] = vr; s_im[c * NS + r] = -vi;
                        }
                    }
                }
                if (lane < NE) { s_re[lane * NS + lane] = (lane < N) ? 1.f : 0.f; s_im[lane * NS + lane] = 0.f; }
            }
        } else {
            ++st_hot;
            // A sample of this pixel's neighbourhood left the FP16 range: FP32 recomputation from the original planes
            // (evd.cpp:537-582 on FMAs).  Lane j accumulates column j of the upper triangle in shared memory.
            for (int e = lane; e < 2 * Cfg::PLANE; e += 32) s_re[e] = 0.f;          // both planes (contiguous)
            float* zv = s_vec;
            __syncwarp();
            for (int w = 0; w < a.nulong; ++w) {
                uint32_t word = mask_word(w);
                while (word) {                                                         // warp-uniform
                    const int bit = __ffs(word) - 1;
                    word &= word - 1;
                    const short2 d = s_off[w * 32 + bit];
                    const int yy = row + d.x, xx = col + d.y;
                    if (yy < 0 || yy >= a.lines || xx < 0 || xx >= a.cols) continue;
                    float2 z = make_float2(0.f, 0.f);
                    if (lane < N) z = __ldg(&a.slc[(long)lane * npix_block + (long)yy * a.cols + xx]);
                    __syncwarp();
                    zv[lane] = z.x; zv[32 + lane] = z.y;
                    __syncwarp();
                    if (lane < N)
                        for (int i = 0; i <= lane; ++i) {
                            const float zr = zv[i], zi = zv[32 + i];
                            s_re[i * NS + lane] += zr * z.x + zi * z.y;                 // z_i * conj(z_j)
                            s_im[i * NS + lane] += zi * z.x - zr * z.y;
                        }
                }
            }
            __syncwarp();
            const float pw = (lane < N) ? sqrtf(s_re[min(lane, NE - 1) * NS + min(lane, NE - 1)]) : CUDART_INF_F;
            s_pw[lane] = pw;
            zero_band = __any_sync(FULLMASK, lane < N && !(pw > 0.f));
            __syncwarp();
            if (lane < N) {
                const float ij = 1.0f / pw;
                for (int i = 0; i < lane; ++i) {
                    const float sc = ij / s_pw[i];
                    const float vr = s_re[i * NS + lane] * sc, vi = s_im[i * NS + lane] * sc;
                    s_re[i * NS + lane] = vr; s_im[i * NS + lane] = vi;
                    s_re[lane * NS + i] = vr; s_im[lane * NS + i] = -vi;
                }
            }
            __syncwarp();
            if (lane < NE) { s_re[lane * NS + lane] = (lane < N) ? 1.f : 0.f; s_im[lane * NS + lane] = 0.f; }
        }
        __syncwarp();
        PHASE_MARK(2)

        // ------------------------- eigen + epilogue --------------------------------------
        float2 o = make_float2(0.f, 0.f);
        float tc = 0.f;
        float2 cmp = make_float2(0.f, 0.f);
        if (solve && zero_band) tc = CUDART_NAN_F;
        if (solve && !zero_band) {
            ++st_pix;
            // the centre pixel's own samples for the compressed SLC: fetched now, needed after the iteration
            float2 zc = make_float2(0.f, 0.f);
            if (lane < N && lane >= k0) zc = __ldg(&a.slc[(long)lane * npix_block + pg]);
            const int r = (lane < NE) ? lane : NE;           // lanes beyond the matrix read the zero row
            // row r of the matrix as pairs of consecutive columns: cr2[k] = (Re C[r][2k], Re C[r][2k+1])
            constexpr int NP2 = NE / 2;
            u64 cr2[NP2], ci2[NP2];
            {
                const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(s_re + r * NS);
                const ulonglong2* pi = reinterpret_cast<const ulonglong2*>(s_im + r * NS);
#pragma unroll
                for (int k = 0; k + 1 < NP2; k += 2) {
                    const ulonglong2 vr = pr[k >> 1], vi = pi[k >> 1];
                    cr2[k] = vr.x; cr2[k + 1] = vr.y; ci2[k] = vi.x; ci2[k + 1] = vi.y;
                }
                if (NP2 & 1) {
                    cr2[NP2 - 1] = *reinterpret_cast<const u64*>(s_re + r * NS + NE - 2);
                    ci2[NP2 - 1] = *reinterpret_cast<const u64*>(s_im + r * NS + NE - 2);
                }
            }
            // rows >= N of the padded matrix are exact zeros, and so is the row of the lanes beyond the padded order
            const float live = (lane < NE) ? 1.f : 0.f;
            if (isstbas) {                                   // evd.cpp:695-706 band limit
#pragma unroll
                for (int k = 0; k < NP2; ++k) {
                    float2 vr = unpack2(cr2[k]), vi = unpack2(ci2[k]);
                    if (abs(2 * k - lane) > BW) { vr.x = 0.f; vi.x = 0.f; }
                    if (abs(2 * k + 1 - lane) > BW) { vr.y = 0.f; vi.y = 0.f; }
                    cr2[k] = pack2(vr.x, vr.y); ci2[k] = pack2(vi.x, vi.y);
                }
            }
            // start vector: the middle column of C reduced to unit modulus (the dominant eigenvector of
            // a coherence matrix has nearly uniform magnitudes, and the middle date is the one most
            // coherent with all others: half an iteration fewer than the first column in the replay)
            float2 x;
            {
                const int ks = N >> 1;
                const int rc = min(lane, NE - 1);
                const float2 v = make_float2(s_re[ks * NS + rc], s_im[ks * NS + rc]);
                const float keep = (isstbas && abs(ks - lane) > BW) ? 0.f : live;
                x = make_float2(v.x * keep, -v.y * keep);
                const float m2 = x.x * x.x + x.y * x.y;
                const float rs = (m2 > 0.f) ? fast_rsqrt(m2) : 0.f;
                x.x *= rs; x.y *= rs;
                float n2 = x.x * x.x + x.y * x.y;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(FULLMASK, n2, s);
                const float sc = rsqrtf(n2);
                x.x *= sc; x.y *= sc;
            }
            PHASE_MARK(3)
            // Power iteration with heavy-ball momentum (see experimental/evd_fast_fp32_fma.cu): x+ = C x / lambda - beta x-
            float lam = 1.f, inv_lam = 1.f, beta = 0.f, rho_prev = -1.f;
            float2 xp = make_float2(0.f, 0.f);
            int it = 0, buf = 0, next_chk = 3, gap = 2;
            bool conv = false;
            const int kMaxIter = 1000;
            const float tol2 = 1.0e-12f;                 // relative residual 1e-6: 3x margin on the 1e-3 rad gate in the worst test pixel
#pragma unroll 1
            for (; it < kMaxIter; ++it) {
                float* xv = s_vec + buf * 64;
                buf ^= 1;
                xv[lane] = x.x; xv[32 + lane] = x.y;
                __syncwarp();
                // y = C x on column pairs with packed FMAs: Re y = sum cr*xr - sum ci*xi,
                // Im y = sum cr*xi + sum ci*xr; 8 independent accumulator pairs
                u64 A0 = 0ull, A1 = 0ull, B0 = 0ull, B1 = 0ull, C0 = 0ull, C1 = 0ull, D0 = 0ull, D1 = 0ull;
                const ulonglong2* xr4 = reinterpret_cast<const ulonglong2*>(xv);
                const ulonglong2* xi4 = reinterpret_cast<const ulonglong2*>(xv + 32);
#pragma unroll
                for (int k = 0; k + 1 < NP2; k += 2) {
                    const ulonglong2 qr = xr4[k >> 1], qi = xi4[k >> 1];
                    A0 = fma2(cr2[k], qr.x, A0); B0 = fma2(ci2[k], qi.x, B0);
                    C0 = fma2(cr2[k], qi.x, C0); D0 = fma2(ci2[k], qr.x, D0);
                    A1 = fma2(cr2[k + 1], qr.y, A1); B1 = fma2(ci2[k + 1], qi.y, B1);
                    C1 = fma2(cr2[k + 1], qi.y, C1); D1 = fma2(ci2[k + 1], qr.y, D1);
                }
                if (NP2 & 1) {
                    const u64 qr = *reinterpret_cast<const u64*>(xv + NE - 2), qi = *reinterpret_cast<const u64*>(xv + 32 + NE - 2);
                    A0 = fma2(cr2[NP2 - 1], qr, A0); B0 = fma2(ci2[NP2 - 1], qi, B0);
                    C0 = fma2(cr2[NP2 - 1], qi, C0); D0 = fma2(ci2[NP2 - 1], qr, D0);
                }
                float yr, yi;
                {                                                  // fold the accumulator pairs with packed operations too
                    const u64 one2 = pack2(1.f, 1.f), neg2 = pack2(-1.f, -1.f);
                    const float2 re = unpack2(fma2(fma2(B1, one2, B0), neg2, fma2(A1, one2, A0)));
                    const float2 im = unpack2(fma2(fma2(D1, one2, D0), one2, fma2(C1, one2, C0)));
                    yr = re.x + re.y;
                    yi = im.x + im.y;
                }
                if (it < next_chk - 1) {
                    const float2 xn = make_float2(fmaf(-beta, xp.x, yr * inv_lam), fmaf(-beta, xp.y, yi * inv_lam));
                    xp = x; x = xn;
                } else if (it == next_chk - 1) {
                    // Rayleigh quotient one iteration ahead of the residual test, so that the two warp
                    // reductions are separate single rounds (the quotient's error is second order in the
                    // residual, one iteration of staleness is far below the tolerance); renormalise here
                    float xy = x.x * yr + x.y * yi, xx = x.x * x.x + x.y * x.y;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        xy += __shfl_xor_sync(FULLMASK, xy, s);
                        xx += __shfl_xor_sync(FULLMASK, xx, s);
                    }
                    const float ixx = fast_rcp(xx);
                    lam = xy * ixx;
                    inv_lam = fast_rcp(lam);
                    const float sc = fast_rsqrt(xx);
                    const float2 xn = make_float2(fmaf(-beta, xp.x, yr * inv_lam) * sc, fmaf(-beta, xp.y, yi * inv_lam) * sc);
                    xp = make_float2(x.x * sc, x.y * sc);
                    x = xn;
                } else {
                    // residual against the quotient of the previous iteration
                    const float rx = yr - lam * x.x, ry = yi - lam * x.y;
                    float rr2 = rx * rx + ry * ry, y2 = yr * yr + yi * yi, xx = x.x * x.x + x.y * x.y;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        rr2 += __shfl_xor_sync(FULLMASK, rr2, s);
                        y2 += __shfl_xor_sync(FULLMASK, y2, s);
                        xx += __shfl_xor_sync(FULLMASK, xx, s);
                    }
                    const float rho2 = rr2 * fast_rcp(lam * lam * xx);   // relative residual^2
                    // converged, or stalled on the FP32 rounding floor of the residual (~1e-13) just above
                    // the tolerance: further iterations cannot improve the vector
                    conv = (rho2 <= tol2) || (rho2 <= 4.f * tol2 && rho_prev > 0.f && rho2 > 0.8f * rho_prev);
                    if (conv) {                                          // final vector: one more plain step
                        const float sc = rsqrtf(y2);
                        x.x = yr * sc; x.y = yi * sc;
                        ++it; break;
                    }
                    // Next test where the residual is predicted to reach the tolerance (at least 2
                    // iterations ahead: the Rayleigh quotient needs the one before).  Observed
                    // per-iteration decay of the residual over the last gap; when the momentum term is
                    // switched on, its asymptotic rate r / (1 + sqrt(1 - r^2)) instead.
                    float rate_l2 = -0.15f;                               // log2 of the expected decay, unknown: ~0.9
                    if (rho_prev > 0.f && rho2 < rho_prev) {
                        const float l2 = __log2f(rho2 * fast_rcp(rho_prev)) * (0.5f / (float)gap);   // log2(rate) < 0
                        rate_l2 = l2;
                        if (beta == 0.f) {
                            // The decay seen this early still contains faster modes, i.e. it
                            // underestimates r = lambda2/lambda1, and heavy-ball momentum loses much more
                            // below its optimum beta = r^2/4 than above it: aim high (0.575 r instead of
                            // 0.5 r), stay below the stability limit 1/4, back off if the residual ever
                            // grows.  Offline replay on 300 coherence matrices of the bench stack: 14.8 ->
                            // 12.9 iterations (exact r from the start would give 11.4).
                            const float rr = exp2f(l2);
                            const float hb = 0.575f * rr;
                            beta = fminf(hb * hb, 0.2f);
                            rate_l2 = __log2f(rr * fast_rcp(1.f + sqrtf(fmaxf(1.f - rr * rr, 0.f))));
                        }
                    } else if (rho_prev > 0.f) {
                        beta *= 0.5f;
                    }
                    {
                        const float need = 0.5f * __log2f(tol2 * fast_rcp(rho2));     // log2 of the factor still missing (< 0)
                        const float m = 1.25f * need * fast_rcp(fminf(rate_l2, -0.01f)) + 0.5f;
                        gap = (rho_prev > 0.f) ? min(max((int)m, 2), 12) : 2;      // first interval: 2 (decay measured over iterations 3..5)
                        next_chk = it + gap;
                    }
                    rho_prev = rho2;
                    const float2 xn = make_float2(fmaf(-beta, xp.x, yr * inv_lam), fmaf(-beta, xp.y, yi * inv_lam));
                    xp = x; x = xn;
                }
            }
            PHASE_MARK(4)
            st_it += it;
            st_cap += conv ? 0 : 1;
            if (lam < 1.0e-6f) tc = -7.f;             // evd.cpp:723-727
            else {
                // ---------------- phase reference (evd.cpp:738-749) -----------------
                float* xv = s_vec + buf * 64;
                buf ^= 1;
                xv[lane] = x.x; xv[32 + lane] = x.y;
                __syncwarp();
                const float2 ref = make_float2(xv[k0], xv[32 + k0]);
                {
                    float ux = x.x * ref.x + x.y * ref.y, uy = x.y * ref.x - x.x * ref.y;
                    const float mm = ux * ux + uy * uy;
                    if (mm == 0.f) {                  // arg(0) = 0 in the reference
                        const float rr = rsqrtf(ref.x * ref.x + ref.y * ref.y);
                        ux = ref.x * rr; uy = -ref.y * rr;
                    } else { const float rr = fast_rsqrt(mm); ux *= rr; uy *= rr; }
                    if (lane == k0) { ux = 1.f; uy = 0.f; }
                    o = make_float2(ux * live, uy * live);
                }
                // ---------------- compressed SLC (evd.cpp:755-762) ------------------
                float cr = 0.f, ci = 0.f;
                if (lane < N && lane >= k0) {
                    cr = zc.x * o.x + zc.y * o.y;
                    ci = zc.y * o.x - zc.x * o.y;
                }
                // ---------------- temporal coherence (evd.cpp:770-786) --------------
                float* ov = s_vec + buf * 64;
                buf ^= 1;
                ov[lane] = o.x; ov[32 + lane] = o.y;
                __syncwarp();
                float wr = 0.f, wi = 0.f;
                const float2* ovr2 = reinterpret_cast<const float2*>(ov);
                const float2* ovi2 = reinterpret_cast<const float2*>(ov + 32);
#pragma unroll
                for (int k = 0; k < NP2; ++k) {
                    const float2 vr = unpack2(cr2[k]), vi = unpack2(ci2[k]);
                    const float2 qr = ovr2[k], qi = ovi2[k];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = 2 * k + h;
                        const float cx = h ? vr.y : vr.x, cy = h ? vi.y : vi.x;
                        const float ojx = h ? qr.y : qr.x, ojy = h ? qi.y : qi.x;
                        // pairs (lane, j) with j > lane, inside the matrix and (STBAS) the band;
                        // e = C_ij / |C_ij| (arg(0) = 0 as in the reference)
                        const bool use = (usemask >> j) & 1u;
                        const float m2 = fmaf(cx, cx, cy * cy);
                        const float rr = use ? fast_rsqrt(m2) : 0.f;
                        const float ex = (m2 > 0.f) ? cx * rr : (use ? 1.f : 0.f);
                        const float ey = (m2 > 0.f) ? cy * rr : 0.f;
                        wr = fmaf(ex, ojx, wr); wr = fmaf(-ey, ojy, wr);
                        wi = fmaf(ex, ojy, wi); wi = fmaf(ey, ojx, wi);
                    }
                }
                float sr = o.x * wr + o.y * wi, si = o.x * wi - o.y * wr;      // conj(o_r) * w_r
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    sr += __shfl_xor_sync(FULLMASK, sr, s);
                    si += __shfl_xor_sync(FULLMASK, si, s);
                    cr += __shfl_xor_sync(FULLMASK, cr, s);
                    ci += __shfl_xor_sync(FULLMASK, ci, s);
                }
                tc = sqrtf(sr * sr + si * si) * inv_pairs;
                const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
                cmp = make_float2(cr * invn, ci * invn);
            }
        }
        if (lane < N) a.out[(long)lane * npix_block + pg] = o;
        if (lane == 0) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
        __syncwarp();
        PHASE_MARK(5)
    }
    PHASE_FLUSH
    if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_pix);
        atomicAdd(&a.stats[1], st_it);
        atomicAdd(&a.stats[3], st_cap);
        atomicAdd(&a.stats[5], st_hot);
    }
}

template <int NE>
static cudaError_t launch_mma_t(const EvdArgs& a, cudaStream_t st) {
    typedef MmaCfg<NE> Cfg;
    const size_t lut = ((size_t)a.nulong * 32 * sizeof(short2) + 15) & ~(size_t)15;
    const size_t smem = lut + (size_t)Cfg::SMEM_PER_WARP * Cfg::WARPS;
    cudaError_t e = cudaFuncSetAttribute(k_evd_mma<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // bands of BAND rows x column segments; segments as long as possible (less halo re-reading at
    // their ends) while there are still >= mult CTAs per SM to even out the tail (one CTA per SM at a
    // time: 8 per SM lost 8 % to the last partial wave, 48 and 96 measured equal)
    const int mult = 48;
    const int nbands = (a.n_lines + Cfg::BAND - 1) / Cfg::BAND;
    int nseg = (nsm * mult + nbands - 1) / nbands;
    if (nseg < 1) nseg = 1;
    int seglen = (a.cols + nseg - 1) / nseg;
    if (seglen < 32) seglen = 32;
    nseg = (a.cols + seglen - 1) / seglen;
    EvdArgs b = a;
    b.tile_pairs = seglen;
    k_evd_mma<NE><<<(unsigned)(nbands * nseg), Cfg::WARPS * 32, smem, st>>>(b);
    return cudaGetLastError();
}

// eigen-phase order: the smallest instantiated even order >= bands
int evd_mma_order(int bands) {
    static const int orders[] = {8, 12, 16, 20, 24, 28, 30, 32};
    if (bands < 2) return 0;
    for (int o : orders) if (bands <= o) return o;
    return 0;
}

cudaError_t launch_evd_mma(const EvdArgs& a, cudaStream_t st) {
    switch (evd_mma_order(a.bands)) {
        case 8: return launch_mma_t<8>(a, st);
        case 12: return launch_mma_t<12>(a, st);
        case 16: return launch_mma_t<16>(a, st);
        case 20: return launch_mma_t<20>(a, st);
        case 24: return launch_mma_t<24>(a, st);
        case 28: return launch_mma_t<28>(a, st);
        case 30: return launch_mma_t<30>(a, st);
        case 32: return launch_mma_t<32>(a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fringe
