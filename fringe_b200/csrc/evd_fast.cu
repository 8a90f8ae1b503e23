// Register-blocked covariance + in-register power iteration (sm_100a), bands <= 30.
//
// This is the kernel the headline benchmark runs (EVD / STBAS, evd.cpp control flow).  Same
// arithmetic contract as the generic kernel in evd_kernels.cu, different mapping:
//
//   covariance  A warp owns two horizontally adjacent pixels.  The upper triangle of each
//               pixel's N x N Hermitian accumulator is cut into a 5 x 5 grid of B x B blocks
//               (B = ceil(N/5) <= 6); the 15 blocks on or above the diagonal go to 15 lanes
//               (lanes 0-14 pixel 0, 16-30 pixel 1).  Per SHP a lane loads the 2B samples its
//               block needs (16-byte loads from the pixel-major stack, served by L1: adjacent
//               pixels share almost all of their window) and issues B*B complex FMAs, i.e.
//               3 complex MACs per loaded sample instead of 0.5 in the generic kernel.  The
//               SHP loop walks the set bits of the window mask (ffs / clear-lowest), so unset
//               neighbours cost nothing.
//   eigen       The normalised coherence is parked in shared memory as a packed triangle
//               (3.7 kB per pixel at N=30) and re-read row-per-lane into registers; the power
//               iteration then needs only the broadcast vector from shared memory (15
//               LDS.128 per 120 FMAs).  Residual / normalisation reductions run every 4th
//               iteration.
//   epilogue    phase reference, compressed SLC and temporal coherence from registers.
#include <math_constants.h>

#include <cstdio>

#include "common.cuh"

namespace fringe {

#define FULLMASK 0xffffffffu
#ifdef FRINGE_DEBUG_TRACE
#define TRACE(...) do { if ((lane & 15) == 0 && blockIdx.x == 0) printf(__VA_ARGS__); } while (0)
#else
#define TRACE(...) do {} while (0)
#endif

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}

template <int B>
struct FastCfg {
    static constexpr int NB = 5;
    static constexpr int NPAD = NB * B;                 // padded matrix order (zpix row length)
    static constexpr int TRI = NPAD * (NPAD + 1) / 2;   // packed upper triangle incl. diagonal
    static constexpr int WARPS = 4;
    // per warp: two packed triangles, two broadcast vectors (double buffered), powers
    static constexpr int SMEM_PER_WARP =
        (((2 * TRI + 2 * 32) * (int)sizeof(float2) + 2 * NPAD * (int)sizeof(float)) + 15) & ~15;
};

__device__ __forceinline__ int tri_index(int i, int j, int n) {   // i <= j
    return i * n - ((i * (i - 1)) >> 1) + (j - i);
}

template <int B>
__global__ void __launch_bounds__(128, 3) k_evd_fast(const EvdArgs a) {
    typedef FastCfg<B> Cfg;
    constexpr int NPAD = Cfg::NPAD;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 4, l = lane & 15;
    const int N = a.bands;                      // <= NPAD, zpix rows are zero padded to NPAD

    unsigned char* base = s_raw + (size_t)warp * Cfg::SMEM_PER_WARP;
    float2* s_tri = reinterpret_cast<float2*>(base);                       // [2][TRI]
    float2* s_vec = s_tri + 2 * Cfg::TRI;                                  // [2][32]
    float* s_pw = reinterpret_cast<float*>(s_vec + 64);                    // [2][NPAD]

    // block coordinates of this lane inside its pixel group
    int bi = 0, bj = 0;
    {
        int k = l;
        bi = 0;
        int rowlen = 5;
        while (bi < 4 && k >= rowlen) { k -= rowlen; ++bi; --rowlen; }
        bj = bi + k;
    }
    const bool blk_active = (l < 15);

    const int WX = 2 * a.Nx + 1, center = a.Ny * WX + a.Nx;
    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2);
    const int BW = a.bandwidth;
    const long npix_block = (long)a.cols * a.lines;

    const int pairs_per_row = (a.cols + 1) >> 1;
    const long total_pairs = (long)a.n_lines * pairs_per_row;
    const long chunk = (total_pairs + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * chunk;
    const long end = min(total_pairs, beg + chunk);
    unsigned long long st_pix = 0, st_it = 0, st_cap = 0;

    for (long pr = beg + warp; pr < end; pr += Cfg::WARPS) {
        const int row = a.first_line + (int)(pr / pairs_per_row);
        const int col0 = 2 * (int)(pr % pairs_per_row);
        const int mycol = col0 + grp;
        const bool pix_exists = mycol < a.cols;
        const long p = (long)row * a.cols + mycol;

        // ------------------------- covariance (evd.cpp:537-564) -------------------------
        float2 acc[B][B];
#pragma unroll
        for (int i = 0; i < B; ++i)
#pragma unroll
            for (int j = 0; j < B; ++j) acc[i][j] = make_float2(0.f, 0.f);
        int npix = 0;
        bool center_on = false;
        if (pix_exists) center_on = (__ldg(&a.wts[p * a.nulong + (center >> 5)]) >> (center & 31)) & 1u;
        for (int w = 0; w < a.nulong; ++w) {
            uint32_t m = (pix_exists && center_on && blk_active) ? __ldg(&a.wts[p * a.nulong + w]) : 0u;
            // uniform trip count: the longer of the two pixels' bit lists in this word
            const int trips = __reduce_max_sync(FULLMASK, __popc(m));
#pragma unroll 1
            for (int t = 0; t < trips; ++t) {
                const bool on = (m != 0u);
                const int f = w * 32 + (on ? (__ffs(m) - 1) : 0);
                m &= (m - 1u);
                const int fy = f / WX;
                const int yy = row + fy - a.Ny, xx = mycol + (f - fy * WX) - a.Nx;
                const bool inb = on && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                npix += inb ? 1 : 0;
                float2 za[B], zb[B];
                if (inb) {
                    const float2* zq = a.zpix + ((long)yy * a.cols + xx) * NPAD;
                    if (B % 2 == 0) {
                        const float4* pa = reinterpret_cast<const float4*>(zq + B * bi);
                        const float4* pb = reinterpret_cast<const float4*>(zq + B * bj);
#pragma unroll
                        for (int i = 0; i < B / 2; ++i) {
                            const float4 va = __ldg(pa + i), vb = __ldg(pb + i);
                            za[2 * i] = make_float2(va.x, va.y); za[2 * i + 1] = make_float2(va.z, va.w);
                            zb[2 * i] = make_float2(vb.x, vb.y); zb[2 * i + 1] = make_float2(vb.z, vb.w);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < B; ++i) { za[i] = __ldg(zq + B * bi + i); zb[i] = __ldg(zq + B * bj + i); }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < B; ++i) { za[i] = make_float2(0.f, 0.f); zb[i] = make_float2(0.f, 0.f); }
                }
#pragma unroll
                for (int i = 0; i < B; ++i)
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        acc[i][j].x = fmaf(za[i].x, zb[j].x, acc[i][j].x);
                        acc[i][j].x = fmaf(za[i].y, zb[j].y, acc[i][j].x);
                        acc[i][j].y = fmaf(za[i].y, zb[j].x, acc[i][j].y);
                        acc[i][j].y = fmaf(-za[i].x, zb[j].y, acc[i][j].y);
                    }
            }
        }
#ifdef FRINGE_DEBUG_TRACE
        if ((lane & 15) == 0) printf("b%d w%d l%d pair %ld row %d col0 %d npix %d center %d\n", blockIdx.x, warp, lane, pr, row, col0, npix, (int)center_on);
#endif
        // group-uniform: enough SHPs?  (evd.cpp:566 hard-codes 2)
        const int npix_grp = __shfl_sync(FULLMASK, npix, grp << 4);   // collective first: no short-circuit around it
        const bool solve_me = pix_exists && center_on && (npix_grp >= 2);

        // ------------------------- coherence (evd.cpp:569-582) --------------------------
        __syncwarp();
        if (blk_active && bi == bj) {
#pragma unroll
            for (int i = 0; i < B; ++i) s_pw[grp * NPAD + B * bi + i] = sqrtf(acc[i][i].x);
        }
        __syncwarp();
        if (blk_active) {
            float pa[B], pb[B];
#pragma unroll
            for (int i = 0; i < B; ++i) { pa[i] = s_pw[grp * NPAD + B * bi + i]; pb[i] = s_pw[grp * NPAD + B * bj + i]; }
            float2* tri = s_tri + grp * Cfg::TRI;
#pragma unroll
            for (int i = 0; i < B; ++i)
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const int gi = B * bi + i, gj = B * bj + j;
                    if (gi < gj && gj < N) {
                        const float inv = 1.0f / (pa[i] * pb[j]);
                        tri[tri_index(gi, gj, N)] = make_float2(acc[i][j].x * inv, acc[i][j].y * inv);
                    }
                }
        }
        __syncwarp();

        // ------------------------- per pixel: eigen + epilogue --------------------------
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const bool exists_g = (col0 + g) < a.cols;
            if (!exists_g) continue;
            const bool solve_g = __shfl_sync(FULLMASK, solve_me ? 1 : 0, g << 4) != 0;
            const long pg = (long)row * a.cols + col0 + g;
            float2 o = make_float2(0.f, 0.f);
            float tc = 0.f;
            float2 cmp = make_float2(0.f, 0.f);
            TRACE("w%d l%d g%d solve %d\n", warp, lane, g, (int)solve_g);
            if (solve_g) {
                ++st_pix;
                const float2* tri = s_tri + g * Cfg::TRI;
                const int r = lane;
                float2 c[NPAD];
#pragma unroll
                for (int j = 0; j < NPAD; ++j) {
                    float2 v = make_float2(0.f, 0.f);
                    if (r < N && j < N) {
                        if (j == r) v = make_float2(1.f, 0.f);
                        else if (j > r) v = tri[tri_index(r, j, N)];
                        else { v = tri[tri_index(j, r, N)]; v.y = -v.y; }
                        if (isstbas && abs(j - r) > BW) v = make_float2(0.f, 0.f);
                    }
                    c[j] = v;
                }
                // start vector: column k0 of C
                float2 x = make_float2(0.f, 0.f);
                if (r < N) {
                    if (r == k0) x = make_float2(1.f, 0.f);
                    else if (r < k0) x = tri[tri_index(r, k0, N)];
                    else { x = tri[tri_index(k0, r, N)]; x.y = -x.y; }
                    if (isstbas && abs(k0 - r) > BW) x = make_float2(0.f, 0.f);
                }
                {
                    const float n2 = wsum(x.x * x.x + x.y * x.y);
                    const float sc = rsqrtf(n2);
                    x.x *= sc; x.y *= sc;
                }
                float lam = 1.f, inv_lam = 1.f;
                int it = 0, buf = 0;
                bool conv = false;
                const int kMaxIter = 1000;
                const float tol2 = 4.0e-12f;
                while (it < kMaxIter && !conv) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float2* xv = s_vec + buf * 32;
                        buf ^= 1;
                        xv[lane] = x;
                        __syncwarp();
                        float yr = 0.f, yi = 0.f;
                        const float4* xv4 = reinterpret_cast<const float4*>(xv);
#pragma unroll
                        for (int j = 0; j < NPAD; j += 2) {
                            if (j + 1 < NPAD) {
                                const float4 q = xv4[j >> 1];
                                yr = fmaf(c[j].x, q.x, yr); yr = fmaf(-c[j].y, q.y, yr);
                                yi = fmaf(c[j].x, q.y, yi); yi = fmaf(c[j].y, q.x, yi);
                                yr = fmaf(c[j + 1].x, q.z, yr); yr = fmaf(-c[j + 1].y, q.w, yr);
                                yi = fmaf(c[j + 1].x, q.w, yi); yi = fmaf(c[j + 1].y, q.z, yi);
                            } else {
                                const float2 q = xv[j];
                                yr = fmaf(c[j].x, q.x, yr); yr = fmaf(-c[j].y, q.y, yr);
                                yi = fmaf(c[j].x, q.y, yi); yi = fmaf(c[j].y, q.x, yi);
                            }
                        }
                        ++it;
                        if (u < 3) { x.x = yr * inv_lam; x.y = yi * inv_lam; }
                        else {
                            // Rayleigh quotient, residual, renormalisation
                            float xy = x.x * yr + x.y * yi, xx = x.x * x.x + x.y * x.y;
#pragma unroll
                            for (int s = 16; s > 0; s >>= 1) {
                                xy += __shfl_xor_sync(FULLMASK, xy, s);
                                xx += __shfl_xor_sync(FULLMASK, xx, s);
                            }
                            lam = xy / xx;
                            const float rx = yr - lam * x.x, ry = yi - lam * x.y;
                            float r2 = rx * rx + ry * ry, y2 = yr * yr + yi * yi;
#pragma unroll
                            for (int s = 16; s > 0; s >>= 1) {
                                r2 += __shfl_xor_sync(FULLMASK, r2, s);
                                y2 += __shfl_xor_sync(FULLMASK, y2, s);
                            }
                            const float sc = rsqrtf(y2);
                            x.x = yr * sc; x.y = yi * sc;
                            inv_lam = 1.0f / lam;
                            conv = (r2 <= tol2 * lam * lam * xx);
                            TRACE("w%d l%d g%d it %d lam %g r2 %g xx %g conv %d\n", warp, lane, g, it, lam, r2, xx, (int)conv);
                        }
                    }
                }
                st_it += it;
                st_cap += conv ? 0 : 1;
                if (lam < 1.0e-6f) tc = -7.f;             // evd.cpp:723-727
                else {
                    // ---------------- phase reference (evd.cpp:738-749) -----------------
                    float2* xv = s_vec + buf * 32;
                    buf ^= 1;
                    xv[lane] = x;
                    __syncwarp();
                    const float2 ref = xv[k0];
                    if (r < N) {
                        float ux = x.x * ref.x + x.y * ref.y, uy = x.y * ref.x - x.x * ref.y;
                        const float mm = ux * ux + uy * uy;
                        if (mm == 0.f) {
                            const float rr = rsqrtf(ref.x * ref.x + ref.y * ref.y);
                            ux = ref.x * rr; uy = -ref.y * rr;
                        } else { const float rr = rsqrtf(mm); ux *= rr; uy *= rr; }
                        if (r == k0) { ux = 1.f; uy = 0.f; }
                        o = make_float2(ux, uy);
                    }
                    // ---------------- compressed SLC (evd.cpp:755-762) ------------------
                    float cr = 0.f, ci = 0.f;
                    if (r < N && r >= k0) {
                        const float2 z = __ldg(&a.zpix[pg * NPAD + r]);
                        cr = z.x * o.x + z.y * o.y;
                        ci = z.y * o.x - z.x * o.y;
                    }
                    // ---------------- temporal coherence (evd.cpp:770-786) --------------
                    float2* ov = s_vec + buf * 32;
                    buf ^= 1;
                    ov[lane] = o;
                    __syncwarp();
                    float wr = 0.f, wi = 0.f;
                    int cnt = 0;
#pragma unroll
                    for (int j = 0; j < NPAD; ++j) {
                        const bool use = (j > r) && (j < N) && (r < N) && (!isstbas || (j - r) <= BW);
                        if (use) {
                            // c[j] may have been zeroed by the STBAS band limit only outside the band
                            const float m2 = c[j].x * c[j].x + c[j].y * c[j].y;
                            float ex = 1.f, ey = 0.f;
                            if (m2 > 0.f) { const float rr = rsqrtf(m2); ex = c[j].x * rr; ey = c[j].y * rr; }
                            const float2 oj = ov[j];
                            wr += ex * oj.x - ey * oj.y;
                            wi += ex * oj.y + ey * oj.x;
                            ++cnt;
                        }
                    }
                    // conj(o_r) * w_r
                    float sr = o.x * wr + o.y * wi, si = o.x * wi - o.y * wr;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        sr += __shfl_xor_sync(FULLMASK, sr, s);
                        si += __shfl_xor_sync(FULLMASK, si, s);
                        cr += __shfl_xor_sync(FULLMASK, cr, s);
                        ci += __shfl_xor_sync(FULLMASK, ci, s);
                    }
                    cnt = __reduce_add_sync(FULLMASK, cnt);
                    tc = sqrtf(sr * sr + si * si) / (float)cnt;
                    const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
                    cmp = make_float2(cr * invn, ci * invn);
                }
            }
            if (lane < N) a.out[(long)lane * npix_block + pg] = o;
            if (lane == 0) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
            __syncwarp();
        }
    }
#ifdef FRINGE_DEBUG_TRACE
    if (lane == 0) printf("b%d w%d done (pairs %ld..%ld)\n", blockIdx.x, warp, beg, end);
#endif
    if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_pix);
        atomicAdd(&a.stats[1], st_it);
        atomicAdd(&a.stats[3], st_cap);
    }
}

template <int B>
static cudaError_t launch_fast_t(const EvdArgs& a, cudaStream_t st) {
    typedef FastCfg<B> Cfg;
    const size_t smem = (size_t)Cfg::SMEM_PER_WARP * Cfg::WARPS;
    cudaError_t e = cudaFuncSetAttribute(k_evd_fast<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_evd_fast<B>, Cfg::WARPS * 32, smem);
    if (occ < 1) occ = 1;
    const long total_pairs = (long)a.n_lines * ((a.cols + 1) / 2);
    long grid = (long)nsm * occ * 8;
    const long maxgrid = (total_pairs + Cfg::WARPS * 4 - 1) / (Cfg::WARPS * 4);
    if (grid > maxgrid) grid = maxgrid;
    if (grid < 1) grid = 1;
    k_evd_fast<B><<<(unsigned)grid, Cfg::WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

int evd_fast_padded_bands(int bands) {
    if (bands < 2 || bands > 30) return 0;
    const int B = (bands + 4) / 5;
    return 5 * (B < 2 ? 2 : B);
}

bool evd_fast_supported(const EvdArgs& a) {
    return a.variant == 0 && (a.method == 0 || a.method == 2) && evd_fast_padded_bands(a.bands) > 0 &&
           a.NP == evd_fast_padded_bands(a.bands);
}

cudaError_t launch_evd_fast(const EvdArgs& a, cudaStream_t st) {
    switch (a.NP / 5) {
        case 2: return launch_fast_t<2>(a, st);
        case 3: return launch_fast_t<3>(a, st);
        case 4: return launch_fast_t<4>(a, st);
        case 5: return launch_fast_t<5>(a, st);
        case 6: return launch_fast_t<6>(a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fringe
