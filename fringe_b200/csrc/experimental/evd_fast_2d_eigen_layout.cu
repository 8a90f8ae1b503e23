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
//               3 complex MACs per loaded sample.  The SHP loop walks the set bits of the
//               window mask (ffs / clear-lowest) with the next SHP's samples prefetched while
//               the current one is accumulated; exhausted or out-of-block bits point at an
//               all-zero sample row, so the loop body is branch-free.
//   eigen       The normalised coherence matrices of both pixels are parked in shared memory
//               (full Hermitian) and re-read as 8 x 4 register sub-blocks on a 4 x 8 lane grid;
//               a matrix-vector product is 128 FMAs + 22 shuffles per lane and touches no
//               shared memory.  Power iteration with heavy-ball momentum; residual /
//               normalisation reductions run every 4th iteration.
//   epilogue    phase reference, compressed SLC and temporal coherence from registers.
//
// Code size matters here: the first version unrolled everything and reached 92 kB of SASS,
// i.e. 3x the instruction cache, and ncu showed "no instruction" as the top stall.  The hot
// loops below are rolled (#pragma unroll 1) around fully unrolled bodies.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace fringe {

#define FULLMASK 0xffffffffu

// single-instruction MUFU approximations (about 1 ulp); arguments here are sums of squares of
// O(1) quantities, far from the denormal range the IEEE-exact expansions would guard
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

template <int B>
struct FastCfg {
    static constexpr int NB = 5;
    static constexpr int NPAD = NB * B;                 // padded matrix order (zpix row length)
    static constexpr int WARPS = 4;
    // per warp: two full matrices, two broadcast vectors (double buffered), powers
    static constexpr int MAT = NPAD * NPAD;             // float2 elements per matrix
    static constexpr int SMEM_PER_WARP =
        (((2 * MAT + 2 * 32) * (int)sizeof(float2) + 2 * NPAD * (int)sizeof(float)) + 15) & ~15;
};

// One SHP's operands for a lane's block: rows B*bi.. and columns B*bj.. of the sample vector.
template <int B>
struct Operands {
    float2 a[B], b[B];
    __device__ __forceinline__ void load(const float2* __restrict__ zq, int oa, int ob) {
        if (B % 2 == 0) {
            const float4* pa = reinterpret_cast<const float4*>(zq + oa);
            const float4* pb = reinterpret_cast<const float4*>(zq + ob);
#pragma unroll
            for (int i = 0; i < B / 2; ++i) {
                const float4 va = __ldg(pa + i), vb = __ldg(pb + i);
                a[2 * i] = make_float2(va.x, va.y); a[2 * i + 1] = make_float2(va.z, va.w);
                b[2 * i] = make_float2(vb.x, vb.y); b[2 * i + 1] = make_float2(vb.z, vb.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < B; ++i) { a[i] = __ldg(zq + oa + i); b[i] = __ldg(zq + ob + i); }
        }
    }
};

template <int B>
__device__ __forceinline__ void accumulate(float2 (&acc)[B][B], const Operands<B>& op) {
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j) {      // acc += a_i * conj(b_j)
            acc[i][j].x = fmaf(op.a[i].x, op.b[j].x, acc[i][j].x);
            acc[i][j].x = fmaf(op.a[i].y, op.b[j].y, acc[i][j].x);
            acc[i][j].y = fmaf(op.a[i].y, op.b[j].x, acc[i][j].y);
            acc[i][j].y = fmaf(-op.a[i].x, op.b[j].y, acc[i][j].y);
        }
}

template <int B>
__global__ void __launch_bounds__(128, 3) k_evd_fast(const EvdArgs a) {
    typedef FastCfg<B> Cfg;
    constexpr int NPAD = Cfg::NPAD;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 4, l = lane & 15;
    const int N = a.bands;                      // <= NPAD, zpix rows are zero padded to NPAD

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
    float2* s_mat = reinterpret_cast<float2*>(base);                       // [2][NPAD][NPAD]
    float2* s_vec = s_mat + 2 * Cfg::MAT;                                  // [2][32]
    float* s_pw = reinterpret_cast<float*>(s_vec + 64);                    // [2][NPAD]

    // block coordinates of this lane inside its pixel group
    int bi = 0, bj = 0;
    {
        int k = l, rowlen = 5;
        while (bi < 4 && k >= rowlen) { k -= rowlen; ++bi; --rowlen; }
        bj = bi + k;
    }
    const bool blk_active = (l < 15);
    const int oa = B * bi, ob = B * bj;

    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2);
    const int BW = a.bandwidth;
    // 2-D eigen layout: lane (ri, cj) owns rows {ri + 4k} x columns {4cj + t}; result row ri + 4cj
    const int ri = lane >> 3, cj = lane & 7;
    const int myrow = ri + 4 * cj;
    // bit k*4+t set: matrix entry (ri+4k, 4cj+t) enters the temporal-coherence sum (row < column,
    // inside the matrix and, for STBAS, inside the band)
    uint32_t usemask = 0u;
    for (int k = 0; k < 8; ++k)
        for (int t = 0; t < 4; ++t) {
            const int rr = ri + 4 * k, cc = 4 * cj + t;
            if (rr < cc && cc < N && (!isstbas || (cc - rr) <= BW)) usemask |= (1u << (k * 4 + t));
        }
    // number of (i<j) pairs in the temporal-coherence sum (evd.cpp:773-784)
    float inv_pairs;
    {
        int cnt = 0;
        for (int i = 0; i < N; ++i) cnt += isstbas ? min(BW, N - 1 - i) : (N - 1 - i);
        inv_pairs = 1.0f / (float)cnt;
    }
    const long npix_block = (long)a.cols * a.lines;
    const float2* zero_row = a.zpix + npix_block * NPAD;    // one all-zero sample vector

    const int pairs_per_row = (a.cols + 1) >> 1;
    const long total_pairs = (long)a.n_lines * pairs_per_row;
    const long chunk = (total_pairs + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * chunk;
    const long end = min(total_pairs, beg + chunk);
    unsigned long long st_pix = 0, st_it = 0, st_cap = 0;

#pragma unroll 1
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
#pragma unroll 1
        for (int w = 0; w < a.nulong; ++w) {
            uint32_t m = (pix_exists && center_on && blk_active) ? __ldg(&a.wts[p * a.nulong + w]) : 0u;
            // uniform trip count: the longer of the two pixels' bit lists in this word
            const int trips = __reduce_max_sync(FULLMASK, __popc(m));
            // address of the next SHP's sample vector (zero row when exhausted / outside block)
            auto next_ptr = [&]() -> const float2* {
                const bool on = (m != 0u);
                const int f = w * 32 + (on ? (__ffs(m) - 1) : 0);
                m &= (m - 1u);
                const short2 d = s_off[f];
                const int yy = row + d.x, xx = mycol + d.y;
                const bool inb = on && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                npix += inb ? 1 : 0;
                return inb ? a.zpix + ((long)yy * a.cols + xx) * NPAD : zero_row;
            };
            // two-stage software pipeline: while one SHP is accumulated the next one's samples
            // are already in flight (exhausted lists read the zero row, so no tail handling)
            Operands<B> opA, opB;
            if (trips > 0) opA.load(next_ptr(), oa, ob);
#pragma unroll 1
            for (int t = 0; t < trips; t += 2) {
                opB.load(next_ptr(), oa, ob);
                accumulate<B>(acc, opA);
                opA.load(next_ptr(), oa, ob);
                accumulate<B>(acc, opB);
            }
        }
        // group-uniform: enough SHPs?  (evd.cpp:566 hard-codes 2).  Collective first: no
        // short-circuit evaluation around a warp shuffle.
        const int npix_grp = __shfl_sync(FULLMASK, npix, grp << 4);
        const bool solve_me = pix_exists && center_on && (npix_grp >= 2);

        // ------------------------- coherence (evd.cpp:569-582) --------------------------
        __syncwarp();
        if (blk_active && bi == bj) {
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int t = oa + i;
                // padded bands get +inf so that their scaled entries come out as exact zeros
                s_pw[grp * NPAD + t] = (t < N) ? sqrtf(acc[i][i].x) : CUDART_INF_F;
            }
        }
        __syncwarp();
        if (blk_active) {
            float ia[B], ib[B];
#pragma unroll
            for (int i = 0; i < B; ++i) { ia[i] = fast_rcp(s_pw[grp * NPAD + oa + i]); ib[i] = fast_rcp(s_pw[grp * NPAD + ob + i]); }
            float2* mat = s_mat + grp * Cfg::MAT;
            float2* up = mat + oa * NPAD + ob;            // block (bi,bj)
            float2* lo = mat + ob * NPAD + oa;            // its mirror
            const bool diag = (bi == bj);
#pragma unroll
            for (int i = 0; i < B; ++i)
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const float s = ia[i] * ib[j];
                    const float2 c = make_float2(acc[i][j].x * s, acc[i][j].y * s);
                    if (!diag || j > i) {                 // off-diagonal entry and its mirror
                        up[i * NPAD + j] = c;
                        lo[j * NPAD + i] = make_float2(c.x, -c.y);
                    }
                    if (j == i && diag) up[i * NPAD + i] = make_float2((oa + i < N) ? 1.f : 0.f, 0.f);
                }
        }
        __syncwarp();

        // ------------------------- per pixel: eigen + epilogue --------------------------
        // 2-D register layout for the matrix-vector products: lane (ri, cj) = (lane>>3, lane&7)
        // holds the 8 x 4 sub-block  rows {ri + 4k}, k<8  x  columns {4cj + t}, t<4  of the padded
        // 32 x 32 matrix.  A product then needs only the 4 vector entries of the lane's columns
        // (8 shuffles) and a reduce-scatter of the 8 partial row sums over the 8 lanes sharing ri
        // (14 shuffles); the result row of lane (ri, cj) is ri + 4cj.  The earlier lane-per-row
        // layout broadcast the whole vector from shared memory to every lane -- 64 shared-memory
        // wavefronts per product -- and made this phase shared-memory-bandwidth bound.
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const bool exists_g = (col0 + g) < a.cols;
            if (!exists_g) continue;
            const bool solve_g = __shfl_sync(FULLMASK, solve_me ? 1 : 0, g << 4) != 0;
            const long pg = (long)row * a.cols + col0 + g;
            float2 o = make_float2(0.f, 0.f);
            float tc = 0.f;
            float2 cmp = make_float2(0.f, 0.f);
            if (solve_g) {
                ++st_pix;
                const float2* mat = s_mat + g * Cfg::MAT;
                float2 c[8][4];
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int rr = ri + 4 * k, cc = 4 * cj + t;
                        const bool in = (rr < N) && (cc < N) && (!isstbas || abs(rr - cc) <= BW);
                        c[k][t] = in ? mat[rr * NPAD + cc] : make_float2(0.f, 0.f);
                    }
                // start vector: column k0 of C (= conj of row k0), entry of my own row
                float2 x;
                {
                    const bool in = (myrow < N) && (!isstbas || abs(myrow - k0) <= BW);
                    const float2 v = in ? mat[k0 * NPAD + myrow] : make_float2(0.f, 0.f);
                    x = make_float2(v.x, -v.y);
                    float n2 = x.x * x.x + x.y * x.y;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(FULLMASK, n2, s);
                    const float sc = rsqrtf(n2);
                    x.x *= sc; x.y *= sc;
                }
                // Power iteration with heavy-ball momentum: x+ = C x / lambda - beta * x-.  beta = 0
                // (plain power iteration) until the decay rate r ~ lambda2/lambda1 of the residual
                // has been observed over one check interval; then beta = (0.95 r / 2)^2, the
                // Chebyshev-optimal value if r is exact and a weaker acceleration if r is
                // underestimated (never a divergence: beta < 1/4).
                float lam = 1.f, inv_lam = 1.f, beta = 0.f, rho_prev = -1.f;
                float2 xp = make_float2(0.f, 0.f);
                int it = 0;
                bool conv = false;
                const int kMaxIter = (a.force_generic >> 1) ? (a.force_generic >> 1) : 1000;   // upper bits: timing experiment only
                const float tol2 = 4.0e-12f;
#pragma unroll 1
                for (; it < kMaxIter; ++it) {
                    // my four column entries live in lanes (t, cj)
                    float2 xc[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        xc[t].x = __shfl_sync(FULLMASK, x.x, t * 8 + cj);
                        xc[t].y = __shfl_sync(FULLMASK, x.y, t * 8 + cj);
                    }
                    float2 y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float yr = 0.f, yi = 0.f;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            yr = fmaf(c[k][t].x, xc[t].x, yr); yr = fmaf(-c[k][t].y, xc[t].y, yr);
                            yi = fmaf(c[k][t].x, xc[t].y, yi); yi = fmaf(c[k][t].y, xc[t].x, yi);
                        }
                        y[k] = make_float2(yr, yi);
                    }
                    // reduce-scatter over cj: after the three steps y[0] is the full sum of row ri + 4cj
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool up = (cj & 4) != 0;
                        const float2 keep = up ? y[k + 4] : y[k], send = up ? y[k] : y[k + 4];
                        y[k].x = keep.x + __shfl_xor_sync(FULLMASK, send.x, 4);
                        y[k].y = keep.y + __shfl_xor_sync(FULLMASK, send.y, 4);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const bool up = (cj & 2) != 0;
                        const float2 keep = up ? y[k + 2] : y[k], send = up ? y[k] : y[k + 2];
                        y[k].x = keep.x + __shfl_xor_sync(FULLMASK, send.x, 2);
                        y[k].y = keep.y + __shfl_xor_sync(FULLMASK, send.y, 2);
                    }
                    {
                        const bool up = (cj & 1) != 0;
                        const float2 keep = up ? y[1] : y[0], send = up ? y[0] : y[1];
                        y[0].x = keep.x + __shfl_xor_sync(FULLMASK, send.x, 1);
                        y[0].y = keep.y + __shfl_xor_sync(FULLMASK, send.y, 1);
                    }
                    const float yr = y[0].x, yi = y[0].y;
                    if ((it & 3) != 3) {
                        const float2 xn = make_float2(fmaf(-beta, xp.x, yr * inv_lam), fmaf(-beta, xp.y, yi * inv_lam));
                        xp = x; x = xn;
                    } else {
                        // Rayleigh quotient, residual, renormalisation
                        float xy = x.x * yr + x.y * yi, xx = x.x * x.x + x.y * x.y;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) {
                            xy += __shfl_xor_sync(FULLMASK, xy, s);
                            xx += __shfl_xor_sync(FULLMASK, xx, s);
                        }
                        lam = xy / xx;
                        const float rx = yr - lam * x.x, ry = yi - lam * x.y;
                        float rr2 = rx * rx + ry * ry, y2 = yr * yr + yi * yi;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) {
                            rr2 += __shfl_xor_sync(FULLMASK, rr2, s);
                            y2 += __shfl_xor_sync(FULLMASK, y2, s);
                        }
                        inv_lam = 1.0f / lam;
                        const float rho2 = rr2 / (lam * lam * xx);           // relative residual^2
                        conv = (rho2 <= tol2);
                        if (conv) {                                          // final vector: one more plain step
                            const float sc = rsqrtf(y2);
                            x.x = yr * sc; x.y = yi * sc;
                            ++it; break;
                        }
                        if (beta == 0.f && rho_prev > 0.f && rho2 < rho_prev) {
                            const float r = sqrtf(sqrtf(sqrtf(rho2 / rho_prev)));   // (rho2 ratio)^(1/8): per-iteration rate
                            const float hb = 0.475f * r;
                            beta = hb * hb;
                        }
                        rho_prev = rho2;
                        // momentum step, then put (x, x-) back on a unit scale
                        const float sc = rsqrtf(xx);
                        const float2 xn = make_float2(fmaf(-beta, xp.x, yr * inv_lam) * sc, fmaf(-beta, xp.y, yi * inv_lam) * sc);
                        xp = make_float2(x.x * sc, x.y * sc);
                        x = xn;
                    }
                }
                st_it += it;
                st_cap += conv ? 0 : 1;
                if (lam < 1.0e-6f) tc = -7.f;             // evd.cpp:723-727
                else {
                    // ---------------- phase reference (evd.cpp:738-749) -----------------
                    const int ref_lane = (k0 & 3) * 8 + (k0 >> 2);            // owner of row k0
                    const float2 ref = make_float2(__shfl_sync(FULLMASK, x.x, ref_lane), __shfl_sync(FULLMASK, x.y, ref_lane));
                    {
                        float ux = x.x * ref.x + x.y * ref.y, uy = x.y * ref.x - x.x * ref.y;
                        const float mm = ux * ux + uy * uy;
                        if (mm == 0.f) {                  // arg(0) = 0 in the reference
                            const float rr = rsqrtf(ref.x * ref.x + ref.y * ref.y);
                            ux = ref.x * rr; uy = -ref.y * rr;
                        } else { const float rr = fast_rsqrt(mm); ux *= rr; uy *= rr; }
                        if (myrow == k0) { ux = 1.f; uy = 0.f; }
                        o = (myrow < N) ? make_float2(ux, uy) : make_float2(0.f, 0.f);
                    }
                    // ---------------- compressed SLC (evd.cpp:755-762) ------------------
                    float cr = 0.f, ci = 0.f;
                    if (myrow < N && myrow >= k0) {
                        const float2 z = __ldg(&a.zpix[pg * NPAD + myrow]);
                        cr = z.x * o.x + z.y * o.y;
                        ci = z.y * o.x - z.x * o.y;
                    }
                    // ---------------- temporal coherence (evd.cpp:770-786) --------------
                    // sum over my sub-block of  e_ij * conj(o_i) * o_j,  e = C_ij/|C_ij| (arg(0) = 0)
                    float2 oc[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        oc[t].x = __shfl_sync(FULLMASK, o.x, t * 8 + cj);
                        oc[t].y = __shfl_sync(FULLMASK, o.y, t * 8 + cj);
                    }
                    float sr = 0.f, si = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float2 ork = make_float2(__shfl_sync(FULLMASK, o.x, ri * 8 + k), __shfl_sync(FULLMASK, o.y, ri * 8 + k));
                        float wr = 0.f, wi = 0.f;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const bool use = (usemask >> (k * 4 + t)) & 1u;
                            const float cx = c[k][t].x, cy = c[k][t].y;
                            const float m2 = fmaf(cx, cx, cy * cy);
                            const float rr = use ? fast_rsqrt(m2) : 0.f;
                            const float ex = (m2 > 0.f) ? cx * rr : (use ? 1.f : 0.f);
                            const float ey = (m2 > 0.f) ? cy * rr : 0.f;
                            wr = fmaf(ex, oc[t].x, wr); wr = fmaf(-ey, oc[t].y, wr);
                            wi = fmaf(ex, oc[t].y, wi); wi = fmaf(ey, oc[t].x, wi);
                        }
                        // conj(o_row) * w
                        sr += ork.x * wr + ork.y * wi;
                        si += ork.x * wi - ork.y * wr;
                    }
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
            if (myrow < N) a.out[(long)myrow * npix_block + pg] = o;
            if (lane == 0) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
            __syncwarp();
        }
    }
    if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_pix);
        atomicAdd(&a.stats[1], st_it);
        atomicAdd(&a.stats[3], st_cap);
    }
}

template <int B>
static cudaError_t launch_fast_t(const EvdArgs& a, cudaStream_t st) {
    typedef FastCfg<B> Cfg;
    const size_t lut = ((size_t)a.nulong * 32 * sizeof(short2) + 15) & ~(size_t)15;
    const size_t smem = lut + (size_t)Cfg::SMEM_PER_WARP * Cfg::WARPS;
    cudaError_t e = cudaFuncSetAttribute(k_evd_fast<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_evd_fast<B>, Cfg::WARPS * 32, smem);
    if (occ < 1) occ = 1;
    const long total_pairs = (long)a.n_lines * ((a.cols + 1) / 2);
    int mult = 8;
    if (const char* e = getenv("FRINGE_EVD_CHUNKS")) { const int v = atoi(e); if (v > 0) mult = v; }
    long grid = (long)nsm * occ * mult;
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

int evd_fast_block(int) { return 0; }      // interleaved complex pixel-major layout

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
