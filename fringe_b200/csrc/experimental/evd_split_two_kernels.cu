// Covariance and eigen solve as two kernels (sm_100a), bands <= 30, EVD / STBAS.
//
// The fused kernel (evd_fast.cu) is pinned at 168 registers by its covariance accumulators and runs
// 12 warps per SM; ncu shows it latency-bound in every phase.  Here the two phases get their own
// launch configuration:
//   k_cov<B>  same mapping as the fused kernel's first half (two pixels per warp, 15 B x B register
//             blocks each, SHP prefetch pipeline) but no shared-memory matrices, so L1 is not carved
//             up; the normalised coherence matrix goes to a global scratch buffer as a full
//             Hermitian [NPAD][NPAD] matrix per pixel (16-byte stores, row segments of one block).
//   k_eig<B>  one warp per pixel, lane = row: 15 coalesced 16-byte loads fetch the row, then the same
//             momentum power iteration and epilogue as the fused kernel -- at ~100 registers, i.e.
//             16 warps per SM.
// The scratch traffic (2 x 8*NPAD^2 bytes per pixel) is ~1.3 TB/s at the rates reached here, about a
// fifth of HBM bandwidth, and overlaps with the arithmetic of the other warps.  The image is walked
// in strips of lines so that the scratch buffer stays bounded.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace fringe {

#define FULLMASK 0xffffffffu

__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

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

struct SplitArgs {
    EvdArgs a;
    float2* cmat;          // [strip pixels][NPAD][NPAD]
    uint8_t* flag;         // [strip pixels] 1 = solve
    int strip_line0;       // first line of the strip, relative to a.first_line
    int strip_lines;
};

// ======================================================================================
template <int B>
__global__ void __launch_bounds__(128, 3) k_cov(const SplitArgs sa) {
    const EvdArgs& a = sa.a;
    constexpr int NPAD = 5 * B;
    constexpr int WARPS = 4;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 4, l = lane & 15;
    const int N = a.bands;

    short2* s_off = reinterpret_cast<short2*>(s_raw);
    const int WX = 2 * a.Nx + 1, W = WX * (2 * a.Ny + 1), center = a.Ny * WX + a.Nx;
    for (int f = threadIdx.x; f < a.nulong * 32; f += blockDim.x) {
        const int fy = f / WX;
        s_off[f] = (f < W) ? make_short2((short)(fy - a.Ny), (short)(f - fy * WX - a.Nx))
                           : make_short2((short)-30000, (short)-30000);
    }
    __syncthreads();
    const int lut_bytes = ((a.nulong * 32 * (int)sizeof(short2)) + 15) & ~15;
    float* s_pw = reinterpret_cast<float*>(s_raw + lut_bytes) + warp * 2 * NPAD;     // [2][NPAD]

    int bi = 0, bj = 0;
    {
        int k = l, rowlen = 5;
        while (bi < 4 && k >= rowlen) { k -= rowlen; ++bi; --rowlen; }
        bj = bi + k;
    }
    const bool blk_active = (l < 15);
    const int oa = B * bi, ob = B * bj;
    const long npix_block = (long)a.cols * a.lines;
    const float2* zero_row = a.zpix + npix_block * NPAD;

    const int pairs_per_row = (a.cols + 1) >> 1;
    const long total_pairs = (long)sa.strip_lines * pairs_per_row;
    const long chunk = (total_pairs + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * chunk;
    const long end = min(total_pairs, beg + chunk);

#pragma unroll 1
    for (long pr = beg + warp; pr < end; pr += WARPS) {
        const int srow = (int)(pr / pairs_per_row);                       // line inside the strip
        const int row = a.first_line + sa.strip_line0 + srow;
        const int col0 = 2 * (int)(pr % pairs_per_row);
        const int mycol = col0 + grp;
        const bool pix_exists = mycol < a.cols;
        const long p = (long)row * a.cols + mycol;
        const long sp = (long)srow * a.cols + mycol;                      // pixel index inside the strip

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

        if (pix_exists && l == 0) sa.flag[sp] = solve_me ? 1 : 0;

        // ------------------------- coherence (evd.cpp:569-582) --------------------------
        __syncwarp();
        if (blk_active && bi == bj) {
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int t = oa + i;
                s_pw[grp * NPAD + t] = (t < N) ? sqrtf(acc[i][i].x) : CUDART_INF_F;
            }
        }
        __syncwarp();
        if (blk_active && solve_me) {
            float ia[B], ib[B];
#pragma unroll
            for (int i = 0; i < B; ++i) { ia[i] = fast_rcp(s_pw[grp * NPAD + oa + i]); ib[i] = fast_rcp(s_pw[grp * NPAD + ob + i]); }
            float2* mat = sa.cmat + sp * (NPAD * NPAD);
            const bool diag = (bi == bj);
            // row segments of the block (and of its conjugate mirror), B complex = 8B bytes each
#pragma unroll
            for (int i = 0; i < B; ++i) {
                float2 seg[B], mir[B];
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const float s1 = ia[i] * ib[j];
                    seg[j] = make_float2(acc[i][j].x * s1, acc[i][j].y * s1);          // C(oa+i, ob+j)
                    const float s2 = ia[j] * ib[i];
                    mir[j] = make_float2(acc[j][i].x * s2, -acc[j][i].y * s2);         // C(ob+i, oa+j) = conj C(oa+j, ob+i)
                }
                if (diag) {
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        if (j < i) seg[j] = mir[j];
                        else if (j == i) seg[j] = make_float2((oa + i < N) ? 1.f : 0.f, 0.f);
                    }
                }
                float2* up = mat + (oa + i) * NPAD + ob;
                float2* lo = mat + (ob + i) * NPAD + oa;
                if (B % 2 == 0) {
#pragma unroll
                    for (int j = 0; j < B; j += 2) {
                        *reinterpret_cast<float4*>(up + j) = make_float4(seg[j].x, seg[j].y, seg[j + 1].x, seg[j + 1].y);
                        if (!diag) *reinterpret_cast<float4*>(lo + j) = make_float4(mir[j].x, mir[j].y, mir[j + 1].x, mir[j + 1].y);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < B; ++j) { up[j] = seg[j]; if (!diag) lo[j] = mir[j]; }
                }
            }
        }
        __syncwarp();
    }
}

// ======================================================================================
// Eigen solve + epilogue.  A warp owns TWO pixels: half-warp h = lane>>4 works on pixel 2w+h and
// each of its lanes holds two matrix rows (l and l+HALF).  Per power iteration a lane then reads
// its own pixel's vector from shared memory (15 LDS.128, two distinct addresses per instruction)
// for 2 x 30 complex FMAs -- half the shared-memory wavefronts per pixel of the lane-per-row
// layout, which ncu showed to be the limit of this phase (short-scoreboard stalls, 60 wavefronts
// per product).
template <int B>
__global__ void __launch_bounds__(128, 3) k_eig(const SplitArgs sa) {
    const EvdArgs& a = sa.a;
    constexpr int NPAD = 5 * B;
    constexpr int HALF = (NPAD + 1) / 2;            // rows l and l+HALF per lane, l < HALF <= 15
    constexpr int WARPS = 4;
    __shared__ __align__(16) float2 s_vec_all[WARPS][2][2][32];     // [warp][buffer][pixel][entry]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l = lane & 15;
    const int N = a.bands;
    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2);
    const int BW = a.bandwidth;
    const int r0 = min(l, HALF - 1), r1 = min(l + HALF, NPAD - 1);
    const bool lane_rows = (l < HALF);
    const bool row1_ok = lane_rows && (l + HALF < NPAD);
    // bit j: pair (row, j) enters the temporal-coherence sum (j > row, inside the matrix / band)
    uint32_t use0 = 0u, use1 = 0u;
    for (int j = 0; j < N; ++j) {
        if (lane_rows && j > r0 && (!isstbas || (j - r0) <= BW)) use0 |= (1u << j);
        if (row1_ok && j > r1 && (!isstbas || (j - r1) <= BW)) use1 |= (1u << j);
    }
    float inv_pairs;
    {
        int cnt = 0;
        for (int i = 0; i < N; ++i) cnt += isstbas ? min(BW, N - 1 - i) : (N - 1 - i);
        inv_pairs = 1.0f / (float)cnt;
    }
    const long npix_block = (long)a.cols * a.lines;
    const long total = (long)sa.strip_lines * a.cols;
    const long npairs = (total + 1) >> 1;
    const long chunk = (npairs + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * chunk;
    const long end = min(npairs, beg + chunk);
    unsigned long long st_pix = 0, st_it = 0, st_cap = 0;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;      // shuffles stay inside the half-warp

#pragma unroll 1
    for (long pw = beg + warp; pw < end; pw += WARPS) {
        const long sp = 2 * pw + half;
        const bool exists = sp < total;
        const long pg = (long)(a.first_line + sa.strip_line0) * a.cols + sp;
        const bool solve = exists && (sa.flag[sp] != 0);
        float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
        float tc = 0.f;
        float2 cmp = make_float2(0.f, 0.f);
        if (__any_sync(FULLMASK, solve)) {
            const float2* mat = sa.cmat + (solve ? sp : 0) * (NPAD * NPAD);
            float2 c0[NPAD], c1[NPAD];
            const float keep0 = (solve && lane_rows) ? 1.f : 0.f, keep1 = (solve && row1_ok) ? 1.f : 0.f;
            if (NPAD % 2 == 0) {
                const float4* p0 = reinterpret_cast<const float4*>(mat + r0 * NPAD);
                const float4* p1 = reinterpret_cast<const float4*>(mat + r1 * NPAD);
#pragma unroll
                for (int j = 0; j < NPAD; j += 2) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
                    if (solve) { v = __ldcs(p0 + (j >> 1)); u = __ldcs(p1 + (j >> 1)); }
                    c0[j] = make_float2(v.x * keep0, v.y * keep0); c0[j + 1] = make_float2(v.z * keep0, v.w * keep0);
                    c1[j] = make_float2(u.x * keep1, u.y * keep1); c1[j + 1] = make_float2(u.z * keep1, u.w * keep1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NPAD; ++j) {
                    float2 v = make_float2(0.f, 0.f), u = v;
                    if (solve) { v = __ldcs(mat + r0 * NPAD + j); u = __ldcs(mat + r1 * NPAD + j); }
                    c0[j] = make_float2(v.x * keep0, v.y * keep0);
                    c1[j] = make_float2(u.x * keep1, u.y * keep1);
                }
            }
            if (isstbas) {                                   // evd.cpp:695-706 band limit
#pragma unroll
                for (int j = 0; j < NPAD; ++j) {
                    if (abs(j - r0) > BW) c0[j] = make_float2(0.f, 0.f);
                    if (abs(j - r1) > BW) c1[j] = make_float2(0.f, 0.f);
                }
            }
            // start vector: column k0 of C = conj of the entries C(k0, r)
            float2 x0 = make_float2(0.f, 0.f), x1 = x0;
            if (solve) {
                const float2 v0 = __ldg(mat + k0 * NPAD + r0), v1 = __ldg(mat + k0 * NPAD + r1);
                const float s0 = (isstbas && abs(k0 - r0) > BW) ? 0.f : keep0;
                const float s1 = (isstbas && abs(k0 - r1) > BW) ? 0.f : keep1;
                x0 = make_float2(v0.x * s0, -v0.y * s0);
                x1 = make_float2(v1.x * s1, -v1.y * s1);
            }
            {
                float n2 = x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y;
#pragma unroll
                for (int s = 8; s > 0; s >>= 1) n2 += __shfl_xor_sync(FULLMASK, n2, s);
                const float sc = solve ? rsqrtf(n2) : 0.f;
                x0.x *= sc; x0.y *= sc; x1.x *= sc; x1.y *= sc;
            }
            // power iteration with heavy-ball momentum (see evd_fast.cu); both pixels in lock-step,
            // a converged (or absent) pixel is frozen
            float lam = 1.f, inv_lam = 1.f, beta = 0.f, rho_prev = -1.f;
            float2 xp0 = make_float2(0.f, 0.f), xp1 = xp0;
            int it = 0, my_it = 0, buf = 0;
            bool conv = !solve;
            const int kMaxIter = 1000;
            const float tol2 = 4.0e-12f;
#pragma unroll 1
            for (; it < kMaxIter; ++it) {
                float2* xv = &s_vec_all[warp][buf][half][0];
                buf ^= 1;
                if (lane_rows) { xv[r0] = x0; if (row1_ok) xv[r1] = x1; }
                __syncwarp();
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;   // row 0: re/im chains
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;   // row 1
                const float4* xv4 = reinterpret_cast<const float4*>(xv);
#pragma unroll
                for (int j = 0; j + 1 < NPAD; j += 2) {
                    const float4 q = xv4[j >> 1];
                    a0 = fmaf(c0[j].x, q.x, a0); a1 = fmaf(-c0[j].y, q.y, a1);
                    b0 = fmaf(c0[j].x, q.y, b0); b1 = fmaf(c0[j].y, q.x, b1);
                    a2 = fmaf(c0[j + 1].x, q.z, a2); a3 = fmaf(-c0[j + 1].y, q.w, a3);
                    b2 = fmaf(c0[j + 1].x, q.w, b2); b3 = fmaf(c0[j + 1].y, q.z, b3);
                    d0 = fmaf(c1[j].x, q.x, d0); d1 = fmaf(-c1[j].y, q.y, d1);
                    e0 = fmaf(c1[j].x, q.y, e0); e1 = fmaf(c1[j].y, q.x, e1);
                    d2 = fmaf(c1[j + 1].x, q.z, d2); d3 = fmaf(-c1[j + 1].y, q.w, d3);
                    e2 = fmaf(c1[j + 1].x, q.w, e2); e3 = fmaf(c1[j + 1].y, q.z, e3);
                }
                if (NPAD % 2 == 1) {
                    const float2 q = xv[NPAD - 1];
                    a0 = fmaf(c0[NPAD - 1].x, q.x, a0); a1 = fmaf(-c0[NPAD - 1].y, q.y, a1);
                    b0 = fmaf(c0[NPAD - 1].x, q.y, b0); b1 = fmaf(c0[NPAD - 1].y, q.x, b1);
                    d0 = fmaf(c1[NPAD - 1].x, q.x, d0); d1 = fmaf(-c1[NPAD - 1].y, q.y, d1);
                    e0 = fmaf(c1[NPAD - 1].x, q.y, e0); e1 = fmaf(c1[NPAD - 1].y, q.x, e1);
                }
                const float2 y0 = make_float2((a0 + a1) + (a2 + a3), (b0 + b1) + (b2 + b3));
                const float2 y1 = make_float2((d0 + d1) + (d2 + d3), (e0 + e1) + (e2 + e3));
                if ((it & 3) != 3) {
                    if (!conv) {
                        const float2 n0 = make_float2(fmaf(-beta, xp0.x, y0.x * inv_lam), fmaf(-beta, xp0.y, y0.y * inv_lam));
                        const float2 n1 = make_float2(fmaf(-beta, xp1.x, y1.x * inv_lam), fmaf(-beta, xp1.y, y1.y * inv_lam));
                        xp0 = x0; xp1 = x1; x0 = n0; x1 = n1;
                    }
                } else {
                    // Rayleigh quotient, residual, renormalisation -- per half-warp
                    float xy = x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y;
                    float xx = x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y;
#pragma unroll
                    for (int s = 8; s > 0; s >>= 1) {
                        xy += __shfl_xor_sync(FULLMASK, xy, s);
                        xx += __shfl_xor_sync(FULLMASK, xx, s);
                    }
                    const float lam_new = xy / xx;
                    const float rx0 = y0.x - lam_new * x0.x, ry0 = y0.y - lam_new * x0.y;
                    const float rx1 = y1.x - lam_new * x1.x, ry1 = y1.y - lam_new * x1.y;
                    float rr2 = rx0 * rx0 + ry0 * ry0 + rx1 * rx1 + ry1 * ry1;
                    float y2 = y0.x * y0.x + y0.y * y0.y + y1.x * y1.x + y1.y * y1.y;
#pragma unroll
                    for (int s = 8; s > 0; s >>= 1) {
                        rr2 += __shfl_xor_sync(FULLMASK, rr2, s);
                        y2 += __shfl_xor_sync(FULLMASK, y2, s);
                    }
                    if (!conv) {
                        lam = lam_new;
                        inv_lam = 1.0f / lam;
                        const float rho2 = rr2 / (lam * lam * xx);
                        if (rho2 <= tol2) {
                            const float sc = rsqrtf(y2);
                            x0 = make_float2(y0.x * sc, y0.y * sc); x1 = make_float2(y1.x * sc, y1.y * sc);
                            conv = true; my_it = it + 1;
                        } else {
                            if (beta == 0.f && rho_prev > 0.f && rho2 < rho_prev) {
                                const float r = sqrtf(sqrtf(sqrtf(rho2 / rho_prev)));
                                const float hb = 0.475f * r;
                                beta = hb * hb;
                            }
                            rho_prev = rho2;
                            const float sc = rsqrtf(xx);
                            const float2 n0 = make_float2(fmaf(-beta, xp0.x, y0.x * inv_lam) * sc, fmaf(-beta, xp0.y, y0.y * inv_lam) * sc);
                            const float2 n1 = make_float2(fmaf(-beta, xp1.x, y1.x * inv_lam) * sc, fmaf(-beta, xp1.y, y1.y * inv_lam) * sc);
                            xp0 = make_float2(x0.x * sc, x0.y * sc); xp1 = make_float2(x1.x * sc, x1.y * sc);
                            x0 = n0; x1 = n1;
                        }
                    }
                    if (__all_sync(FULLMASK, conv)) { ++it; break; }
                }
            }
            if (solve && l == 0) { ++st_pix; st_it += conv ? my_it : it; st_cap += conv ? 0 : 1; }
            // ---------------- epilogue (evd.cpp:738-786), per half-warp ----------------
            float2* xv = &s_vec_all[warp][buf][half][0];
            buf ^= 1;
            if (lane_rows) { xv[r0] = x0; if (row1_ok) xv[r1] = x1; }
            __syncwarp();
            const float2 ref = xv[k0];
            auto phasor = [&](float2 x, int r) -> float2 {
                float ux = x.x * ref.x + x.y * ref.y, uy = x.y * ref.x - x.x * ref.y;
                const float mm = ux * ux + uy * uy;
                if (mm == 0.f) {                  // arg(0) = 0 in the reference
                    const float rr = rsqrtf(ref.x * ref.x + ref.y * ref.y);
                    ux = ref.x * rr; uy = -ref.y * rr;
                } else { const float rr = fast_rsqrt(mm); ux *= rr; uy *= rr; }
                if (r == k0) { ux = 1.f; uy = 0.f; }
                return make_float2(ux, uy);
            };
            const bool good = solve && !(lam < 1.0e-6f);                 // evd.cpp:723-727
            if (good) {
                if (lane_rows && r0 < N) o0 = phasor(x0, r0);
                if (row1_ok && r1 < N) o1 = phasor(x1, r1);
            }
            float cr = 0.f, ci = 0.f;
            if (good) {
                if (lane_rows && r0 < N && r0 >= k0) {
                    const float2 z = __ldg(&a.zpix[pg * NPAD + r0]);
                    cr += z.x * o0.x + z.y * o0.y; ci += z.y * o0.x - z.x * o0.y;
                }
                if (row1_ok && r1 < N && r1 >= k0) {
                    const float2 z = __ldg(&a.zpix[pg * NPAD + r1]);
                    cr += z.x * o1.x + z.y * o1.y; ci += z.y * o1.x - z.x * o1.y;
                }
            }
            float2* ov = &s_vec_all[warp][buf][half][0];
            buf ^= 1;
            if (lane_rows) { ov[r0] = o0; if (row1_ok) ov[r1] = o1; }
            __syncwarp();
            float w0r = 0.f, w0i = 0.f, w1r = 0.f, w1i = 0.f;
            const float4* ov4 = reinterpret_cast<const float4*>(ov);
#pragma unroll
            for (int j = 0; j < NPAD; ++j) {
                float2 oj;
                if (NPAD % 2 == 0) {
                    const float4 q = ov4[j >> 1];
                    oj = (j & 1) ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
                } else oj = ov[j];
                {
                    const bool use = (use0 >> j) & 1u;
                    const float m2 = fmaf(c0[j].x, c0[j].x, c0[j].y * c0[j].y);
                    const float rr = use ? fast_rsqrt(m2) : 0.f;
                    const float ex = (m2 > 0.f) ? c0[j].x * rr : (use ? 1.f : 0.f);
                    const float ey = (m2 > 0.f) ? c0[j].y * rr : 0.f;
                    w0r = fmaf(ex, oj.x, w0r); w0r = fmaf(-ey, oj.y, w0r);
                    w0i = fmaf(ex, oj.y, w0i); w0i = fmaf(ey, oj.x, w0i);
                }
                {
                    const bool use = (use1 >> j) & 1u;
                    const float m2 = fmaf(c1[j].x, c1[j].x, c1[j].y * c1[j].y);
                    const float rr = use ? fast_rsqrt(m2) : 0.f;
                    const float ex = (m2 > 0.f) ? c1[j].x * rr : (use ? 1.f : 0.f);
                    const float ey = (m2 > 0.f) ? c1[j].y * rr : 0.f;
                    w1r = fmaf(ex, oj.x, w1r); w1r = fmaf(-ey, oj.y, w1r);
                    w1i = fmaf(ex, oj.y, w1i); w1i = fmaf(ey, oj.x, w1i);
                }
            }
            float sr = (o0.x * w0r + o0.y * w0i) + (o1.x * w1r + o1.y * w1i);
            float si = (o0.x * w0i - o0.y * w0r) + (o1.x * w1i - o1.y * w1r);
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) {
                sr += __shfl_xor_sync(FULLMASK, sr, s);
                si += __shfl_xor_sync(FULLMASK, si, s);
                cr += __shfl_xor_sync(FULLMASK, cr, s);
                ci += __shfl_xor_sync(FULLMASK, ci, s);
            }
            if (good) {
                tc = sqrtf(sr * sr + si * si) * inv_pairs;
                const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
                cmp = make_float2(cr * invn, ci * invn);
            } else if (solve) tc = -7.f;
            (void)hmask;
        }
        if (exists) {
            if (lane_rows && r0 < N) a.out[(long)r0 * npix_block + pg] = o0;
            if (row1_ok && r1 < N) a.out[(long)r1 * npix_block + pg] = o1;
            if (l == 0) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
        }
        __syncwarp();
    }
    if (a.stats && l == 0) {
        atomicAdd(&a.stats[0], st_pix);
        atomicAdd(&a.stats[1], st_it);
        atomicAdd(&a.stats[3], st_cap);
    }
}

template <int B>
static cudaError_t launch_split_t(const EvdArgs& a, float2* cmat, uint8_t* flag, long strip_pixels_cap, cudaStream_t st, int* launches) {
    constexpr int NPAD = 5 * B;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const size_t lut = ((size_t)a.nulong * 32 * sizeof(short2) + 15) & ~(size_t)15;
    const size_t smem_cov = lut + (size_t)4 * 2 * NPAD * sizeof(float);
    int strip = (int)(strip_pixels_cap / a.cols);
    if (strip < 1) return cudaErrorInvalidValue;
    if (strip > a.n_lines) strip = a.n_lines;
    for (int l0 = 0; l0 < a.n_lines; l0 += strip) {
        SplitArgs sa;
        sa.a = a; sa.cmat = cmat; sa.flag = flag; sa.strip_line0 = l0; sa.strip_lines = min(strip, a.n_lines - l0);
        const long pairs = (long)sa.strip_lines * ((a.cols + 1) / 2);
        long g1 = (long)nsm * 3 * 8;
        const long m1 = (pairs + 15) / 16;
        if (g1 > m1) g1 = m1;
        if (g1 < 1) g1 = 1;
        k_cov<B><<<(unsigned)g1, 128, smem_cov, st>>>(sa);
        const long pix = ((long)sa.strip_lines * a.cols + 1) / 2;      // pixel pairs
        long g2 = (long)nsm * 3 * 8;
        const long m2 = (pix + 15) / 16;
        if (g2 > m2) g2 = m2;
        if (g2 < 1) g2 = 1;
        k_eig<B><<<(unsigned)g2, 128, 0, st>>>(sa);
        if (launches) *launches += 2;
    }
    return cudaGetLastError();
}

size_t evd_split_bytes_per_pixel(int bands) {
    const int np = evd_fast_padded_bands(bands);
    return (size_t)np * np * sizeof(float2) + 1;
}

cudaError_t launch_evd_split(const EvdArgs& a, void* scratch, size_t scratch_bytes, cudaStream_t st, int* launches) {
    const size_t per_pixel = evd_split_bytes_per_pixel(a.bands);
    const long cap = (long)(scratch_bytes / per_pixel);
    const int NP = a.NP;
    float2* cmat = reinterpret_cast<float2*>(scratch);
    uint8_t* flag = reinterpret_cast<uint8_t*>(scratch) + (size_t)cap * NP * NP * sizeof(float2);
    if (launches) *launches = 0;
    switch (NP / 5) {
        case 2: return launch_split_t<2>(a, cmat, flag, cap, st, launches);
        case 3: return launch_split_t<3>(a, cmat, flag, cap, st, launches);
        case 4: return launch_split_t<4>(a, cmat, flag, cap, st, launches);
        case 5: return launch_split_t<5>(a, cmat, flag, cap, st, launches);
        case 6: return launch_split_t<6>(a, cmat, flag, cap, st, launches);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fringe
