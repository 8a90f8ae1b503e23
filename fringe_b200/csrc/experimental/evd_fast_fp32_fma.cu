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
//               (full Hermitian, row stride = padded order) and re-read row-per-lane into
//               registers; the power iteration then needs only the broadcast vector from
//               shared memory (15 LDS.128 per 120 FMAs, 8 independent FMA chains).
//               Residual / normalisation reductions run every 4th iteration.
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

// profiling build only (-DFRINGE_PHASE_CLOCKS): per-phase warp cycles into stats[4..]
#ifdef FRINGE_PHASE_CLOCKS
#define PHASE_DECL long long ph_t = clock64(); unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
#define PHASE_MARK(k) { const long long ph_n = clock64(); ph[k] += (unsigned long long)(ph_n - ph_t); ph_t = ph_n; }
#define PHASE_FLUSH if (a.stats && lane == 0) { for (int k = 0; k < 6; ++k) atomicAdd(&a.stats[4 + k], ph[k]); }
#else
#define PHASE_DECL
#define PHASE_MARK(k)
#define PHASE_FLUSH
#endif

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
    // ... and the SHP index lists of the two pixels (2 x 64 ints)
    static constexpr int SMEM_PER_WARP =
        (((2 * MAT + 2 * 32) * (int)sizeof(float2) + 2 * NPAD * (int)sizeof(float) + 128 * (int)sizeof(int)) + 15) & ~15;
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
    int* s_list = reinterpret_cast<int*>(s_pw + 2 * NPAD);                 // [2][64]

    // block coordinates of this lane inside its pixel group
    int bi = 0, bj = 0;
    {
        int k = l, rowlen = 5;
        while (bi < 4 && k >= rowlen) { k -= rowlen; ++bi; --rowlen; }
        bj = bi + k;
    }
    const bool blk_active = (l < 15);
    if (!blk_active) { bi = 0; bj = 0; }           // lanes 15 / 31 shadow block (0,0); their results are never stored
    const int oa = B * bi, ob = B * bj;

    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2);
    const int BW = a.bandwidth;
    // bit j set: the pair (lane, j) enters the temporal-coherence sum (j > lane, inside the
    // matrix and, for STBAS, inside the band)
    uint32_t usemask = 0u;
    for (int j = 0; j < N; ++j)
        if (j > lane && (!isstbas || (j - lane) <= BW)) usemask |= (1u << j);
    // number of (i<j) pairs in the temporal-coherence sum (evd.cpp:773-784)
    float inv_pairs;
    {
        int cnt = 0;
        for (int i = 0; i < N; ++i) cnt += isstbas ? min(BW, N - 1 - i) : (N - 1 - i);
        inv_pairs = 1.0f / (float)cnt;
    }
    const long npix_block = (long)a.cols * a.lines;

    const int pairs_per_row = (a.cols + 1) >> 1;
    const long total_pairs = (long)a.n_lines * pairs_per_row;
    const long chunk = (total_pairs + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * chunk;
    const long end = min(total_pairs, beg + chunk);
    unsigned long long st_pix = 0, st_it = 0, st_cap = 0;
    PHASE_DECL

#pragma unroll 1
    for (long pr = beg + warp; pr < end; pr += Cfg::WARPS) {
        const int row = a.first_line + (int)(pr / pairs_per_row);
        const int col0 = 2 * (int)(pr % pairs_per_row);
        const int mycol = col0 + grp;
        const bool pix_exists = mycol < a.cols;

        // ------------------------- covariance (evd.cpp:537-564) -------------------------
        float2 acc[B][B];
#pragma unroll
        for (int i = 0; i < B; ++i)
#pragma unroll
            for (int j = 0; j < B; ++j) acc[i][j] = make_float2(0.f, 0.f);
        // SHP address lists.  All 32 lanes turn one 32-bit mask word of one pixel into sample-vector
        // indices at once (bit -> window offset -> bounds check, ballot + popcount for the
        // compacted slot) and park them in shared memory; the accumulation loop below then costs
        // one shared-memory read and one multiply per SHP instead of ~30 integer instructions
        // (ncu: the dispatch stalls of this phase sat on exactly those instructions).  Lists are
        // built for two mask words (64 window positions) at a time.
        int npix = 0;                                 // group-uniform count of usable SHPs
        const long p0 = (long)row * a.cols + col0;
        const bool ex1 = (col0 + 1) < a.cols;
        const bool con0 = (__ldg(&a.wts[p0 * a.nulong + (center >> 5)]) >> (center & 31)) & 1u;
        const bool con1 = ex1 && ((__ldg(&a.wts[(p0 + 1) * a.nulong + (center >> 5)]) >> (center & 31)) & 1u);
        const bool center_on = grp ? con1 : con0;
        const int zero_idx = (int)npix_block;         // index of the all-zero sample vector
#pragma unroll 1
        for (int w0 = 0; w0 < a.nulong; w0 += 2) {
            int n0 = 0, n1 = 0;                       // list lengths of the two pixels (warp-uniform)
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int w = w0 + c;
                    const bool live = (w < a.nulong) && (g ? con1 : con0);
                    const uint32_t word = live ? __ldg(&a.wts[(p0 + g) * a.nulong + w]) : 0u;
                    const short2 d = s_off[min(w, a.nulong - 1) * 32 + lane];
                    const int yy = row + d.x, xx = col0 + g + d.y;
                    const bool ok = ((word >> lane) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                    const uint32_t V = __ballot_sync(FULLMASK, ok);
                    int& n = g ? n1 : n0;
                    if (ok) s_list[g * 64 + n + __popc(V & ((1u << lane) - 1u))] = yy * a.cols + xx;
                    n += __popc(V);
                }
            }
            const int trips = (max(n0, n1) + 1) & ~1;   // SHPs are consumed two at a time
            // pad the shorter list(s) with the zero vector
            for (int k = n0 + lane; k < trips; k += 32) s_list[k] = zero_idx;
            for (int k = n1 + lane; k < trips; k += 32) s_list[64 + k] = zero_idx;
            __syncwarp();
            PHASE_MARK(0)
            npix += grp ? n1 : n0;
            const int* lst = s_list + grp * 64;
            // two-stage software pipeline: while one SHP is accumulated the next one's samples
            // are already in flight
            Operands<B> opA, opB;
            if (trips > 0) opA.load(a.zpix + (long)lst[0] * NPAD, oa, ob);
#pragma unroll 1
            for (int t = 0; t < trips; t += 2) {
                opB.load(a.zpix + (long)lst[t + 1] * NPAD, oa, ob);
                accumulate<B>(acc, opA);
                opA.load(a.zpix + (long)lst[min(t + 2, trips - 1)] * NPAD, oa, ob);
                accumulate<B>(acc, opB);
            }
            PHASE_MARK(1)
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
        PHASE_MARK(2)

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
            if (solve_g) {
                ++st_pix;
                const int r = (lane < NPAD) ? lane : (NPAD - 1);
                float2 c[NPAD];
                if (NPAD % 2 == 0) {
                    const float4* rowp = reinterpret_cast<const float4*>(s_mat + g * Cfg::MAT + r * NPAD);
#pragma unroll
                    for (int j = 0; j < NPAD; j += 2) {
                        const float4 v = rowp[j >> 1];
                        c[j] = make_float2(v.x, v.y); c[j + 1] = make_float2(v.z, v.w);
                    }
                } else {
                    const float2* rp2 = s_mat + g * Cfg::MAT + r * NPAD;
#pragma unroll
                    for (int j = 0; j < NPAD; ++j) c[j] = rp2[j];
                }
                // rows >= N of the padded matrix are exact zeros; only lanes beyond the padded
                // order (which re-read the last row) have to be silenced
                const float live = (lane < NPAD) ? 1.f : 0.f;
                if (isstbas) {                                   // evd.cpp:695-706 band limit
#pragma unroll
                    for (int j = 0; j < NPAD; ++j)
                        if (abs(j - lane) > BW) c[j] = make_float2(0.f, 0.f);
                }
                // start vector: column k0 of C (warm-starting from the previous pixel's eigenvector was
                // measured at -3 % but makes results depend on the block schedule at the 1e-7 level)
                float2 x;
                {
                    const float2 v = s_mat[g * Cfg::MAT + k0 * NPAD + r];
                    const float keep = (isstbas && abs(k0 - lane) > BW) ? 0.f : live;
                    x = make_float2(v.x * keep, -v.y * keep);
                    {   // unit-modulus start: the dominant eigenvector of a coherence matrix has nearly uniform magnitudes
                        const float m2 = x.x * x.x + x.y * x.y;
                        const float rs = (m2 > 0.f) ? fast_rsqrt(m2) : 0.f;
                        x.x *= rs; x.y *= rs;
                    }
                    float n2 = x.x * x.x + x.y * x.y;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(FULLMASK, n2, s);
                    const float sc = rsqrtf(n2);
                    x.x *= sc; x.y *= sc;
                }
                // Power iteration with heavy-ball momentum: x+ = C x / lambda - beta * x-.  beta = 0
                // (plain power iteration) until the decay rate r ~ lambda2/lambda1 of the residual
                // has been observed over one check interval; then beta = (0.95 r / 2)^2, which is
                // the optimal Chebyshev-type acceleration if r is exact and merely a weaker
                // acceleration if r is underestimated (never a divergence: beta < 1/4).
                PHASE_MARK(3)
                float lam = 1.f, inv_lam = 1.f, beta = 0.f, rho_prev = -1.f;
                float2 xp = make_float2(0.f, 0.f);
                int it = 0, buf = 0;
                bool conv = false;
                const int kMaxIter = (a.force_generic >> 1) ? (a.force_generic >> 1) : 1000;   // upper bits: timing experiment only
                const float tol2 = 4.0e-12f;
#pragma unroll 1
                for (; it < kMaxIter; ++it) {
                    float2* xv = s_vec + buf * 32;
                    buf ^= 1;
                    xv[lane] = x;
                    __syncwarp();
                    // y = C x with 8 independent FMA chains
                    float r0 = 0.f, r1 = 0.f, r2a = 0.f, r3 = 0.f, i0 = 0.f, i1 = 0.f, i2 = 0.f, i3 = 0.f;
                    const float4* xv4 = reinterpret_cast<const float4*>(xv);
#pragma unroll
                    for (int j = 0; j + 1 < NPAD; j += 2) {
                        const float4 q = xv4[j >> 1];
                        r0 = fmaf(c[j].x, q.x, r0); r1 = fmaf(-c[j].y, q.y, r1);
                        i0 = fmaf(c[j].x, q.y, i0); i1 = fmaf(c[j].y, q.x, i1);
                        r2a = fmaf(c[j + 1].x, q.z, r2a); r3 = fmaf(-c[j + 1].y, q.w, r3);
                        i2 = fmaf(c[j + 1].x, q.w, i2); i3 = fmaf(c[j + 1].y, q.z, i3);
                    }
                    if (NPAD % 2 == 1) {
                        const float2 q = xv[NPAD - 1];
                        r0 = fmaf(c[NPAD - 1].x, q.x, r0); r1 = fmaf(-c[NPAD - 1].y, q.y, r1);
                        i0 = fmaf(c[NPAD - 1].x, q.y, i0); i1 = fmaf(c[NPAD - 1].y, q.x, i1);
                    }
                    const float yr = ((r0 + r1) + (r2a + r3)) * live, yi = ((i0 + i1) + (i2 + i3)) * live;
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
                PHASE_MARK(4)
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
                        const float2 z = __ldg(&a.zpix[pg * NPAD + lane]);
                        cr = z.x * o.x + z.y * o.y;
                        ci = z.y * o.x - z.x * o.y;
                    }
                    // ---------------- temporal coherence (evd.cpp:770-786) --------------
                    float2* ov = s_vec + buf * 32;
                    buf ^= 1;
                    ov[lane] = o;
                    __syncwarp();
                    float wr = 0.f, wi = 0.f;
                    const float4* ov4 = reinterpret_cast<const float4*>(ov);
#pragma unroll
                    for (int j = 0; j < NPAD; ++j) {
                        // pairs (lane, j) with j > lane, inside the matrix and (STBAS) the band;
                        // e = C_ij / |C_ij| (arg(0) = 0 as in the reference)
                        const bool use = (usemask >> j) & 1u;
                        const float m2 = fmaf(c[j].x, c[j].x, c[j].y * c[j].y);
                        const float rr = use ? fast_rsqrt(m2) : 0.f;
                        const float ex = (m2 > 0.f) ? c[j].x * rr : (use ? 1.f : 0.f);
                        const float ey = (m2 > 0.f) ? c[j].y * rr : 0.f;
                        float2 oj;
                        if (NPAD % 2 == 0) {
                            const float4 q = ov4[j >> 1];
                            oj = (j & 1) ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
                        } else oj = ov[j];
                        wr = fmaf(ex, oj.x, wr); wr = fmaf(-ey, oj.y, wr);
                        wi = fmaf(ex, oj.y, wi); wi = fmaf(ey, oj.x, wi);
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
    }
    PHASE_FLUSH
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

int evd_fast_block(int) { return 0; }      // this kernel reads the interleaved complex pixel-major layout

bool evd_fast_supported(const EvdArgs& a) {
    if (a.zblock < 0)                           // tensor-pipe layout (evd_mma.cu)
        return a.variant == 0 && (a.method == 0 || a.method == 2) && a.NP == 64 && evd_mma_order(a.bands) > 0;
    return a.variant == 0 && (a.method == 0 || a.method == 2) && evd_fast_padded_bands(a.bands) > 0 &&
           a.NP == evd_fast_padded_bands(a.bands);
}

cudaError_t launch_evd_fast(const EvdArgs& a, cudaStream_t st) {
    if (a.zblock < 0) return launch_evd_mma(a, st);
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
