// FP64 maximum-likelihood (MLE / EMI) and phase_link kernel for bands <= 32 (sm_100a).
//
// What the reference does per pixel (src/evd/evd.cpp:537-688 with method MLE -- the default,
// evd.hpp:66 -- and src/phase_link/phase_link.cpp:503-618): masked sample covariance (float
// products, double sums), coherence C, two positive-definiteness gates through LAPACK zheevr,
// inv(|C|) through zpotrf + zpotri, Hadamard product M = inv(|C|) o C, smallest eigenpair of M
// (zheevr), sentinels -2 / -4 / -5 / -6 / -7 on the way (phase_link: EVD fallback instead).
//
// How it is done here: one warp per pixel, and every N x N matrix of the chain lives in
// registers with one Hermitian row per lane -- lane i holds row i, statically indexed, so
// nothing spills and there is no shared-memory matrix traffic in the factorisations.
//   covariance    pairs (i < j) dealt round-robin to the 32 lanes; per SHP the N samples are
//                 staged once in shared memory, every lane forms its pairs' products exactly as
//                 libgcc's complex multiply rounds them and adds them in double (the double sums
//                 of float products are exact, so the order of the SHPs does not matter)
//   Cholesky      right-looking, column k broadcast through a double-buffered shared-memory
//                 vector: 1 LDS.128 + 4 DFMA per (k, j).  The rank-1 update runs over the *full*
//                 Hermitian row, so that after step i lane i also holds the (unscaled) conjugate
//                 of column i: the backward substitution needs no transposition
//   solves        forward / backward substitution with the pivot value broadcast by shuffles,
//                 10-14 instructions per step, no shared memory
//   gates         lambda_min(C) >= 1e-6 as Cholesky of C - 1e-6 I; pixels with fewer SHPs than
//                 bands are rank deficient and take the -2 sentinel without any arithmetic;
//                 lambda_min(|C|) >= 1e-6 is read off inv(|C|) (max row sum / max diagonal
//                 bracket 1 / lambda_min) and only decided by a second factorisation in between
//   inverse       real Cholesky, then lane c solves L L^T x = e_c privately from the packed
//                 factor in shared memory (uniform addresses: broadcasts)
//   eigen         inverse iteration with Cholesky-certified shifts (a successful factorisation
//                 of M - sigma I proves sigma < lambda_min, so the iteration cannot lock onto a
//                 wrong eigenvalue).  Rayleigh quotient and residual come for free from the
//                 solve (M y = sigma y + x), shifts follow the Kato-Temple bound with the gap
//                 estimated from the observed decay; start vector = phases of the middle column
//                 of C (scripts/sim_mle_solver.py: 3.0 factorisations + 5.2 solves per pixel
//                 against 4.4 + 8.9 for the round-1 policy)
//   epilogue      phase reference, compressed SLC, temporal coherence as evd.cpp:738-786
#include <math_constants.h>

#include "common.cuh"

namespace fringe {

namespace {

#define FULLM 0xffffffffu

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLM, v, o));
    return v;
}
__device__ __forceinline__ float wsumf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
    return v;
}

template <int NT>
struct MleCfg {
    static constexpr int WARPS = 4;
    static constexpr int MIN_CTAS = NT <= 20 ? 4 : 3;          // register budget 128 / 168 per thread
    static constexpr int BAND = 4;                            // rows of a CTA's pixel band
    static constexpr int NPAIR = NT * (NT - 1) / 2;
    static constexpr int NSLOT = (NPAIR + 31) / 32;           // covariance pairs per lane
    static constexpr int NPACK = NT * (NT + 1) / 2;           // packed lower triangle
    // per warp: Mp complex double packed | col 2 x NT complex double | Lr real packed | dv NT |
    //           Cf float2 packed | zs 2 x NT float2 | list 64 ints
    static constexpr int SMEM_PER_WARP = 16 * NPACK + 32 * NT + 8 * NPACK + 8 * NT + 8 * NPACK + 16 * NT + 256;
};

// ---------------------------------------------------------------------------------------
// Complex Cholesky of the Hermitian matrix held one row per lane (fr + i fi), in place.
// On success lane i holds L(i, k) for k < i (scaled), the unscaled trailing values a(i, j), j > i,
// as of step i (= conj(L(j, i)) / rs_i), and rs_own = 1 / L(i, i).  Returns false (warp-uniform)
// on a non-positive pivot, the failure LAPACK zpotrf reports.
// ---------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ bool chol_c_reg(double (&fr)[NT], double (&fi)[NT], double& rs_own, double2* colb,
                                           int N, int lane) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        if (k < N && ok) {
            const double d = __shfl_sync(FULLM, fr[k], k);
            ok = __all_sync(FULLM, d > 0.0);             // d is the same in every lane; the vote tells the compiler so
            if (ok) {
                const double rs = rsqrt(d);
                rs_own = (lane == k) ? rs : rs_own;
                // column k is scaled in every lane: rows > k get L(i,k), rows < k keep their final a(i,k) times
                // rs_k (the backward substitution accounts for the factor), row k gets sqrt(d)
                const double lr = fr[k] * rs, li = fi[k] * rs;
                fr[k] = lr; fi[k] = li;
                double2* cb = colb + (k & 1) * NT;
                if (lane < NT) cb[lane] = make_double2(lr, li);
                __syncwarp();
                // rows <= k are final: they run the update with a zero multiplier instead of a branch.
                // Columns >= N (padding up to NT) are updated too; nothing ever reads them.
                const bool act = lane > k;
                const double mr = act ? lr : 0.0, mi = act ? li : 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j) {           // constant trip count: j > k folds once k is unrolled
                    if (j > k) {
                        const double2 v = cb[j];         // a(i,j) -= l(i,k) conj(l(j,k))
                        fr[j] = fma(-mr, v.x, fr[j]); fr[j] = fma(-mi, v.y, fr[j]);
                        fi[j] = fma(-mi, v.x, fi[j]); fi[j] = fma(mr, v.y, fi[j]);
                    }
                }
            }
        }
    }
    return ok;
}

// x <- (L L^H)^-1 x with the factor left by chol_c_reg; lane i holds x(i).  Branch-free: a lane's
// coefficient is zero on the steps that do not concern it.
//   forward   L y = x, column oriented: step k broadcasts y(k) = x(k) rs_k, rows > k subtract L(i,k) y(k)
//   backward  L^H z = y: lane i holds u(i,k) = conj(L(k,i)) rs_k / rs_i for k > i, so with t(k) = z(k) / rs_k
//             z(i) = (y(i) - rs_i sum_k u(i,k) t(k)) rs_i  and  t(i) = y(i) - rs_i acc(i)
template <int NT>
__device__ __forceinline__ void solve_c_reg(const double (&fr)[NT], const double (&fi)[NT], double rs_own, int N,
                                            int lane, double& xr, double& xi) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        if (k < N) {
            const double yr = __shfl_sync(FULLM, xr * rs_own, k), yi = __shfl_sync(FULLM, xi * rs_own, k);
            const double cr = (lane > k) ? fr[k] : 0.0, ci = (lane > k) ? fi[k] : 0.0;
            xr = fma(-cr, yr, xr); xr = fma(ci, yi, xr);
            xi = fma(-cr, yi, xi); xi = fma(-ci, yr, xi);
        }
    }
    xr *= rs_own; xi *= rs_own;                        // y(i)
    double ar = 0.0, ai = 0.0;
#pragma unroll
    for (int k = NT - 1; k >= 0; --k) {
        if (k < N) {
            const double tr = __shfl_sync(FULLM, fma(-ar, rs_own, xr), k), ti = __shfl_sync(FULLM, fma(-ai, rs_own, xi), k);
            const double cr = (lane < k) ? fr[k] : 0.0, ci = (lane < k) ? fi[k] : 0.0;
            ar = fma(cr, tr, ar); ar = fma(-ci, ti, ar);
            ai = fma(cr, ti, ai); ai = fma(ci, tr, ai);
        }
    }
    xr = fma(-ar, rs_own, xr) * rs_own; xi = fma(-ai, rs_own, xi) * rs_own;     // z(i)
}

// Real Cholesky of the symmetric matrix held one row per lane; the factor goes to shared memory
// (packed lower, Lr[i(i+1)/2 + k] = L(i,k) for k < i) and dv[k] = 1 / L(k,k).
template <int NT>
__device__ __forceinline__ bool chol_r_reg(double (&g)[NT], double* Lr, double* dv, int N, int lane) {
    bool ok = true;
    const int tri = lane * (lane + 1) / 2;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        if (k < N && ok) {
            const double d = __shfl_sync(FULLM, g[k], k);
            ok = __all_sync(FULLM, d > 0.0);
            if (ok) {
                const double rs = rsqrt(d);
                const double l = g[k] * rs;
                if (lane > k && lane < N) Lr[tri + k] = l;
                if (lane == k) dv[k] = rs;
                __syncwarp();
                const double ml = (lane > k) ? l : 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (j > k) g[j] = fma(-ml, Lr[j * (j + 1) / 2 + k], g[j]);     // rows >= N of Lr stay zero
                }
            }
        }
    }
    __syncwarp();
    return ok;
}

// Row r of a Hermitian matrix stored packed lower (row-major) as complex double: element j of the row
// times `scale`, plus `dadd` on the diagonal.  Every array element is written (columns >= N get zeros),
// so whatever the arrays held before is dead from here on.
template <int NT>
__device__ __forceinline__ void load_row_d(const double2* __restrict__ Mp, int r, int tri_r, int N, double scale,
                                           double dadd, double (&fr)[NT], double (&fi)[NT]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int jj = min(j, N - 1);
        const bool low = (jj <= r);
        const double2 v = Mp[low ? tri_r + jj : jj * (jj + 1) / 2 + r];
        const double re = scale * v.x + ((jj == r) ? dadd : 0.0);
        const double im = (jj == r) ? 0.0 : (low ? scale * v.y : -scale * v.y);
        fr[j] = (j < N) ? re : 0.0;
        fi[j] = (j < N) ? im : 0.0;
    }
}
// the same from the FP32 copy of C
template <int NT>
__device__ __forceinline__ void load_row_f(const float2* __restrict__ Cf, int r, int tri_r, int N, double scale,
                                           double dadd, double (&fr)[NT], double (&fi)[NT]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int jj = min(j, N - 1);
        const bool low = (jj <= r);
        const float2 v = Cf[low ? tri_r + jj : jj * (jj + 1) / 2 + r];
        const double re = scale * (double)v.x + ((jj == r) ? dadd : 0.0);
        const double im = (jj == r) ? 0.0 : (low ? scale * (double)v.y : -scale * (double)v.y);
        fr[j] = (j < N) ? re : 0.0;
        fi[j] = (j < N) ? im : 0.0;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(MleCfg<NT>::WARPS * 32, MleCfg<NT>::MIN_CTAS) k_mle(const EvdArgs a) {
    typedef MleCfg<NT> Cfg;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.bands;                                   // 2 <= N <= NT <= 32

    // CTA-wide table: window bit index -> (dy, dx)
    short2* s_off = reinterpret_cast<short2*>(s_raw);
    const int WX = 2 * a.Nx + 1, W = WX * (2 * a.Ny + 1), center = a.Ny * WX + a.Nx;
    for (int f = threadIdx.x; f < a.nulong * 32; f += blockDim.x) {
        const int fy = f / WX;
        s_off[f] = (f < W) ? make_short2((short)(fy - a.Ny), (short)(f - fy * WX - a.Nx))
                           : make_short2((short)-30000, (short)-30000);
    }
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int lut_bytes = ((a.nulong * 32 * (int)sizeof(short2)) + 15) & ~15;

    unsigned char* base = s_raw + lut_bytes + (size_t)warp * Cfg::SMEM_PER_WARP;
    double2* Mp = reinterpret_cast<double2*>(base);                       // packed lower: C, later M
    double2* colb = Mp + Cfg::NPACK;                                       // [2][NT] broadcast vectors
    double* Lr = reinterpret_cast<double*>(colb + 2 * NT);                 // packed real factor
    double* dv = Lr + Cfg::NPACK;                                          // [NT] powers / 1 / L(k,k)
    float2* Cf = reinterpret_cast<float2*>(dv + NT);                       // packed lower C, FP32
    float2* zs = Cf + Cfg::NPACK;                                          // [2][NT] staged samples
    int* s_list = reinterpret_cast<int*>(zs + 2 * NT);                     // [64]

    for (int e = lane; e < Cfg::NPACK; e += 32) Lr[e] = 0.0;            // rows >= N are never written again
    __syncwarp();

    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2), ismle = (a.method == 1);
    const bool pl = (a.variant == 1);
    const int BW = a.bandwidth;
    const int NP = a.NP;
    const long npix_block = (long)a.cols * a.lines;
    const int need = pl ? a.min_neighbors : 2;                             // evd.cpp:566 / phase_link.cpp:524
    const int E = N * (N - 1) / 2;
    const int nslot = (E + 31) >> 5;
    const int r = min(lane, N - 1);                                        // my row (lanes >= N shadow the last one)
    const int tri_r = r * (r + 1) / 2;
    const bool live = lane < N;

    // covariance pairs of this lane: slot s <-> pair e = 32 s + lane of the strict upper triangle, row-major
    int pij[Cfg::NSLOT];
#pragma unroll
    for (int s = 0; s < Cfg::NSLOT; ++s) {
        const int e = s * 32 + lane;
        int t = 0, off = 0;
        if (e < E) { while (off + (N - 1 - t) <= e) { off += N - 1 - t; ++t; } }
        pij[s] = (e < E) ? (t | ((t + 1 + e - off) << 8)) : 0;
    }

    // work: the CTA owns a band of BAND rows x a column segment; warps draw pixels from a shared
    // counter in column-major order (windows of concurrently processed pixels overlap in L1, and
    // sentinel pixels, which finish early, do not leave warps idle)
    const int nbands_img = (a.n_lines + Cfg::BAND - 1) / Cfg::BAND;
    const int band = blockIdx.x % nbands_img, seg = blockIdx.x / nbands_img;
    const int seglen = a.tile_pairs;
    const int c0 = seg * seglen, c1 = min(a.cols, c0 + seglen);
    const int row0 = a.first_line + band * Cfg::BAND;
    const int rows = min(Cfg::BAND, a.first_line + a.n_lines - row0);
    const int total = (c1 - c0) * rows;
    unsigned int st_pix = 0, st_fact = 0, st_steps = 0, st_dp = 0, st_cap = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;

#pragma unroll 1
    for (;;) {
        int kdraw = 0;
        if (lane == 0) kdraw = atomicAdd(&s_next, 1);
        kdraw = __shfl_sync(FULLM, kdraw, 0);
        if (kdraw >= total) break;
        const int row = row0 + kdraw % rows;
        const int col = c0 + kdraw / rows;
        const long pg = (long)row * a.cols + col;
        const uint32_t* mwords = a.wts + pg * a.nulong;
        const bool center_on = __all_sync(FULLM, (__ldg(&mwords[center >> 5]) >> (center & 31)) & 1u);

        float tc = 0.f;
        bool have_vec = false;
        double vxr = 0.0, vxi = 0.0;                          // eigenvector component of this lane

        // SHP list of mask words w0, w0 + 1 -> s_list (if store), returns its length
        auto build_list = [&](int w0, bool store) -> int {
            int n = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int w = w0 + c;
                const uint32_t word = (w < a.nulong) ? __ldg(&mwords[w]) : 0u;
                const short2 d = s_off[min(w, a.nulong - 1) * 32 + lane];
                const int yy = row + d.x, xx = col + d.y;
                const bool ok = ((word >> lane) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                const uint32_t V = __ballot_sync(FULLM, ok);
                if (store && ok) s_list[n + __popc(V & lt_mask)] = yy * a.cols + xx;
                n += __popc(V);
            }
            return n;
        };

        int npix = 0;
        if (center_on) {
            for (int w0 = 0; w0 < a.nulong; w0 += 2) npix += build_list(w0, false);
        }
        bool go = center_on && npix >= need;
        if (go && !pl && ismle && npix < N) { tc = -2.f; go = false; }     // rank(C) <= npix < N: lambda_min = 0 (evd.cpp:613-617)

        if (go) {
            ++st_pix;
            // ---------------- covariance (evd.cpp:537-564) --------------------------------
            {
                double ar[Cfg::NSLOT], ai[Cfg::NSLOT];
#pragma unroll
                for (int s = 0; s < Cfg::NSLOT; ++s) { ar[s] = 0.0; ai[s] = 0.0; }
                double pw = 0.0;
                int parity = 0;
#pragma unroll 1
                for (int w0 = 0; w0 < a.nulong; w0 += 2) {
                    __syncwarp();
                    const int n = build_list(w0, true);
                    __syncwarp();
                    float2 znext = make_float2(0.f, 0.f);
                    if (n > 0 && live) znext = __ldg(&a.zpix[(long)s_list[0] * NP + lane]);
#pragma unroll 1
                    for (int c = 0; c < n; ++c) {
                        const float2 z = znext;
                        if (c + 1 < n && live) znext = __ldg(&a.zpix[(long)s_list[c + 1] * NP + lane]);
                        float2* zb = zs + parity * NT;
                        parity ^= 1;
                        if (live) {
                            zb[lane] = z;
                            // |z|^2: float hypot, squared and summed in double (evd.cpp:558)
                            const float hy = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)z.x, (double)z.x),
                                                                         __dmul_rn((double)z.y, (double)z.y)));
                            pw += (double)hy * (double)hy;
                        }
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < Cfg::NSLOT; ++s) {
                            if (s < nslot) {
                                const float2 zi = zb[pij[s] & 0xff], zj = zb[pij[s] >> 8];
                                // z_i * conj(z_j) rounded term by term as libgcc does (evd.cpp:557)
                                const float pr = __fadd_rn(__fmul_rn(zi.x, zj.x), __fmul_rn(zi.y, zj.y));
                                const float pi = __fsub_rn(__fmul_rn(zi.y, zj.x), __fmul_rn(zi.x, zj.y));
                                ar[s] += (double)pr; ai[s] += (double)pi;
                            }
                        }
                    }
                }
                // coherence (evd.cpp:569-582): C_ij = sum / sqrt(P_i P_j), packed lower (row j, column i)
                __syncwarp();
                if (live) dv[lane] = pw;
                __syncwarp();
#pragma unroll
                for (int s = 0; s < Cfg::NSLOT; ++s) {
                    if (s * 32 + lane < E) {
                        const int i = pij[s] & 0xff, j = pij[s] >> 8;
                        const double den = sqrt(dv[i] * dv[j]);
                        const double cx = ar[s] / den, cy = ai[s] / den;
                        Mp[j * (j + 1) / 2 + i] = make_double2(cx, -cy);
                        Cf[j * (j + 1) / 2 + i] = make_float2((float)cx, -(float)cy);
                    }
                }
                if (live) { Mp[tri_r + r] = make_double2(1.0, 0.0); Cf[tri_r + r] = make_float2(1.f, 0.f); }
                __syncwarp();
            }

            double fr[NT], fi[NT];
            double rs_own = 1.0;
            bool failed = false, run_evd = false, c_in_mp = true;

            // ---------------- gate 1: lambda_min(C) >= 1e-6 (evd.cpp:608-617) ---------------
            if (!pl) {
                load_row_d<NT>(Mp, r, tri_r, N, 1.0, -1.0e-6, fr, fi);
                ++st_fact;
                if (!chol_c_reg<NT>(fr, fi, rs_own, colb, N, lane)) { tc = -2.f; failed = true; }
            }
            // ---------------- |C|, gate 2, inverse (evd.cpp:619-653) ------------------------
            double dmax = 0.0;
            if (!failed) {
                ++st_dp;
                double xv[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) xv[j] = 0.0;
                bool have_inv = false;
                double shift = 0.0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    double g[NT];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int jj = min(j, N - 1);
                        const bool low = (jj <= r);
                        const double2 v = Mp[low ? tri_r + jj : jj * (jj + 1) / 2 + r];
                        const double m = (jj == r) ? 1.0 - shift : sqrt(fma(v.x, v.x, v.y * v.y));
                        g[j] = (j < N) ? m : 0.0;
                    }
                    __syncwarp();
                    const bool okr = chol_r_reg<NT>(g, Lr, dv, N, lane);
                    if (pass == 1) { if (!okr) { tc = -4.f; failed = true; } break; }
                    if (!okr) {                                   // |C| itself is not positive definite
                        if (pl) run_evd = true; else { tc = -4.f; failed = true; }
                        break;
                    }
                    // lane c: column c of inv(|C|) = (L L^T)^-1 e_c, private substitutions
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        if (i < N) {
                            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                            for (int m = 0; m < NT; ++m) { if (m < i) s = fma(-Lr[i * (i + 1) / 2 + m], xv[m], s); }
                            xv[i] = s * dv[i];
                        }
                    }
#pragma unroll
                    for (int i = NT - 1; i >= 0; --i) {
                        if (i < N) {
                            double s = xv[i];
#pragma unroll
                            for (int m = 0; m < NT; ++m) { if (m > i) s = fma(-Lr[m * (m + 1) / 2 + i], xv[m], s); }
                            xv[i] = s * dv[i];
                        }
                    }
                    have_inv = true;
                    if (pl) break;
                    // gate 2: lambda_min(|C|) = 1 / lambda_max(inv); max diagonal <= lambda_max <= max row sum
                    double rowsum = 0.0, dg = 0.0;
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (j < N) { rowsum += fabs(xv[j]); if (j == lane) dg = xv[j]; } }
                    if (!live) { rowsum = 0.0; dg = 0.0; }
                    const double mrow = wmax(rowsum), mdiag = wmax(dg);
                    if (__all_sync(FULLM, mrow <= 0.999e6)) break;     // passes for certain
                    if (__all_sync(FULLM, mdiag >= 1.001e6)) { tc = -4.f; failed = true; break; }
                    shift = 1.0e-6;                               // in between: decide by factorising |C| - 1e-6 I
                }
                // ---------------- M = inv(|C|) o C (evd.cpp:655-657): every lane rewrites the lower part
                // of its own row in place (nobody else touches those entries)
                if (!failed && !run_evd && have_inv) {
                    if (live) {
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            if (j <= lane) {
                                const double2 v = Mp[tri_r + j];
                                Mp[tri_r + j] = make_double2(xv[j] * v.x, (j == lane) ? 0.0 : xv[j] * v.y);
                                if (j == lane) dmax = xv[j];
                            }
                        }
                    }
                    c_in_mp = false;
                    __syncwarp();
                    dmax = wmax(dmax);
                }
            }

            // ---------------- eigen solves ------------------------------------------------------
            // stage 0: smallest eigenpair of M (MLE).  stage 1 (phase_link only): dominant eigenvector of C
            // (phase_link.cpp:586-600) by FP64 power iteration with momentum, and if that stalls by the
            // same certified inverse iteration applied to N I - C.
#pragma unroll 1
            for (int stage = 0; stage < 2; ++stage) {
                bool do_inverse_iteration = false;
                double xr = 0.0, xi = 0.0, scale_d = 1.0;
                if (stage == 0) {
                    if (failed || run_evd) continue;
                    // start vector: phases of the middle column of C (the MLE phases are close to them)
                    const int km = N >> 1;
                    const float2 cst = Cf[(r >= km) ? tri_r + km : km * (km + 1) / 2 + r];
                    xr = cst.x; xi = (r >= km) ? cst.y : -cst.y;
                    const double m2 = xr * xr + xi * xi;
                    if (m2 > 0.0) { const double s = rsqrt(m2); xr *= s; xi *= s; } else { xr = 1.0; xi = 0.0; }
                    { const double s = rsqrt((double)N); xr *= s; xi *= s; }
                    if (!live) { xr = 0.0; xi = 0.0; }
                    scale_d = dmax;
                    do_inverse_iteration = true;
                } else {
                    if (failed || !run_evd) break;
                    if (c_in_mp) load_row_d<NT>(Mp, r, tri_r, N, 1.0, 0.0, fr, fi);
                    else load_row_f<NT>(Cf, r, tri_r, N, 1.0, 0.0, fr, fi);
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (j == k0) { xr = fr[j]; xi = fi[j]; } }     // start: column k0 of C
                    if (!live) { xr = 0.0; xi = 0.0; }
                    double xpr = 0.0, xpi = 0.0;
                    { const double s = rsqrt(wsum(xr * xr + xi * xi)); xr *= s; xi *= s; }
                    double lam = 1.0, beta = 0.0, rho_prev = -1.0;
                    int next_chk = 2, gap = 2;
                    bool got = false;
#pragma unroll 1
                    for (int it = 0; it < 400; ++it) {
                        double2* cb = colb + (it & 1) * NT;
                        if (live) cb[lane] = make_double2(xr, xi);
                        __syncwarp();
                        double yr = 0.0, yi = 0.0;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            if (j < N) {
                                const double2 v = cb[j];
                                yr = fma(fr[j], v.x, yr); yr = fma(-fi[j], v.y, yr);
                                yi = fma(fr[j], v.y, yi); yi = fma(fi[j], v.x, yi);
                            }
                        }
                        if (!live) { yr = 0.0; yi = 0.0; }
                        if (it == next_chk) {
                            const double xy = wsum(xr * yr + xi * yi), xx = wsum(xr * xr + xi * xi);
                            lam = xy / xx;
                            const double rx = yr - lam * xr, ry = yi - lam * xi;
                            const double r2 = wsum(rx * rx + ry * ry), y2 = wsum(yr * yr + yi * yi);
                            const double rho2 = r2 / (lam * lam * xx);
                            if (__all_sync(FULLM, rho2 <= 1.0e-18)) {   // one more plain step, then done
                                const double s = rsqrt(y2);
                                xr = yr * s; xi = yi * s; got = true;
                                break;
                            }
                            if (__all_sync(FULLM, rho_prev > 0.0 && rho2 < rho_prev)) {
                                if (beta == 0.0) {
                                    const double rr = pow(rho2 / rho_prev, 0.5 / (double)gap);
                                    beta = fmin(0.575 * rr * 0.575 * rr, 0.2);
                                }
                            } else if (rho_prev > 0.0) beta *= 0.5;
                            rho_prev = rho2;
                            next_chk = it + gap;
                            const double s = rsqrt(xx), il = 1.0 / lam;
                            const double nr = (yr * il - beta * xpr) * s, ni = (yi * il - beta * xpi) * s;
                            xpr = xr * s; xpi = xi * s; xr = nr; xi = ni;
                        } else if (it < 2) {                       // lambda still unknown: plain normalised steps
                            const double s = rsqrt(wsum(yr * yr + yi * yi));
                            xpr = 0.0; xpi = 0.0; xr = yr * s; xi = yi * s;
                        } else {
                            const double il = 1.0 / lam;
                            const double nr = yr * il - beta * xpr, ni = yi * il - beta * xpi;
                            xpr = xr; xpi = xi; xr = nr; xi = ni;
                        }
                    }
                    if (got) { vxr = xr; vxi = xi; have_vec = true; break; }
                    ++st_cap;                                      // slow or stalled: certified inverse iteration on N I - C
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (j == k0) { xr = fr[j]; xi = fi[j]; } }
                    if (!live) { xr = 0.0; xi = 0.0; }
                    { const double s = rsqrt(wsum(xr * xr + xi * xi)); xr *= s; xi *= s; }
                    scale_d = (double)N;
                    do_inverse_iteration = true;
                }
                if (!do_inverse_iteration) continue;

                // -------- certified inverse iteration: smallest eigenpair of
                //          stage 0: M (packed in Mp)      stage 1: N I - C
                // pending: 0 first shift (just below zero), 1 proposal, 2 retreat, 3 restore
                int pending = 0, attempts = 0, since = 0, it = 0;
                double back = 1.0e-12 * scale_d, sig = -back, sigma_ok = 0.0, hi = CUDART_INF;
                double rho = 0.0, res_prev = -1.0, retreat = 0.0;
                bool solved = false, gave_up = false;
#pragma unroll 1
                while (!solved && !gave_up) {
                    if (stage == 0) load_row_d<NT>(Mp, r, tri_r, N, 1.0, -sig, fr, fi);
                    else if (c_in_mp) load_row_d<NT>(Mp, r, tri_r, N, -1.0, (double)N - sig, fr, fi);   // N I - C: diagonal N - 1
                    else load_row_f<NT>(Cf, r, tri_r, N, -1.0, (double)N - sig, fr, fi);
                    ++st_fact;
                    const bool ok = chol_c_reg<NT>(fr, fi, rs_own, colb, N, lane);
                    if (!ok) {
                        if (pending == 0) { if (++attempts >= 6) gave_up = true; else { back *= 1.0e3; sig = -back; } continue; }
                        hi = fmin(hi, sig);
                        if (pending == 1) {                         // the proposal was too bold: the classic rho - 2 res, or halfway
                            double mid = retreat;
                            if (__any_sync(FULLM, !(mid > sigma_ok) || !(mid < sig))) mid = 0.5 * (sigma_ok + sig);
                            sig = mid; pending = 2; continue;
                        }
                        if (pending == 2) { sig = sigma_ok; pending = 3; continue; }
                        gave_up = true; continue;
                    }
                    sigma_ok = sig;
                    since = 0; res_prev = -1.0;
                    bool reshift = false;
#pragma unroll 1
                    while (!reshift && !solved) {
                        if (it >= 200) { ++st_cap; solved = true; break; }      // iteration cap: best vector so far
                        ++it; ++st_steps;
                        const double x0r = xr, x0i = xi;
                        solve_c_reg<NT>(fr, fi, rs_own, N, lane, xr, xi);
                        if (!live) { xr = 0.0; xi = 0.0; }
                        double n2 = xr * xr + xi * xi;
                        double cr = xr * x0r + xi * x0i;                        // y^H x
                        double ci = xr * x0i - xi * x0r;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            n2 += __shfl_xor_sync(FULLM, n2, o);
                            cr += __shfl_xor_sync(FULLM, cr, o);
                            ci += __shfl_xor_sync(FULLM, ci, o);
                        }
                        const double inv_n = rsqrt(n2);
                        xr *= inv_n; xi *= inv_n;                              // y^ = y / |y|
                        cr *= inv_n; ci *= inv_n;                              // c = y^^H x
                        rho = sigma_ok + cr * inv_n;                           // M y^ = sigma y^ + x / |y|
                        const double rr = x0r - (cr * xr - ci * xi), ri = x0i - (cr * xi + ci * xr);
                        const double res = sqrt(wsum(rr * rr + ri * ri)) * inv_n;   // |(x - c y^)| / |y|
                        if (__all_sync(FULLM, res <= 1.0e-11 * scale_d)) { solved = true; break; }
                        ++since;
                        if (__all_sync(FULLM, since >= 2 && res_prev > 0.0 && res < res_prev && res > 1.0e-8 * scale_d)) {
                            // Kato-Temple: lambda_min >= rho - res^2 / (lambda_2 - rho), the gap from the decay
                            // q = (lambda_1 - sigma) / (lambda_2 - sigma) seen over the last step
                            const double q = res / res_prev;
                            const double gapl = (rho - sigma_ok) * (1.0 / q - 1.0);
                            double dlt = fmax(2.0 * res * res / fmax(gapl, 1.0e-300), 1.0e-10 * scale_d);
                            dlt = fmin(dlt, 2.0 * res);
                            const double p = rho - dlt;
                            if (__all_sync(FULLM, p > sigma_ok + 0.25 * (rho - sigma_ok) && p < hi)) {
                                sig = p; pending = 1; retreat = rho - 2.0 * res; reshift = true;
                            }
                        }
                        res_prev = res;
                    }
                }
                if (stage == 0) {
                    if (gave_up) { if (pl) run_evd = true; else { tc = -6.f; failed = true; } }
                    else if (!pl && __all_sync(FULLM, rho < 1.0e-6)) { tc = -7.f; failed = true; }
                    else { vxr = xr; vxi = xi; have_vec = true; }
                } else {
                    if (gave_up) tc = -6.f; else { vxr = xr; vxi = xi; have_vec = true; }
                }
            }
        }

        // -------- phase reference, compression, temporal coherence (evd.cpp:738-786) --------
        float2 o = make_float2(0.f, 0.f);
        float2 cmp = make_float2(0.f, 0.f);
        if (have_vec) {
            // rotate in double so that the reference component is real positive
            const double refr = __shfl_sync(FULLM, vxr, k0), refi = __shfl_sync(FULLM, vxi, k0);
            const double rn = 1.0 / fmax(sqrt(refr * refr + refi * refi), 1e-300);
            const double ur = (vxr * refr + vxi * refi) * rn, ui = (vxi * refr - vxr * refi) * rn;
            float ux = (float)ur, uy = (float)ui;
            const float m = sqrtf(ux * ux + uy * uy);
            if (m == 0.f) { ux = 1.f; uy = 0.f; }              // arg(0) = 0 in the reference
            else { ux /= m; uy /= m; }
            if (lane == k0) { ux = 1.f; uy = 0.f; }
            if (live) o = make_float2(ux, uy);
            float cr = 0.f, cim = 0.f;
            if (live && lane >= k0) {
                const float2 z = __ldg(&a.zpix[pg * NP + lane]);
                cr = z.x * ux + z.y * uy;                        // z * conj(o)
                cim = z.y * ux - z.x * uy;
            }
            cr = wsumf(cr); cim = wsumf(cim);
            const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
            cmp = make_float2(cr * invn, cim * invn);
            __syncwarp();
            if (live) zs[lane] = o;
            __syncwarp();
            // lane j sums the pairs (i < j): e(i,j) conj(o_i) o_j, e = C_ij / |C_ij|, C_ij = conj(Cf(j,i))
            float sr = 0.f, si = 0.f;
            int cnt = 0;
            if (live) {
                for (int i = 0; i < lane; ++i) {
                    if (isstbas && (lane - i) > BW) continue;
                    const float2 c = Cf[tri_r + i];
                    const float mm = sqrtf(c.x * c.x + c.y * c.y);
                    float ex = 1.f, ey = 0.f;
                    if (mm > 0.f) { ex = c.x / mm; ey = -c.y / mm; }
                    const float2 oi = zs[i];
                    const float tx = ex * oi.x + ey * oi.y, ty = ey * oi.x - ex * oi.y;     // e * conj(o_i)
                    sr += tx * o.x - ty * o.y;
                    si += tx * o.y + ty * o.x;
                    ++cnt;
                }
            }
            sr = wsumf(sr); si = wsumf(si);
            cnt = __reduce_add_sync(FULLM, cnt);
            tc = sqrtf(sr * sr + si * si) / (float)cnt;
        }
        if (live) a.out[(long)lane * npix_block + pg] = o;
        if (lane == 0) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
        __syncwarp();
    }
    if (a.stats && lane == 0) {
        atomicAdd(&a.stats[0], (unsigned long long)st_pix);
        atomicAdd(&a.stats[1], (unsigned long long)st_steps);
        atomicAdd(&a.stats[2], (unsigned long long)st_dp);
        atomicAdd(&a.stats[3], (unsigned long long)st_cap);
        atomicAdd(&a.stats[4], (unsigned long long)st_fact);
    }
}

// ---------------------------------------------------------------------------------------
template <int NT>
static cudaError_t launch_mle_t(const EvdArgs& a, cudaStream_t st) {
    typedef MleCfg<NT> Cfg;
    const size_t lut = ((size_t)a.nulong * 32 * sizeof(short2) + 15) & ~(size_t)15;
    const size_t smem = lut + (size_t)Cfg::SMEM_PER_WARP * Cfg::WARPS;
    cudaError_t e = cudaFuncSetAttribute(k_mle<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // bands of BAND rows x column segments, >= 16 CTAs per SM so the tail evens out
    const int nbands = (a.n_lines + Cfg::BAND - 1) / Cfg::BAND;
    int nseg = (nsm * 16 + nbands - 1) / nbands;
    if (nseg < 1) nseg = 1;
    int seglen = (a.cols + nseg - 1) / nseg;
    if (seglen < 16) seglen = 16;
    nseg = (a.cols + seglen - 1) / seglen;
    EvdArgs b = a;
    b.tile_pairs = seglen;
    k_mle<NT><<<(unsigned)(nbands * nseg), Cfg::WARPS * 32, smem, st>>>(b);
    return cudaGetLastError();
}

// register-row order: the smallest instantiated order >= bands (0 = not covered by this kernel)
int evd_mle_order(int bands) {
    static const int orders[] = {12, 16, 20, 24, 28, 32};
    if (bands < 2) return 0;
    for (int o : orders) if (bands <= o) return o;
    return 0;
}

cudaError_t launch_evd_mle(const EvdArgs& a, cudaStream_t st) {
    switch (evd_mle_order(a.bands)) {
#ifdef FRINGE_MLE_ONLY                                   // development builds: one instantiation
        case FRINGE_MLE_ONLY: return launch_mle_t<FRINGE_MLE_ONLY>(a, st);
#else
        case 12: return launch_mle_t<12>(a, st);
        case 16: return launch_mle_t<16>(a, st);
        case 20: return launch_mle_t<20>(a, st);
        case 24: return launch_mle_t<24>(a, st);
        case 28: return launch_mle_t<28>(a, st);
        case 32: return launch_mle_t<32>(a, st);
#endif
    }
    return cudaErrorInvalidValue;
}

}  // namespace fringe
