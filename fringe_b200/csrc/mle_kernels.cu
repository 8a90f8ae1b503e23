// FP64 maximum-likelihood (MLE / EMI) and phase_link kernel for bands <= 32 (sm_100a).
//
// What the reference does per pixel (src/evd/evd.cpp:537-688 with method MLE -- the default,
// evd.hpp:66 -- and src/phase_link/phase_link.cpp:503-618): masked sample covariance (float
// products, double sums), coherence C, two positive-definiteness gates through LAPACK zheevr,
// inv(|C|) through zpotrf + zpotri, Hadamard product M = inv(|C|) o C, smallest eigenpair of M
// (zheevr), sentinels -2 / -4 / -5 / -6 / -7 on the way (phase_link: EVD fallback instead).
//
// How it is done here: one warp per pixel, lane = matrix row.  The design is shaped by two
// measurements on the B200 (profiles/r2_mle_*): the round-1 kernel ran at 12 % warps active with
// three-way bank conflicts in its left-looking factorisations, and a first rewrite that kept the
// matrices in registers with fully unrolled factorisations (csrc/experimental/
// mle_register_rows_unrolled.cu) was no faster because its 300 kB of straight-line code thrashed
// the instruction cache (85 % of all stall samples "no instruction", 10 % issue slots used).
// So: every quadratic-size loop is rolled and runs on packed matrices in shared memory; only the
// two triangular sweeps of the inverse iteration, whose code is linear in N, are unrolled and run
// from registers.
//   covariance    pairs (i < j) dealt round-robin to the 32 lanes; per SHP the N samples are
//                 staged once in shared memory, every lane forms its pairs' products exactly as
//                 libgcc's complex multiply rounds them and adds them in double (double sums of
//                 float products are exact, so the order of the SHPs does not matter)
//   Cholesky      right-looking on the packed column-major lower triangle (columns contiguous:
//                 lanes read consecutive addresses), panels of 4 columns: the trailing update
//                 reads and writes every entry once per panel and applies a rank-4 update from
//                 registers (6 shared-memory operations per 16 DFMA)
//   solves        the factor's row and column of a lane are loaded into registers once per
//                 factorisation; forward / backward substitution broadcast the pivot value by
//                 shuffles, branch-free (a lane's coefficient is zero on steps that do not concern it)
//   gates         lambda_min(C) >= 1e-6 as Cholesky of C - 1e-6 I; pixels with fewer SHPs than
//                 bands are rank deficient and take the -2 sentinel without any arithmetic;
//                 lambda_min(|C|) >= 1e-6 is read off inv(|C|) (max diagonal <= 1 / lambda_min <=
//                 max row sum) and only decided by a second factorisation in between
//   inverse       real Cholesky, then lane c solves L L^T x = e_c privately (uniform addresses
//                 into the factor: broadcasts)
//   eigen         inverse iteration with Cholesky-certified shifts (a successful factorisation
//                 of M - sigma I proves sigma < lambda_min, so the iteration cannot lock onto a
//                 wrong eigenvalue).  Rayleigh quotient and residual come for free from the
//                 solve (M y = sigma y + x), shifts follow the Kato-Temple bound with the gap
//                 estimated from the observed decay; start vector = phases of the middle column
//                 of C (offline replay: 3.0 factorisations + 5.2 solves per pixel against
//                 4.4 + 8.9 for the round-1 policy)
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
    // one CTA per SM; warps limited by registers (65536 / threads) and shared memory (227 kB)
#ifdef FRINGE_MLE_WARPS
    static constexpr int WARPS = FRINGE_MLE_WARPS;
#else
    static constexpr int WARPS = NT <= 16 ? 16 : (NT <= 24 ? 12 : (NT <= 28 ? 8 : 7));
#endif
    static constexpr int MIN_CTAS = 1;
    static constexpr int MAX_WORK = 1024;                    // pixels per CTA (band rows x segment columns)
    static constexpr int BAND = 4;                            // rows of a CTA's pixel band
    static constexpr int NPAIR = NT * (NT - 1) / 2;
    static constexpr int NSLOT = (NPAIR + 31) / 32;           // covariance pairs per lane
    static constexpr int NPACK = NT * (NT + 1) / 2;           // packed lower triangle
    // per warp: Mp, F complex double packed | Lr real packed | rsv, dv NT doubles | Cf float2 packed |
    //           zs 2 x NT float2 | list 64 ints
    static constexpr int SMEM_PER_WARP = 16 * NPACK + 16 * NPACK + 8 * NPACK + 16 * NT + 8 * NPACK + 16 * NT + 256 + 64 * NT;   // + pan [NT][4] complex
};

// Packed column-major lower triangle with the static pitch NT (the instantiation's order; rows and
// columns >= N are padding): column j starts at coff(j), element (i, j), i >= j, sits at coff(j) + i - j.
// Lanes that walk down a column read consecutive addresses, and for an unrolled j the offset is an
// immediate.
template <int NT>
__device__ __forceinline__ int coff(int j) { return (j * (2 * NT - j + 1)) >> 1; }

// one predicated shared-memory store (the compiler turns `if (p) *q = v` in these loops into a branch)
__device__ __forceinline__ void sts_pred(double2* q, double2 v, bool pred) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(q);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.shared.v2.f64 [%0], {%1, %2};\n\t}"
                 :: "r"(addr), "d"(v.x), "d"(v.y), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts_pred(double* q, double v, bool pred) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(q);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}"
                 :: "r"(addr), "d"(v), "r"((int)pred) : "memory");
}

// ---------------------------------------------------------------------------------------
// Complex Cholesky F = L L^H in shared memory, in place, lane = row; rsv[k] = 1 / L(k,k).
// Panels of 4 columns: a lane takes its row's 4 panel entries into registers, the panel is factored
// there (pivots and the L(j,k) of the other panel columns arrive by shuffles: no shared-memory
// traffic, no barriers), written back, and a row-major copy `pan` feeds the broadcast operands of the
// rank-4 trailing update.  Returns false (warp-uniform) on a non-positive pivot, the failure LAPACK
// zpotrf reports.
// ---------------------------------------------------------------------------------------
template <int NT>
__device__ __noinline__ bool chol_c_smem(double2* F, double2* pan, double* rsv, int N, int lane) {
    const int row = min(lane, N - 1);
    const bool live = lane < N;
#pragma unroll 1
    for (int kb = 0; kb < N; kb += 4) {
        const int kend = min(kb + 4, N);
        double2 p[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = min(kb + t, N - 1);
            p[t] = F[coff<NT>(k) + max(row, k) - k];
        }
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = kb + t;
            if (k < N) {
                const double d = __shfl_sync(FULLM, p[t].x, k);
                ok = ok && __all_sync(FULLM, d > 0.0);
                const double rs = rsqrt(d);
                if (lane == k) rsv[k] = rs;
                const bool act = lane > k && live;
                p[t].x = act ? p[t].x * rs : (lane == k ? d * rs : 0.0);
                p[t].y = act ? p[t].y * rs : 0.0;
#pragma unroll
                for (int u = t + 1; u < 4; ++u) {
                    if (kb + u < N) {                          // a(i,j) -= l(i,k) conj(l(j,k)), j = kb + u
                        const double vr = __shfl_sync(FULLM, p[t].x, kb + u), vi = __shfl_sync(FULLM, p[t].y, kb + u);
                        const bool on = lane >= kb + u;
                        const double lr = on ? p[t].x : 0.0, li = on ? p[t].y : 0.0;
                        p[u].x = fma(-lr, vr, p[u].x); p[u].x = fma(-li, vi, p[u].x);
                        p[u].y = fma(-li, vr, p[u].y); p[u].y = fma(lr, vi, p[u].y);
                    }
                }
            }
        }
        if (!ok) return false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = min(kb + t, N - 1);
            sts_pred(&F[coff<NT>(k) + max(row, k) - k], p[t], kb + t < N && lane >= kb + t && live);
        }
        if (kend < N) {                                       // trailing update by the finished panel (always 4 wide here)
            const bool own = lane >= kend && live;
            double2 lp[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                sts_pred(&pan[row * 4 + t], p[t], own);
                lp[t] = own ? p[t] : make_double2(0.0, 0.0);
            }
            __syncwarp();
            int cjm = coff<NT>(kend) - kend;                  // coff(j) - j, advanced with j
#pragma unroll 1
            for (int j = kend; j < N; ++j) {
                double2* q = F + cjm + max(row, j);
                double2 a = *q;
                const double2* pj = pan + j * 4;               // L(j, kb .. kb+3), broadcast
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double2 v = pj[t];
                    a.x = fma(-lp[t].x, v.x, a.x); a.x = fma(-lp[t].y, v.y, a.x);
                    a.y = fma(-lp[t].y, v.x, a.y); a.y = fma(lp[t].x, v.y, a.y);
                }
                sts_pred(q, a, lane >= j && live);
                cjm += NT - j - 1;
            }
        }
        __syncwarp();
    }
    return true;
}

// the same for a real symmetric matrix; dv[k] = 1 / L(k,k); pan is used as [N][4] doubles
template <int NT>
__device__ __noinline__ bool chol_r_smem(double* A, double* pan, double* dv, int N, int lane) {
    const int row = min(lane, N - 1);
    const bool live = lane < N;
#pragma unroll 1
    for (int kb = 0; kb < N; kb += 4) {
        const int kend = min(kb + 4, N);
        double p[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = min(kb + t, N - 1);
            p[t] = A[coff<NT>(k) + max(row, k) - k];
        }
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = kb + t;
            if (k < N) {
                const double d = __shfl_sync(FULLM, p[t], k);
                ok = ok && __all_sync(FULLM, d > 0.0);
                const double rs = rsqrt(d);
                if (lane == k) dv[k] = rs;
                p[t] = (lane > k && live) ? p[t] * rs : (lane == k ? d * rs : 0.0);
#pragma unroll
                for (int u = t + 1; u < 4; ++u) {
                    if (kb + u < N) {
                        const double v = __shfl_sync(FULLM, p[t], kb + u);
                        p[u] = fma(-((lane >= kb + u) ? p[t] : 0.0), v, p[u]);
                    }
                }
            }
        }
        if (!ok) return false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = min(kb + t, N - 1);
            sts_pred(&A[coff<NT>(k) + max(row, k) - k], p[t], kb + t < N && lane >= kb + t && live);
        }
        if (kend < N) {
            const bool own = lane >= kend && live;
            double lp[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                sts_pred(&pan[row * 4 + t], p[t], own);
                lp[t] = own ? p[t] : 0.0;
            }
            __syncwarp();
            int cjm = coff<NT>(kend) - kend;
#pragma unroll 1
            for (int j = kend; j < N; ++j) {
                double* q = A + cjm + max(row, j);
                double a = *q;
                const double* pj = pan + j * 4;
#pragma unroll
                for (int t = 0; t < 4; ++t) a = fma(-lp[t], pj[t], a);
                sts_pred(q, a, lane >= j && live);
                cjm += NT - j - 1;
            }
        }
        __syncwarp();
    }
    return true;
}

// Row `row` of the Hermitian matrix whose lower triangle is packed in P: element j into fr[j] + i fi[j]
// (the diagonal entry is read as stored; columns >= N deliver padding nobody uses).
template <int NT>
__device__ __forceinline__ void load_row(const double2* __restrict__ P, int row, double (&fr)[NT], double (&fi)[NT]) {
    const double2* up = P + coff<NT>(row) - row;           // P(j, row), j > row, at up[j]
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const bool low = (j <= row);
        const double2 v = low ? P[coff<NT>(j) - j + row] : up[j];
        fr[j] = v.x;
        fi[j] = low ? v.y : -v.y;
    }
}

// x <- (L L^H)^-1 x; lane i holds x(i), row i of L in fr/fi[k], k < i, and the conjugate of column i in
// fr/fi[k], k > i (as load_row delivers them from the packed factor), rs_own = 1 / L(i,i).
//   forward   L y = x, column oriented: step k broadcasts y(k) = x(k) rs_k, rows > k subtract L(i,k) y(k)
//   backward  L^H z = y: step k broadcasts z(k) = (y(k) - acc(k)) rs_k, rows < k add conj(L(k,i)) z(k)
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
            const double zr = __shfl_sync(FULLM, (xr - ar) * rs_own, k), zi = __shfl_sync(FULLM, (xi - ai) * rs_own, k);
            // fr/fi[k], k > lane, hold conj(L(k,lane)) already (load_row conjugates the upper part)
            const double cr = (lane < k) ? fr[k] : 0.0, ci = (lane < k) ? fi[k] : 0.0;
            ar = fma(cr, zr, ar); ar = fma(-ci, zi, ar);
            ai = fma(cr, zi, ai); ai = fma(ci, zr, ai);
        }
    }
    xr = (xr - ar) * rs_own; xi = (xi - ai) * rs_own;   // z(i)
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
    __shared__ int s_nwork;
    __shared__ unsigned short s_work[MleCfg<NT>::MAX_WORK];       // pixels of this CTA that reach the solver
    if (threadIdx.x == 0) s_nwork = 0;
    __syncthreads();
    const int lut_bytes = ((a.nulong * 32 * (int)sizeof(short2)) + 15) & ~15;

    unsigned char* base = s_raw + lut_bytes + (size_t)warp * Cfg::SMEM_PER_WARP;
    double2* Mp = reinterpret_cast<double2*>(base);                       // packed lower: C, later M
    double2* F = Mp + Cfg::NPACK;                                          // packed factor; scratch of the inverse
    double* Lr = reinterpret_cast<double*>(F + Cfg::NPACK);                // packed real matrix / factor
    double* rsv = Lr + Cfg::NPACK;                                         // [NT] 1 / L(k,k) of the complex factor
    double* dv = rsv + NT;                                                 // [NT] powers, then 1 / L(k,k) of the real factor
    float2* Cf = reinterpret_cast<float2*>(dv + NT);                       // packed lower C, FP32
    float2* zs = Cf + Cfg::NPACK;                                          // [2][NT] staged samples
    int* s_list = reinterpret_cast<int*>(zs + 2 * NT);                     // [64]
    double2* pan = reinterpret_cast<double2*>(s_list + 64);                // [NT][4] panel rows of the factorisation in progress

    const int k0 = a.mini_stack_count - 1;
    const bool isstbas = (a.method == 2), ismle = (a.method == 1);
    const bool pl = (a.variant == 1);
    const int BW = a.bandwidth;
    const int NP = a.NP;
    const long npix_block = (long)a.cols * a.lines;
    const int need = pl ? a.min_neighbors : 2;                             // evd.cpp:566 / phase_link.cpp:524
    const int E = N * (N - 1) / 2;
    const int nslot = (E + 31) >> 5;
    constexpr int npack = Cfg::NPACK;                                     // elementwise passes run over the padding too
    const int r = min(lane, N - 1);                                        // my row (lanes >= N shadow the last one)
    const int cr_own = coff<NT>(r);
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

    // All warps of the CTA walk through the phases of a pixel together (__syncthreads between them): the
    // instruction working set of the SM is then one phase at a time.  Measured on the version with
    // independent warps: 64 % instruction-cache hit rate and the GPC-level instruction cache at 95 % of
    // its request bandwidth -- the kernel was bound by instruction fetch, not by any arithmetic pipe.
    const int nwarps = blockDim.x >> 5;

    // ---- pass 0: which pixels reach the solver?  The others (mask centre off, too few SHPs, or fewer SHPs
    // than bands under MLE: rank(C) <= npix < N, lambda_min = 0, sentinel -2 of evd.cpp:613-617) get their
    // outputs here, so that no warp idles at the phase barriers below on their behalf.
#pragma unroll 1
    for (int k = warp; k < total; k += nwarps) {
        const int row = row0 + k % rows, col = c0 + k / rows;
        const long pg = (long)row * a.cols + col;
        const uint32_t* mwords = a.wts + pg * a.nulong;
        const bool center_on = __all_sync(FULLM, (__ldg(&mwords[center >> 5]) >> (center & 31)) & 1u);
        int npix = 0;
        if (center_on) {
            for (int w = lane; w < a.nulong * 32; w += 32) {
                const uint32_t word = __ldg(&mwords[w >> 5]);
                const short2 d = s_off[w];
                const int yy = row + d.x, xx = col + d.y;
                const bool ok = ((word >> (w & 31)) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                npix += __popc(__ballot_sync(FULLM, ok));
            }
        }
        const bool rank_deficient = !pl && ismle && npix < N;
        if (center_on && npix >= need && !rank_deficient) {
            if (lane == 0) s_work[atomicAdd(&s_nwork, 1)] = (unsigned short)k;
        } else {
            float code = 0.f;
            if (center_on && npix >= need) {
                // rank deficient: -2, unless some band is zero in every SHP -- then the coherence matrix holds NaNs,
                // LAPACK's zheevr reports failure and the reference writes -1 before it ever looks at the eigenvalue
                // (evd.cpp:608-612).  Happens in the sequential chain, where compressed SLCs of failed pixels are 0.
                bool nonzero = false;
                for (int w = lane; w < a.nulong * 32; w += 32) {
                    const uint32_t word = __ldg(&mwords[w >> 5]);
                    const short2 d = s_off[w];
                    const int yy = row + d.x, xx = col + d.y;
                    uint32_t V = __ballot_sync(FULLM, ((word >> (w & 31)) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols);
                    while (V) {
                        const int src = __ffs(V) - 1;
                        V &= V - 1;
                        const int qy = __shfl_sync(FULLM, yy, src), qx = __shfl_sync(FULLM, xx, src);
                        if (live) { const float2 z = __ldg(&a.zpix[((long)qy * a.cols + qx) * NP + lane]); nonzero = nonzero || z.x != 0.f || z.y != 0.f; }
                    }
                }
                code = __all_sync(FULLM, nonzero || !live) ? -2.f : -1.f;
            }
            if (live) a.out[(long)lane * npix_block + pg] = make_float2(0.f, 0.f);
            if (lane == 0) { a.tcorr[pg] = code; a.comp[pg] = make_float2(0.f, 0.f); }
        }
    }
    __syncthreads();
    const int nwork = s_nwork;

#pragma unroll 1
    for (int kbase = 0; kbase < nwork; kbase += nwarps) {
        const bool mine = kbase + warp < nwork;
        const int kdraw = s_work[min(kbase + warp, nwork - 1)];
        const int row = row0 + kdraw % rows;
        const int col = c0 + kdraw / rows;
        const long pg = (long)row * a.cols + col;
        const uint32_t* mwords = a.wts + pg * a.nulong;
        const bool go = mine;

        float tc = 0.f;
        bool have_vec = false;
        double vxr = 0.0, vxi = 0.0;                          // eigenvector component of this lane
        bool failed = false, run_evd = false, c_in_mp = true, have_inv = false, okr = false;
        double dmax = 0.0;

        // SHP list of mask words w0, w0 + 1 -> s_list, returns its length
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

        // F = scale * Mp + dadd I on the packed triangle
        auto assemble = [&](double scale, double dadd) {
            for (int e = lane; e < npack; e += 32) { const double2 v = Mp[e]; F[e] = make_double2(scale * v.x, scale * v.y); }
            __syncwarp();
            if (live) { const double2 v = F[cr_own]; F[cr_own] = make_double2(v.x + dadd, 0.0); }
            __syncwarp();
        };
        auto fill_abs = [&](double shift) {
            // |C| elementwise on the packed triangle; padding entries are forced finite (they meet zero
            // multipliers in the unrolled substitutions of the inverse)
            __syncwarp();
            for (int e = lane; e < npack; e += 32) {
                const double2 v = Mp[e];
                const double m = sqrt(fma(v.x, v.x, v.y * v.y));
                Lr[e] = (m <= 2.0) ? m : 0.0;
            }
            __syncwarp();
            if (live) Lr[cr_own] = 1.0 - shift;
            __syncwarp();
        };

        // ================= phase 1: covariance (evd.cpp:537-564) =================
        if (go) {
            ++st_pix;
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
                // a band that is zero in every SHP: NaNs in the coherence matrix, LAPACK fails, sentinel -1
                // (evd.cpp:608-612, phase_link.cpp:540-545)
                if (__any_sync(FULLM, live && !(pw > 0.0))) { tc = -1.f; failed = true; }
                // coherence (evd.cpp:569-582): C_ij = sum / sqrt(P_i P_j); stored is C(j,i) = conj(C_ij), j > i
                __syncwarp();
                if (live) dv[lane] = pw;
                __syncwarp();
#pragma unroll
                for (int s = 0; s < Cfg::NSLOT; ++s) {
                    if (s * 32 + lane < E) {
                        const int i = pij[s] & 0xff, j = pij[s] >> 8;
                        const double den = sqrt(dv[i] * dv[j]);
                        const double cx = ar[s] / den, cy = ai[s] / den;
                        const int idx = coff<NT>(i) + j - i;
                        Mp[idx] = make_double2(cx, -cy);
                        Cf[idx] = make_float2((float)cx, -(float)cy);
                    }
                }
                if (live) { Mp[cr_own] = make_double2(1.0, 0.0); Cf[cr_own] = make_float2(1.f, 0.f); }
                __syncwarp();
            }

        }
        __syncthreads();
        // ================= phase 2: gate 1, lambda_min(C) >= 1e-6 (evd.cpp:608-617) =================
        if (go && !pl && !failed) {
            assemble(1.0, -1.0e-6);
            ++st_fact;
            if (!chol_c_smem<NT>(F, pan, rsv, N, lane)) { tc = -2.f; failed = true; }
        }
        // ================= phase 3: |C| and its factor (evd.cpp:619-653); same barrier interval (both are
        // factorisation code) =================
        if (go && !failed) {
            ++st_dp;
            fill_abs(0.0);
            okr = chol_r_smem<NT>(Lr, reinterpret_cast<double*>(pan), dv, N, lane);
            if (!okr) {                                       // |C| itself is not positive definite
                if (pl) run_evd = true; else { tc = -4.f; failed = true; }
            }
        }
        __syncthreads();
        // ================= phase 4: inverse, gate 2, M = inv(|C|) o C =================
        if (go && okr) {
            double xv[NT];                                   // lane c: column c (= row c) of inv(|C|)
#pragma unroll
            for (int j = 0; j < NT; ++j) xv[j] = 0.0;
            // lane c solves L L^T x = e_c privately: the factor is read at uniform addresses (broadcasts,
            // immediate offsets), the vector stays in registers
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                if (i < N) {
                    double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int m = 0; m < NT; ++m) { if (m < i) sacc = fma(-Lr[coff<NT>(m) + i - m], xv[m], sacc); }
                    xv[i] = sacc * dv[i];
                }
            }
#pragma unroll
            for (int i = NT - 1; i >= 0; --i) {
                if (i < N) {
                    double sacc = xv[i];
#pragma unroll
                    for (int m = 0; m < NT; ++m) { if (m > i) sacc = fma(-Lr[coff<NT>(i) + m - i], xv[m], sacc); }   // xv[m >= N] = 0
                    xv[i] = sacc * dv[i];
                }
            }
            have_inv = true;
            if (!pl) {
                // gate 2: lambda_min(|C|) = 1 / lambda_max(inv); max diagonal <= lambda_max <= max row sum;
                // in between the two bounds a factorisation of |C| - 1e-6 I decides
                double rowsum = 0.0, dg = 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j) { rowsum += fabs(xv[j]); dg = (j == lane) ? xv[j] : dg; }
                if (!live) { rowsum = 0.0; dg = 0.0; }
                const double mrow = wmax(rowsum), mdiag = wmax(dg);
                if (!__all_sync(FULLM, mrow <= 0.999e6)) {
                    if (__all_sync(FULLM, mdiag >= 1.001e6)) { tc = -4.f; failed = true; }
                    else {
                        fill_abs(1.0e-6);
                        if (!chol_r_smem<NT>(Lr, reinterpret_cast<double*>(pan), dv, N, lane)) { tc = -4.f; failed = true; }
                    }
                }
            }
            // every lane rewrites the lower part of its own row of Mp in place (evd.cpp:655-657)
            if (!failed) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (j < N) {
                        double2* q = Mp + coff<NT>(j) - j + max(r, j);
                        const double2 v = *q;
                        sts_pred(q, make_double2(xv[j] * v.x, (j == lane) ? 0.0 : xv[j] * v.y), j <= lane && live);
                        dmax = (j == lane) ? xv[j] : dmax;
                    }
                }
                c_in_mp = false;
                __syncwarp();
                dmax = wmax(live ? dmax : 0.0);
            }
        }
        __syncthreads();
        // ================= phase 5: eigen solves =================
        if (go) {
            // ---------------- eigen solves ------------------------------------------------------
            // stage 0: smallest eigenpair of M (MLE).  stage 1 (phase_link only): dominant eigenvector of C
            // (phase_link.cpp:586-600) by FP64 power iteration with momentum, and if that stalls by the
            // same certified inverse iteration applied to N I - C.
            double fr[NT], fi[NT];
#pragma unroll 1
            for (int stage = 0; stage < 2; ++stage) {
                bool do_inverse_iteration = false;
                double xr = 0.0, xi = 0.0, scale_d = 1.0;
                if (stage == 0) {
                    if (failed || run_evd) continue;
                    // start vector: phases of the middle column of C (the MLE phases are close to them)
                    const int km = N >> 1;
                    const float2 cst = Cf[(r >= km) ? coff<NT>(km) + r - km : cr_own + km - r];
                    xr = cst.x; xi = (r >= km) ? cst.y : -cst.y;
                    const double m2 = xr * xr + xi * xi;
                    if (m2 > 0.0) { const double s = rsqrt(m2); xr *= s; xi *= s; } else { xr = 1.0; xi = 0.0; }
                    { const double s = rsqrt((double)N); xr *= s; xi *= s; }
                    if (!live) { xr = 0.0; xi = 0.0; }
                    scale_d = dmax;
                    do_inverse_iteration = true;
                } else {
                    if (failed || !run_evd) break;
                    if (!c_in_mp) {                              // Mp holds M by now: C back from its FP32 copy
                        __syncwarp();
                        for (int e = lane; e < npack; e += 32) { const float2 v = Cf[e]; Mp[e] = make_double2((double)v.x, (double)v.y); }
                        __syncwarp();
                        c_in_mp = true;
                    }
                    load_row<NT>(Mp, r, fr, fi);
                    const double2 ck0 = Mp[(r >= k0) ? coff<NT>(k0) + r - k0 : cr_own + k0 - r];   // start: column k0 of C
                    xr = ck0.x; xi = (r >= k0) ? ck0.y : -ck0.y;
                    if (!live) { xr = 0.0; xi = 0.0; }
                    double xpr = 0.0, xpi = 0.0;
                    { const double s = rsqrt(wsum(xr * xr + xi * xi)); xr *= s; xi *= s; }
                    double lam = 1.0, beta = 0.0, rho_prev = -1.0;
                    int next_chk = 2, gap = 2;
                    bool got = false;
                    double2* cb = F;                             // [2][NT] broadcast vectors
#pragma unroll 1
                    for (int it = 0; it < 400; ++it) {
                        double2* cbi = cb + (it & 1) * NT;
                        if (live) cbi[lane] = make_double2(xr, xi);
                        __syncwarp();
                        double yr = 0.0, yi = 0.0;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            if (j < N) {
                                const double2 v = cbi[j];
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
                    __syncwarp();
                    if (got) { vxr = xr; vxi = xi; have_vec = true; break; }
                    ++st_cap;                                      // slow or stalled: certified inverse iteration on N I - C
                    xr = ck0.x; xi = (r >= k0) ? ck0.y : -ck0.y;
                    if (!live) { xr = 0.0; xi = 0.0; }
                    { const double s = rsqrt(wsum(xr * xr + xi * xi)); xr *= s; xi *= s; }
                    scale_d = (double)N;
                    do_inverse_iteration = true;
                }
                if (!do_inverse_iteration) continue;

                // -------- certified inverse iteration: smallest eigenpair of
                //          stage 0: M (packed in Mp)      stage 1: N I - C (C packed in Mp)
                // pending: 0 first shift (just below zero), 1 proposal, 2 retreat, 3 restore
                int pending = 0, attempts = 0, since = 0, it = 0;
                double back = 1.0e-12 * scale_d, sig = -back, sigma_ok = 0.0, hi = CUDART_INF;
                double rho = 0.0, res_prev = -1.0, retreat = 0.0;
                bool solved = false, gave_up = false;
#pragma unroll 1
                while (!solved && !gave_up) {
                    if (stage == 0) assemble(1.0, -sig);
                    else assemble(-1.0, (double)N - sig);          // N I - C: diagonal N - 1
                    ++st_fact;
                    const bool ok = chol_c_smem<NT>(F, pan, rsv, N, lane);
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
                    load_row<NT>(F, r, fr, fi);
                    const double rs_own = rsv[r];
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

        __syncthreads();
        // ================= phase 6: phase reference, compression, temporal coherence (evd.cpp:738-786) =================
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
                    const float2 c = Cf[coff<NT>(i) + lane - i];
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
        if (live && mine) a.out[(long)lane * npix_block + pg] = o;
        if (lane == 0 && mine) { a.tcorr[pg] = tc; a.comp[pg] = cmp; }
        __syncthreads();
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
    if (seglen * Cfg::BAND > Cfg::MAX_WORK) seglen = Cfg::MAX_WORK / Cfg::BAND;
    nseg = (a.cols + seglen - 1) / seglen;
    EvdArgs b = a;
    b.tile_pairs = seglen;
    k_mle<NT><<<(unsigned)(nbands * nseg), Cfg::WARPS * 32, smem, st>>>(b);
    return cudaGetLastError();
}

// order of the instantiation that serves `bands` (0 = not covered by this kernel)
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
