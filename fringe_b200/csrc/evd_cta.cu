// phase_link for 32 < bands <= 104: one CTA per pixel, the N x N coherence matrix held on chip.
//
// What the reference does per pixel (src/phase_link/phase_link.cpp:479-666): sample covariance over the SHPs (float
// products, double sums), coherence C, zheevr as a NaN probe (:540-545), zpotrf / zpotri of |C| and the smallest
// eigenvector of inv(|C|) o C (:547-584), and -- whenever |C| is not positive definite -- the dominant eigenvector of C
// by zheevr (:586-600).  With more dates than SHPs (configs[2]: 100 dates, ~33 SHPs in an 11x5 window) C is singular and
// |C| indefinite, so practically every pixel ends in that last branch.
//
// The warp-per-pixel kernel (k_evd<HR,DP,GS>) keeps a 485 kB FP64 workspace per warp in global memory at N = 100 and
// re-reads the neighbours' samples once per chunk of 512 matrix entries.  Here a CTA owns the pixel:
//   * the SHPs' sample vectors are staged once in shared memory; thread t owns one 4x4 tile of the upper triangle of the
//     covariance in registers (16 complex double accumulators) and reads 8 samples per SHP for 16 entries; products are
//     rounded the way libgcc's complex multiply rounds them, summed in double in raster order (evd.cpp:557), as the
//     reference does -- bit for bit the sums of k_evd<.,true,.>;
//   * C (FP64, full Hermitian) and |C| (packed lower triangle) go to shared memory; an LDL^T sweep of |C| with one
//     barrier per column decides positive definiteness.  PD pixels (the MLE branch proper) are appended to a work list
//     and solved by the warp-per-pixel kernel afterwards -- rare where this kernel is chosen;
//   * the fallback's dominant eigenpair: FP64 power iteration with the heavy-ball momentum rule of power_iteration_dp
//     (evd_kernels.cu), four threads per matrix row, **the row strip of each thread in registers** (N = 100: 25 complex
//     doubles), so an iteration reads only the vector from shared memory: it runs at the FP64 pipe's rate
//     (4 N^2 DFMA per iteration) instead of the 16 N^2 bytes of shared-memory traffic per iteration a resident matrix costs;
//     pixels that do not reach a 1e-9 residual within 400 iterations join the work list too;
//   * phase reference, compressed SLC and temporal coherence from the same registers.
// Pixels are drawn from a global counter (persistent CTAs; neighbouring CTAs work on neighbouring pixels, so the SHPs'
// samples come from L2).
#include <math_constants.h>

#include "common.cuh"

namespace fringe {
namespace {

#define FULLM 0xffffffffu

// profiling build only (-DFRINGE_PHASE_CLOCKS): cycles of thread 0 per phase into stats[8..15] -- [0] pixel draw + SHP list,
// [1] staging loads, [2] covariance, [3] coherence + |C|, [4] LDL^T test, [5] strip load + power iteration, [6] epilogue
#ifdef FRINGE_PHASE_CLOCKS
#define CPH_DECL long long cph_t = clock64(); unsigned long long cph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CPH_MARK(k) { const long long cph_n = clock64(); cph[k] += (unsigned long long)(cph_n - cph_t); cph_t = cph_n; }
#define CPH_FLUSH if (a.stats && tid == 0) { for (int k = 0; k < 8; ++k) atomicAdd(&a.stats[8 + k], cph[k]); }
#else
#define CPH_DECL
#define CPH_MARK(k)
#define CPH_FLUSH
#endif

struct CtaLayout {
    int ld;        // row stride of C in double2 units; ld % 8 == 4: the 64-byte pieces four lanes read from two rows of a wavefront
                   // fall into different halves of the 128 bytes
    int npad;      // staged sample vector, float2 units (multiple of 4)
    int cap;       // SHPs staged at a time
    int nx;        // length of the broadcast vectors (>= 4 * CP)
    size_t off_a, off_pow, off_dinv, off_xd, off_xo, off_red, off_list, off_misc, bytes;
};

__host__ __device__ inline CtaLayout cta_layout(int N, int W, int CP) {
    CtaLayout L;
    L.ld = ((N + 3) & ~7) + 4;
    if (L.ld < N) L.ld += 8;
    L.npad = (N + 3) & ~3;
    L.nx = 4 * CP;
    size_t o = (size_t)N * L.ld * sizeof(double2);
    L.off_a = o;
    const size_t a_bytes = (size_t)N * (N + 1) / 2 * sizeof(double);
    int cap = (int)(a_bytes / ((size_t)L.npad * sizeof(float2)));
    if (cap < 8) cap = 8;
    if (cap > W) cap = W;
    L.cap = cap;
    const size_t z_bytes = (size_t)cap * L.npad * sizeof(float2);
    o += ((a_bytes > z_bytes ? a_bytes : z_bytes) + 15) & ~(size_t)15;
    L.off_pow = o;  o += (size_t)L.nx * sizeof(double);
    L.off_dinv = o; o += (size_t)L.nx * sizeof(double);
    L.off_xd = o;   o += 2 * (size_t)L.nx * sizeof(double2);
    L.off_xo = o;   o += 2 * (size_t)L.nx * sizeof(float2);
    L.off_red = o;  o += 4 * 16 * sizeof(double2);
    L.off_list = o; o += ((size_t)W * sizeof(int) + 15) & ~(size_t)15;
    L.off_misc = o; o += 16 * sizeof(int);
    L.bytes = o;
    return L;
}

// sum over the CTA of two doubles, the same bits in every thread (one barrier; four rotating buffers, so a buffer is
// written again only three barriers after its last reader)
__device__ __forceinline__ double2 cta_sum2(double a, double b, double2* red, int& slot, int lane, int warp, int nw) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULLM, a, o);
        b += __shfl_xor_sync(FULLM, b, o);
    }
    double2* buf = red + (slot & 3) * 16;
    ++slot;
    if (lane == 0) buf[warp] = make_double2(a, b);
    __syncthreads();
    double2 s = make_double2(0.0, 0.0);
    for (int w = 0; w < nw; ++w) { const double2 v = buf[w]; s.x += v.x; s.y += v.y; }
    return s;
}

template <int CP>
struct CtaCfg { static constexpr int THREADS = 32 * ((16 * CP + 31) / 32); };

template <int CP>
__global__ void __launch_bounds__(CtaCfg<CP>::THREADS) k_evd_cta(const EvdArgs a) {
    constexpr int NT = CtaCfg<CP>::THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int N = a.bands, NP = a.NP;
    const int WX = 2 * a.Nx + 1, W = WX * (2 * a.Ny + 1), center = a.Ny * WX + a.Nx;
    const CtaLayout L = cta_layout(N, W, CP);
    const int ld = L.ld, npad = L.npad;
    double2* Cd = reinterpret_cast<double2*>(smem);
    double* A = reinterpret_cast<double*>(smem + L.off_a);          // |C|, packed lower triangle (row r at r(r+1)/2) ...
    float2* zs = reinterpret_cast<float2*>(smem + L.off_a);         // ... after the staged samples are done with
    double* rp = reinterpret_cast<double*>(smem + L.off_pow);       // band powers, then their inverse square roots
    double* dinv = reinterpret_cast<double*>(smem + L.off_dinv);
    double2* xd = reinterpret_cast<double2*>(smem + L.off_xd);      // [2][nx]
    float2* xo = reinterpret_cast<float2*>(smem + L.off_xo);        // [2][nx]
    double2* red = reinterpret_cast<double2*>(smem + L.off_red);
    int* list = reinterpret_cast<int*>(smem + L.off_list);
    volatile int* misc = reinterpret_cast<volatile int*>(smem + L.off_misc);

    const int k0 = a.mini_stack_count - 1;
    const int need = (a.variant == 0) ? 2 : a.min_neighbors;
    const long npix_block = (long)a.cols * a.lines;
    const long total = (long)a.n_lines * a.cols;
    const int r = tid >> 2, p = tid & 3;            // solver phases: matrix row and column part (columns 4k + p)
    const bool rowp = (r < N) && (p == 0);          // the lane that speaks for row r

    // covariance phase: tile (TI, TJ), TI <= TJ, of 4x4 entries
    const int nt = (N + 3) >> 2, ntiles = nt * (nt + 1) / 2;
    int TI = -1, TJ = -1;
    if (tid < ntiles) {
        int t = tid, I = 0;
        while (t >= nt - I) { t -= nt - I; ++I; }
        TI = I; TJ = I + t;
    }
    for (int i = tid; i < 2 * L.nx; i += NT) xd[i] = make_double2(0.0, 0.0);    // the tails beyond N stay zero
    int slot = 0;
    unsigned long long st_pix = 0, st_it = 0;
    CPH_DECL

    for (;;) {
        __syncthreads();                               // every thread is done with the previous pixel
        if (tid == 0) misc[0] = atomicAdd(&a.worklist[0], 1);
        __syncthreads();
        const long i = misc[0];
        if (i >= total) break;
        const long pix = (long)a.first_line * a.cols + i;
        const int ci = (int)(pix / a.cols), cj = (int)(pix - (long)ci * a.cols);

        // ---- SHP list in raster order (warp 0) -----------------------------------------------
        if (warp == 0) {
            int base = 0;
            for (int f0 = 0; f0 < W; f0 += 32) {
                const int f = f0 + lane;
                bool on = false;
                int q = 0;
                if (f < W) {
                    const uint32_t wd = __ldg(&a.wts[pix * a.nulong + (f >> 5)]);
                    const int fy = f / WX;
                    const int yy = ci + fy - a.Ny, xx = cj + (f - fy * WX) - a.Nx;
                    on = ((wd >> (f & 31)) & 1u) && yy >= 0 && yy < a.lines && xx >= 0 && xx < a.cols;
                    q = yy * a.cols + xx;
                }
                const unsigned b = __ballot_sync(FULLM, on);
                if (on) list[base + __popc(b & ((1u << lane) - 1u))] = q;
                base += __popc(b);
            }
            if (lane == 0) {
                misc[1] = base;
                misc[2] = (int)((__ldg(&a.wts[pix * a.nulong + (center >> 5)]) >> (center & 31)) & 1u);
            }
        }
        __syncthreads();
        CPH_MARK(0)
        const int S = misc[1];
        float tc = 0.f;
        bool solved = false;
        float2 o = make_float2(0.f, 0.f), cmp = make_float2(0.f, 0.f);

        if (misc[2] && S >= need) {
            // ---- covariance (evd.cpp:537-564 / phase_link.cpp:500-527) ------------------------
            double2 acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = make_double2(0.0, 0.0);
            double pw = 0.0;
            for (int c0 = 0; c0 < S; c0 += L.cap) {
                const int ns = min(L.cap, S - c0);
                if (c0 > 0) __syncthreads();
                for (int idx = tid; idx < ns * npad; idx += NT) {
                    const int s = idx / npad, t = idx - s * npad;
                    zs[idx] = (t < N) ? __ldg(&a.zpix[(long)list[c0 + s] * NP + t]) : make_float2(0.f, 0.f);
                }
                __syncthreads();
                CPH_MARK(1)
                if (TI >= 0) {
                    const float2* zi0 = zs + 4 * TI;
                    const float2* zj0 = zs + 4 * TJ;
#pragma unroll 2
                    for (int s = 0; s < ns; ++s) {
                        const float4 i01 = *reinterpret_cast<const float4*>(zi0 + s * npad);
                        const float4 i23 = *reinterpret_cast<const float4*>(zi0 + s * npad + 2);
                        const float4 j01 = *reinterpret_cast<const float4*>(zj0 + s * npad);
                        const float4 j23 = *reinterpret_cast<const float4*>(zj0 + s * npad + 2);
                        const float zix[4] = {i01.x, i01.z, i23.x, i23.z}, ziy[4] = {i01.y, i01.w, i23.y, i23.w};
                        const float zjx[4] = {j01.x, j01.z, j23.x, j23.z}, zjy[4] = {j01.y, j01.w, j23.y, j23.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float pr = __fadd_rn(__fmul_rn(zix[u], zjx[v]), __fmul_rn(ziy[u], zjy[v]));
                                const float pi = __fsub_rn(__fmul_rn(ziy[u], zjx[v]), __fmul_rn(zix[u], zjy[v]));
                                acc[4 * u + v].x += (double)pr;
                                acc[4 * u + v].y += (double)pi;
                            }
                    }
                }
                // |z|^2 as the reference forms it: float hypot, squared and summed in double (:558); lane part p takes
                // every fourth SHP
                if (r < N) {
                    for (int s = ((p - c0) & 3); s < ns; s += 4) {
                        const float2 z = zs[s * npad + r];
                        const float hy = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)z.x, (double)z.x),
                                                                     __dmul_rn((double)z.y, (double)z.y)));
                        pw += (double)hy * (double)hy;
                    }
                }
            }
            pw += __shfl_xor_sync(FULLM, pw, 1);
            pw += __shfl_xor_sync(FULLM, pw, 2);
            ++st_pix;
            // a band that is zero in every SHP: NaNs in C, zheevr reports failure, the reference writes -1
            const int zero_band = __syncthreads_or((r < N) && !(pw > 0.0));      // also: the staged samples are consumed
            CPH_MARK(2)
            if (zero_band) {
                tc = -1.f;
            } else {
                if (rowp) rp[r] = 1.0 / sqrt(pw);
                __syncthreads();
                // ---- coherence (evd.cpp:569-582) and |C| -------------------------------------
                if (TI >= 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int ti = 4 * TI + u, tj = 4 * TJ + v;
                            if (ti < tj && tj < N) {
                                const double sc = rp[ti] * rp[tj];
                                const double2 c = make_double2(acc[4 * u + v].x * sc, acc[4 * u + v].y * sc);
                                Cd[ti * ld + tj] = c;
                                Cd[tj * ld + ti] = make_double2(c.x, -c.y);
                                A[((tj * (tj + 1)) >> 1) + ti] = sqrt(c.x * c.x + c.y * c.y);
                            }
                        }
                }
                if (tid < N) {
                    Cd[tid * ld + tid] = make_double2(1.0, 0.0);
                    A[((tid * (tid + 1)) >> 1) + tid] = 1.0;
                }
                __syncthreads();
                CPH_MARK(3)

                // ---- is |C| positive definite?  LDL^T, column by column (zpotrf's verdict, phase_link.cpp:556-566)
                // A[r][k] holds U = L D (undivided) for the finished columns k, dinv[k] = 1 / D_k
                bool pd = true;
                {
                    const double* Ur = A + ((r * (r + 1)) >> 1);
                    for (int j = 0; j < N; ++j) {
                        const double* Uj = A + ((j * (j + 1)) >> 1);
                        double s = 0.0;
                        const bool act = (r >= j) && (r < N);
                        if (act)
                            for (int k = p; k < j; k += 4) s = fma(Ur[k], Uj[k] * dinv[k], s);
                        s += __shfl_xor_sync(FULLM, s, 1);
                        s += __shfl_xor_sync(FULLM, s, 2);
                        bool bad = false;
                        if (act && p == 0) {
                            const double v = Ur[j] - s;
                            A[((r * (r + 1)) >> 1) + j] = v;
                            if (r == j) { bad = !(v > 0.0); dinv[j] = 1.0 / v; }
                        }
                        if (__syncthreads_or(bad)) { pd = false; break; }
                    }
                }

                CPH_MARK(4)
                bool defer = pd;          // the MLE branch proper: left to the warp-per-pixel kernel
                if (!pd) {
                    // ---- dominant eigenpair of C (phase_link.cpp:586-600): FP64 power iteration with momentum ----
                    double2 c[CP];
#pragma unroll
                    for (int k = 0; k < CP; ++k) {
                        const int j = 4 * k + p;
                        c[k] = (r < N && j < N) ? Cd[r * ld + j] : make_double2(0.0, 0.0);
                    }
                    double2 x = (r < N) ? Cd[r * ld + k0] : make_double2(0.0, 0.0), xp = make_double2(0.0, 0.0);
                    double2 vfin = make_double2(0.0, 0.0);
                    bool got = false;
                    const double nrm = cta_sum2(rowp ? x.x * x.x + x.y * x.y : 0.0, 0.0, red, slot, lane, warp, NW).x;
                    if (nrm > 0.0) {
                        double sc = 1.0 / sqrt(nrm);
                        x.x *= sc; x.y *= sc;
                        int cur = 0;
                        if (rowp) xd[r] = x;
                        __syncthreads();
                        double lam = 1.0, beta = 0.0, rho_prev = -1.0;
                        int next_chk = 2;
                        constexpr int gap = 2;
                        int it = 0;
                        for (; it < 400; ++it) {
                            const double2* xv = xd + cur * L.nx + p;
                            double yr0 = 0.0, yi0 = 0.0, yr1 = 0.0, yi1 = 0.0;
#pragma unroll
                            for (int k = 0; k < CP; ++k) {
                                const double2 xj = xv[4 * k];
                                if (k & 1) {
                                    yr1 = fma(c[k].x, xj.x, yr1); yr1 = fma(-c[k].y, xj.y, yr1);
                                    yi1 = fma(c[k].x, xj.y, yi1); yi1 = fma(c[k].y, xj.x, yi1);
                                } else {
                                    yr0 = fma(c[k].x, xj.x, yr0); yr0 = fma(-c[k].y, xj.y, yr0);
                                    yi0 = fma(c[k].x, xj.y, yi0); yi0 = fma(c[k].y, xj.x, yi0);
                                }
                            }
                            double2 y = make_double2(yr0 + yr1, yi0 + yi1);
                            y.x += __shfl_xor_sync(FULLM, y.x, 1); y.y += __shfl_xor_sync(FULLM, y.y, 1);
                            y.x += __shfl_xor_sync(FULLM, y.x, 2); y.y += __shfl_xor_sync(FULLM, y.y, 2);
                            if (it == next_chk) {
                                const double2 s1 = cta_sum2(rowp ? x.x * y.x + x.y * y.y : 0.0,
                                                            rowp ? x.x * x.x + x.y * x.y : 0.0, red, slot, lane, warp, NW);
                                const double xx = s1.y;
                                lam = s1.x / xx;
                                const double rx = y.x - lam * x.x, ry = y.y - lam * x.y;
                                const double2 s2 = cta_sum2(rowp ? rx * rx + ry * ry : 0.0,
                                                            rowp ? y.x * y.x + y.y * y.y : 0.0, red, slot, lane, warp, NW);
                                const double rho2 = s2.x / (lam * lam * xx);
                                if (rho2 <= 1.0e-18) {                         // one more plain step, then done
                                    sc = 1.0 / sqrt(s2.y);
                                    vfin = make_double2(y.x * sc, y.y * sc);
                                    got = true;
                                    ++it;
                                    break;
                                }
                                if (rho_prev > 0.0 && rho2 < rho_prev) {
                                    if (beta == 0.0) {
                                        const double rr = sqrt(sqrt(rho2 / rho_prev));       // (rho2 / rho_prev)^(0.5 / gap)
                                        beta = fmin(0.575 * rr * 0.575 * rr, 0.2);
                                    }
                                } else if (rho_prev > 0.0) beta *= 0.5;
                                rho_prev = rho2;
                                next_chk = it + gap;
                                sc = 1.0 / sqrt(xx);
                                const double il = 1.0 / lam;
                                const double2 xn = make_double2((y.x * il - beta * xp.x) * sc, (y.y * il - beta * xp.y) * sc);
                                xp = make_double2(x.x * sc, x.y * sc);
                                x = xn;
                            } else if (it < 2) {                               // lambda still unknown: plain normalised steps
                                const double y2 = cta_sum2(rowp ? y.x * y.x + y.y * y.y : 0.0, 0.0, red, slot, lane, warp, NW).x;
                                sc = 1.0 / sqrt(y2);
                                xp = make_double2(0.0, 0.0);
                                x = make_double2(y.x * sc, y.y * sc);
                            } else {
                                const double il = 1.0 / lam;
                                const double2 xn = make_double2(y.x * il - beta * xp.x, y.y * il - beta * xp.y);
                                xp = x;
                                x = xn;
                            }
                            cur ^= 1;
                            if (rowp) xd[cur * L.nx + r] = x;
                            __syncthreads();
                        }
                        st_it += it;
                        CPH_MARK(5)
                        if (got) {
                            // ---- phase reference, compression, temporal coherence (evd.cpp:738-786) ----
                            cur ^= 1;
                            if (rowp) xd[cur * L.nx + r] = vfin;
                            __syncthreads();
                            const double2 ref = xd[cur * L.nx + k0];
                            const double rn = 1.0 / fmax(hypot(ref.x, ref.y), 1e-300);
                            const double qx = ref.x * rn, qy = ref.y * rn;                 // v * conj(ref / |ref|), then FP32
                            const float2 vf = make_float2((float)(vfin.x * qx + vfin.y * qy), (float)(vfin.y * qx - vfin.x * qy));
                            if (rowp) xo[r] = vf;
                            __syncthreads();
                            const float2 rf = xo[k0];
                            float cr = 0.f, cim = 0.f;
                            if (r < N) {
                                float ux = vf.x * rf.x + vf.y * rf.y;
                                float uy = vf.y * rf.x - vf.x * rf.y;
                                const float m = sqrtf(ux * ux + uy * uy);
                                if (m == 0.f) {                                           // arg(0) = 0 in the reference
                                    const float mr = sqrtf(rf.x * rf.x + rf.y * rf.y);
                                    ux = rf.x / mr; uy = -rf.y / mr;
                                } else { ux /= m; uy /= m; }
                                if (r == k0) { ux = 1.f; uy = 0.f; }
                                o = make_float2(ux, uy);
                                if (p == 0) {
                                    xo[L.nx + r] = o;
                                    if (r >= k0) {
                                        const float2 z = __ldg(&a.zpix[pix * NP + r]);
                                        cr = z.x * ux + z.y * uy;                         // z * conj(o)
                                        cim = z.y * ux - z.x * uy;
                                    }
                                }
                            }
                            const double2 sc2 = cta_sum2((double)cr, (double)cim, red, slot, lane, warp, NW);   // barrier: xo[1] complete
                            const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
                            cmp = make_float2((float)sc2.x * invn, (float)sc2.y * invn);
                            float sr = 0.f, si = 0.f;
                            if (r < N) {
                                const float2* ov = xo + L.nx + p;
#pragma unroll
                                for (int k = 0; k < CP; ++k) {
                                    const int j = 4 * k + p;
                                    if (j > r && j < N) {
                                        const float cx = (float)c[k].x, cy = (float)c[k].y;
                                        const float m = sqrtf(cx * cx + cy * cy);
                                        float ex = 1.f, ey = 0.f;
                                        if (m > 0.f) { ex = cx / m; ey = cy / m; }
                                        const float2 oj = ov[4 * k];
                                        const float tx = ex * o.x + ey * o.y, ty = ey * o.x - ex * o.y;   // e * conj(o_r) * o_j
                                        sr += tx * oj.x - ty * oj.y;
                                        si += tx * oj.y + ty * oj.x;
                                    }
                                }
                            }
                            const double2 st = cta_sum2((double)sr, (double)si, red, slot, lane, warp, NW);
                            const float fr = (float)st.x, fi = (float)st.y;
                            tc = sqrtf(fr * fr + fi * fi) / (float)((N * (N - 1)) >> 1);
                            solved = true;
                        }
                    }
                    if (!got) defer = true;       // not converged: certified inverse iteration of the warp-per-pixel kernel
                }
                if (defer) {
                    if (tid == 0) {
                        const int at = atomicAdd(&a.worklist[1], 1);
                        a.worklist[2 + at] = (int)pix;
                    }
                    --st_pix;                     // counted by the kernel that solves it
                    continue;
                }
            }
        }
        if (rowp) a.out[(long)r * npix_block + pix] = solved ? o : make_float2(0.f, 0.f);
        if (tid == 0) { a.tcorr[pix] = tc; a.comp[pix] = cmp; }
        CPH_MARK(6)
    }
    CPH_FLUSH
    if (a.stats && tid == 0) {
        atomicAdd(&a.stats[0], st_pix);
        atomicAdd(&a.stats[1], st_it);
        atomicAdd(&a.stats[2], st_pix);
    }
}

template <int CP>
cudaError_t launch_cta_t(const EvdArgs& a, cudaStream_t st) {
    const int W = (2 * a.Nx + 1) * (2 * a.Ny + 1);
    const CtaLayout L = cta_layout(a.bands, W, CP);
    cudaError_t e = cudaFuncSetAttribute(k_evd_cta<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_evd_cta<CP>, CtaCfg<CP>::THREADS, L.bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const long total = (long)a.n_lines * a.cols;
    long grid = (long)nsm * per_sm;
    if (grid > total) grid = total;
    e = cudaMemsetAsync(a.worklist, 0, 2 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    k_evd_cta<CP><<<(unsigned)grid, CtaCfg<CP>::THREADS, L.bytes, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

// column parts per thread the kernel is instantiated for (bands <= 4 * CP); 0 = not covered
int evd_cta_order(int bands, int Nx, int Ny, int method, int variant) {
    if (variant != 1 || bands <= 32) return 0;           // phase_link only; bands <= 32 has k_mle
    (void)method;
    static const int orders[] = {12, 16, 20, 23, 25, 26};
    for (int cp : orders) {
        if (bands > 4 * cp) continue;
        const int W = (2 * Nx + 1) * (2 * Ny + 1);
        if (cta_layout(bands, W, cp).bytes > 227 * 1024) return 0;
        return cp;
    }
    return 0;
}

cudaError_t launch_evd_cta(const EvdArgs& a, cudaStream_t st) {
    switch (evd_cta_order(a.bands, a.Nx, a.Ny, a.method, a.variant)) {
        case 12: return launch_cta_t<12>(a, st);
        case 16: return launch_cta_t<16>(a, st);
        case 20: return launch_cta_t<20>(a, st);
        case 23: return launch_cta_t<23>(a, st);
        case 25: return launch_cta_t<25>(a, st);
        case 26: return launch_cta_t<26>(a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fringe
