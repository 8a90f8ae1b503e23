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
//   * |C| (packed lower triangle) goes to shared memory; an LDL^T sweep in panels of eight columns (the rows of one warp,
//     the 8x8 diagonal block factored in registers, three barriers per panel) decides positive definiteness.  PD pixels
//     (the MLE branch proper) are appended to a work list and solved by the warp-per-pixel kernel afterwards -- rare
//     where this kernel is chosen;
//   * the fallback's dominant eigenpair, Gram path (padded S <= N / 2 and <= 48): C = B B^H with B = D^-1/2 Z (N x S), so
//     the eigenvector is B u with u the dominant eigenvector of G = B^H B (S x S).  G by tiles on all warps, squared
//     twice (G^4), then a plain power iteration applying G^4 four times per step on S rows x 4 lanes; C itself is never
//     stored, only the unit phasors exp(i arg C_ij) the temporal coherence needs;
//   * otherwise (more SHPs): FP64 power iteration on C with the heavy-ball momentum rule of power_iteration_dp
//     (evd_kernels.cu), four threads per matrix row, most of a thread's row strip in registers; that product is bound by
//     the bytes the shared-memory crossbar delivers for the vector broadcast, which is why the Gram path exists;
//   * pixels that do not reach a 1e-9 residual join the work list too;
//   * phase reference, compressed SLC and temporal coherence behind either solver.
// Pixels are drawn from a global counter (persistent CTAs; neighbouring CTAs work on neighbouring pixels, so the SHPs'
// samples come from L2); the last warp draws the next pixel and fetches its mask words ahead of time.  One CTA per SM at
// 100 dates (209 kB of shared memory), two up to 64 dates.  DESIGN.md 3.6b has the measurements.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"

namespace fringe {
namespace {

#define FULLM 0xffffffffu

// profiling build only (-DFRINGE_PHASE_CLOCKS): cycles of thread 0 per phase into stats[8..15] -- [0] pixel draw + SHP list,
// [1] staging loads, [2] covariance, [3] coherence + |C|, [4] LDL^T test, [5] power iteration (either path), [6] epilogue,
// [7] Gram path: B, G, its squarings and v = B u
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
                   // fall into different halves of the 128 bytes (path that iterates on C itself)
    int npad;      // staged sample vector, float2 units (multiple of 4)
    int cap;       // SHPs staged at a time
    int nx;        // length of the broadcast vectors (>= 4 * CP)
    size_t r0_bytes, a_bytes, off_a, off_pow, off_dd, off_dinv, off_xd, off_xo, off_red, off_list, off_misc, bytes;
};

__host__ __device__ inline CtaLayout cta_layout(int N, int W, int CP) {
    CtaLayout L;
    L.ld = ((N + 3) & ~7) + 4;
    if (L.ld < N) L.ld += 8;
    L.npad = (N + 3) & ~3;
    L.nx = 4 * CP;
    // region 0: C (N x ld double2) -- before that the staged samples, which sit at its END (the Gram path builds its
    // scaled copy of them from the start of the region while they are still needed)
    L.r0_bytes = (size_t)N * L.ld * sizeof(double2);
    int cap = (int)(L.r0_bytes / ((size_t)L.npad * sizeof(float2)));
    if (cap > W) cap = W;
    L.cap = cap;
    size_t o = L.r0_bytes;
    L.off_a = o;
    L.a_bytes = (((size_t)N * (N + 1) / 2 * sizeof(double)) + 15) & ~(size_t)15;
    o += L.a_bytes;
    L.off_pow = o;  o += (size_t)L.nx * sizeof(double);
    L.off_dd = o;   o += (size_t)L.nx * sizeof(double);
    L.off_dinv = o; o += (size_t)L.nx * sizeof(double);
    L.off_xd = o;   o += 2 * (size_t)L.nx * sizeof(double2);
    L.off_xo = o;   o += (size_t)L.nx * sizeof(float2);
    L.off_red = o;  o += 4 * 64 * sizeof(double);
    L.off_list = o; o += ((size_t)W * sizeof(int) + 15) & ~(size_t)15;
    L.off_misc = o; o += 64 * sizeof(int);          // [0..15] scalars, [16..47] the next pixel's mask words
    L.bytes = o;
    return L;
}

// Sum over the CTA of K doubles that only the first lanes of the four-lane groups carry (lanes 0, 4, 8, ...: zero
// elsewhere), the same bits in every thread: three shuffle rounds, one barrier, the per-warp partials added as a tree.
// Four rotating buffers [K][16 warps] (slots of absent warps stay zero), so a buffer is rewritten only three barriers
// after its last reader.
template <int K>
__device__ __forceinline__ void cta_sum_rows(double (&v)[K], double* red, int& slot, int lane, int warp) {
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1)
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(FULLM, v[k], o);
    double* buf = red + (slot & 3) * 64;
    ++slot;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[k * 16 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double2* q = reinterpret_cast<const double2*>(buf + k * 16);
        const double2 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6], q7 = q[7];
        v[k] = (((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y))) +
               (((q4.x + q4.y) + (q5.x + q5.y)) + ((q6.x + q6.y) + (q7.x + q7.y)));
    }
}

// the same among the first nwg warps only (named barrier 1): the Gram-matrix iteration occupies S rows x 4 lanes
template <int K>
__device__ __forceinline__ void grp_sum_rows(double (&v)[K], double* red, int& slot, int lane, int warp, int nwg) {
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1)
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(FULLM, v[k], o);
    double* buf = red + (slot & 3) * 64;
    ++slot;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[k * 16 + warp] = v[k];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nwg * 32) : "memory");
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = 0.0;
        for (int w = 0; w < nwg; ++w) t += buf[k * 16 + w];
        v[k] = t;
    }
}

// 1 / x and 1 / sqrt(x) for positive normal x: the hardware seed (~2^-22) and two Newton steps -- ~50 cycles of latency
// where the IEEE division / square root sequences take several hundred; the solver's scalars need no last-bit rounding
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x * r, r, 1.0);
    r = fma(0.5 * r, e, r);
    e = fma(-x * r, r, 1.0);
    return fma(0.5 * r, e, r);
}

// sum_k g[k] * x[4k] for the first NK strip entries, four independent accumulator chains
template <int NK>
__device__ __forceinline__ double2 strip_dot(const double2 (&g)[12], const double2* xv) {
    double yr[2] = {0.0, 0.0}, yi[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const double2 xj = xv[4 * k];
        yr[k & 1] = fma(g[k].x, xj.x, yr[k & 1]); yr[k & 1] = fma(-g[k].y, xj.y, yr[k & 1]);
        yi[k & 1] = fma(g[k].x, xj.y, yi[k & 1]); yi[k & 1] = fma(g[k].y, xj.x, yi[k & 1]);
    }
    return make_double2(yr[0] + yr[1], yi[0] + yi[1]);
}

template <int CP>
struct CtaCfg {
    static constexpr int THREADS = 32 * ((16 * CP + 31) / 32);
    // strip entries held in registers; the rest is read from the shared-memory copy every iteration.  Thirteen warps leave
    // 128 registers per thread (four warps on one scheduler): 100+ strip registers would spill to local memory, i.e. to L2
    static constexpr int KR = (THREADS > 384) ? CP - 8 : CP;
    // up to 64 dates two CTAs fit one SM (shared memory and, at 128 registers, the register file): twice the warps to
    // cover the latency chains of every phase
    static constexpr int MIN_CTAS = (THREADS <= 256) ? 2 : 1;
};
#ifndef CTA_NSQ
#define CTA_NSQ 2            // squarings of the Gram matrix before the power iteration
#endif
constexpr int KG = 12;          // Gram path: strip entries per thread, i.e. at most 48 SHPs

__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

template <int CP>
__global__ void __launch_bounds__(CtaCfg<CP>::THREADS, CtaCfg<CP>::MIN_CTAS) k_evd_cta(const EvdArgs a) {
    constexpr int NT = CtaCfg<CP>::THREADS, KR = CtaCfg<CP>::KR;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.bands, NP = a.NP;
    const int WX = 2 * a.Nx + 1, W = WX * (2 * a.Ny + 1), center = a.Ny * WX + a.Nx;
    const CtaLayout L = cta_layout(N, W, CP);
    const int ld = L.ld, npad = L.npad;
    double2* Cd = reinterpret_cast<double2*>(smem);                 // C: upper triangle (and the lower one in the tail columns >= 4 KR)
    double* A = reinterpret_cast<double*>(smem + L.off_a);          // |C|, packed lower triangle (row r at r(r+1)/2); later the Gram matrix
    double* rpw = reinterpret_cast<double*>(smem + L.off_pow);      // inverse square roots of the band powers
    double* rp = reinterpret_cast<double*>(smem + L.off_dd);        // the pivots D of the LDL^T sweep
    double* dinv = reinterpret_cast<double*>(smem + L.off_dinv);
    double2* xd = reinterpret_cast<double2*>(smem + L.off_xd);      // [2][nx]
    float2* xo = reinterpret_cast<float2*>(smem + L.off_xo);        // [nx]
    double* red = reinterpret_cast<double*>(smem + L.off_red);
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
    for (int i = tid; i < 4 * 64; i += NT) red[i] = 0.0;
    int slot = 0;
    unsigned long long st_pix = 0, st_it = 0;
    CPH_DECL

    // The next pixel is drawn, and its mask words fetched, by the last warp while the others work on the current one
    // (misc[8] = its index, misc[9] = words present in wpre): three dependent global round trips off the critical path
    volatile uint32_t* wpre = reinterpret_cast<volatile uint32_t*>(misc + 16);
    constexpr int NW = NT / 32;
    if (tid == 0) { misc[8] = atomicAdd(&a.worklist[0], 1); misc[9] = 0; }
    for (;;) {
        __syncthreads();                               // every thread is done with the previous pixel
        const long i = misc[8];
        const bool pre = misc[9] != 0;
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
                    const uint32_t wd = pre ? wpre[f >> 5] : __ldg(&a.wts[pix * a.nulong + (f >> 5)]);
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
                const uint32_t cw = pre ? wpre[center >> 5] : __ldg(&a.wts[pix * a.nulong + (center >> 5)]);
                misc[2] = (int)((cw >> (center & 31)) & 1u);
            }
        }
        __syncthreads();
        CPH_MARK(0)
        int nxt_pending = 0;
        if (warp == NW - 1 && lane == 0) nxt_pending = atomicAdd(&a.worklist[0], 1);      // consumed in publish_next
        auto publish_next = [&]() {
            if (warp == NW - 1) {
                const int nxt = __shfl_sync(FULLM, nxt_pending, 0);
                const bool have = (nxt < total) && (a.nulong <= 32);
                if (have && lane < a.nulong) wpre[lane] = __ldg(&a.wts[((long)a.first_line * a.cols + nxt) * a.nulong + lane]);
                if (lane == 0) { misc[8] = nxt; misc[9] = have ? 1 : 0; }
            }
        };
        const int S = misc[1];
        float tc = 0.f;
        bool solved = false;
        float2 o = make_float2(0.f, 0.f), cmp = make_float2(0.f, 0.f);

        // Gram path: with fewer SHPs than dates, C = B B^H (B = D^-1/2 Z, N x S) has rank S, and its dominant eigenvector is
        // B u with u the dominant eigenvector of the S x S Gram matrix G = B^H B -- (N/S)^2 fewer operations per iteration.
        // Taken when the padded S is at most N / 2 (G fits the |C| buffer; B, the unit phasors of C and the staged samples
        // fit region 0) and the SHPs were staged in one piece.
        // staged samples [min(S, cap)][npad], ending with region 0
        const size_t zs_bytes = (size_t)min(S, L.cap) * npad * sizeof(float2);
        float2* zs = reinterpret_cast<float2*>(smem + L.r0_bytes - zs_bytes);
        const int SB = (S + 3) & ~3, SBP = SB + 1;
        const int ntp = (ntiles + 7) & ~7;
        const size_t bd_bytes = (size_t)N * SBP * sizeof(double2), e_bytes = (size_t)16 * ntp * sizeof(float2);
        const size_t h_bytes = (size_t)SB * SBP * sizeof(double2);             // second S x S buffer of the squarings, over the samples
        const size_t tail_bytes = zs_bytes > h_bytes ? zs_bytes : h_bytes;
        const bool gpath = (SB <= 4 * KG) && (S <= L.cap) && (h_bytes <= L.a_bytes) &&
                           (bd_bytes + e_bytes + tail_bytes <= L.r0_bytes);
        double2* Bd = reinterpret_cast<double2*>(smem);                      // [N][SBP]
        float2* E = reinterpret_cast<float2*>(smem + bd_bytes);              // [16][ntp]: unit phasors of this thread's tile

        if (misc[2] && S >= need) {
            // the centre pixel's own sample of this row, for the compressed SLC (in flight during everything below)
            const float2 zc = (rowp && r >= k0) ? __ldg(&a.zpix[pix * NP + r]) : make_float2(0.f, 0.f);
            // ---- covariance (evd.cpp:537-564 / phase_link.cpp:500-527) ------------------------
            bool pd = true;
            int zero_band = 0;
            ++st_pix;
          do {
            double2 acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = make_double2(0.0, 0.0);
            double pw = 0.0;
            for (int c0 = 0; c0 < S; c0 += L.cap) {
                const int ns = min(L.cap, S - c0);
                if (c0 > 0) __syncthreads();
                // eight loads in flight per thread before the first store (one round trip per batch, not per element)
                for (int b0 = 0; b0 < ns * npad; b0 += 8 * NT) {
                    float2 zv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = b0 + u * NT + tid;
                        zv[u] = make_float2(0.f, 0.f);
                        if (idx < ns * npad) {
                            const int s = idx / npad, t = idx - s * npad;
                            if (t < N) zv[u] = __ldg(&a.zpix[(long)list[c0 + s] * NP + t]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = b0 + u * NT + tid;
                        if (idx < ns * npad) zs[idx] = zv[u];
                    }
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
            // a band that is zero in every SHP: NaNs in C, zheevr reports failure, the reference writes -1
            zero_band = __syncthreads_or((r < N) && !(pw > 0.0));      // also: the staged samples are consumed
            publish_next();
            CPH_MARK(2)
            if (zero_band) break;
            {
                if (rowp) rpw[r] = fast_rsqrt(pw);
                __syncthreads();
                // ---- coherence (evd.cpp:569-582) and |C| -------------------------------------
                // C: the upper triangle only (the solver's strips read the mirror image), except for the tail columns
                // the strips fetch from here in every iteration
                if (TI >= 0) {
                    const bool mirror = (TI >= KR);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int ti = 4 * TI + u, tj = 4 * TJ + v;
                            if (ti < tj && tj < N) {
                                const double sc = rpw[ti] * rpw[tj];
                                const double2 c = make_double2(acc[4 * u + v].x * sc, acc[4 * u + v].y * sc);
                                const double m2 = c.x * c.x + c.y * c.y;
                                A[tri(tj) + ti] = (m2 > 0.0) ? m2 * fast_rsqrt(m2) : 0.0;          // |C_ij| to a few ulp
                                if (gpath) {                       // C itself is only needed as exp(i arg C) (temporal coherence)
                                    const float cx = (float)c.x, cy = (float)c.y;
                                    const float f2 = cx * cx + cy * cy;
                                    float2 e = make_float2(1.f, 0.f);
                                    if (f2 > 0.f) { const float im = rsqrtf(f2); e = make_float2(cx * im, cy * im); }
                                    E[(4 * u + v) * ntp + tid] = e;
                                } else {
                                    Cd[ti * ld + tj] = c;
                                    if (mirror) Cd[tj * ld + ti] = make_double2(c.x, -c.y);
                                }
                            }
                        }
                }
                if (tid < N) {
                    if (!gpath) Cd[tid * ld + tid] = make_double2(1.0, 0.0);
                    A[tri(tid) + tid] = 1.0;
                }
                __syncthreads();
                CPH_MARK(3)

                // ---- is |C| positive definite?  (zpotrf's verdict, phase_link.cpp:556-566) --------------------
                // LDL^T in panels of eight columns = the rows of one warp, three barriers per panel: (A) every row
                // takes the finished columns' contribution off its panel entries, (B) the panel's warp factors the 8x8
                // diagonal block, (C) the rows below finish their panel columns in registers.  A[r][k] = L[r][k] for
                // finished columns k, rp[k] = D_k, dinv[k] = 1 / D_k.
                pd = true;
                {
                    double* Ar = A + tri(r);
                    for (int j0 = 0; j0 < N; j0 += 8) {
                        const int jw = min(8, N - j0);
                        if (j0 > 0 && r >= j0 && r < N) {
                            const int ca = j0 + p, cb = j0 + p + 4;              // this lane's two panel columns
                            const bool ha = (ca <= r) && (p < jw), hb = (cb <= r) && (p + 4 < jw);
                            const double* La = A + tri(min(ca, N - 1));
                            const double* Lb = A + tri(min(cb, N - 1));
                            double sa = 0.0, sb = 0.0;
#pragma unroll 4
                            for (int k = 0; k < j0; ++k) {
                                const double t = Ar[k] * rp[k];
                                sa = fma(t, La[k], sa);
                                sb = fma(t, Lb[k], sb);
                            }
                            if (ha) Ar[ca] -= sa;
                            if (hb) Ar[cb] -= sb;
                        }
                        __syncthreads();
                        bool bad = false;
                        if (warp == (j0 >> 3)) {
                            // the 8x8 diagonal block, factored redundantly by every lane in registers (no exchange, no
                            // barrier inside); rows / columns beyond N behave like an identity block
                            double v[8][8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int j = 0; j <= i; ++j)
                                    v[i][j] = (i < jw) ? A[tri(j0 + i) + j0 + j] : (i == j ? 1.0 : 0.0);
                            double dv[8], di[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                dv[c] = v[c][c];
                                bad = bad || !(dv[c] > 0.0);
                                di[c] = fast_rcp(dv[c]);
#pragma unroll
                                for (int j = c + 1; j < 8; ++j) {
                                    const double l = v[j][c] * di[c];
#pragma unroll
                                    for (int i = j; i < 8; ++i) v[i][j] = fma(-v[i][c], l, v[i][j]);     // V[i][j] -= V[i][c] L[j][c]
                                }
#pragma unroll
                                for (int i = c + 1; i < 8; ++i) v[i][c] *= di[c];                          // column c: V -> L
                            }
                            // every lane holds the same values and stores them all (same address, same bits: one wavefront per
                            // store); a lane-dependent choice of the entry would turn v[][] into a local-memory array
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int j = 0; j < i; ++j)
                                    if (i < jw) A[tri(j0 + i) + j0 + j] = v[i][j];
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                if (c < jw) { rp[j0 + c] = dv[c]; dinv[j0 + c] = di[c]; }
                        }
                        if (__syncthreads_or(bad)) { pd = false; break; }
                        if (r >= j0 + 8 && r < N && p == 0) {                            // then jw == 8
                            double u[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) u[c] = Ar[j0 + c];
#pragma unroll
                            for (int c = 0; c < 8; ++c)
#pragma unroll
                                for (int b = c + 1; b < 8; ++b) u[b] = fma(-u[c], A[tri(j0 + b) + j0 + c], u[b]);
#pragma unroll
                            for (int c = 0; c < 8; ++c) Ar[j0 + c] = u[c] * dinv[j0 + c];
                        }
                        __syncthreads();
                    }
                }
                CPH_MARK(4)
            }
          } while (false);
            if (zero_band) {
                tc = -1.f;
            } else {
                bool defer = pd;          // the MLE branch proper: left to the warp-per-pixel kernel
                if (!pd) {
                    // ---- dominant eigenpair of C (phase_link.cpp:586-600) ---------------------------------------
                    // phase reference, compression, temporal coherence (evd.cpp:738-786) behind either solver; tsum adds
                    // this thread's share of sum_{i<j} exp(i arg C_ij) conj(o_i) o_j, with the phasors o in xo
                    auto finish = [&](const double2 vfin, auto&& tsum) {
                        if (rowp) xd[r] = vfin;
                        __syncthreads();
                        const double2 ref = xd[k0];
                        const double rn = 1.0 / fmax(sqrt(ref.x * ref.x + ref.y * ref.y), 1e-300);
                        const double qx = ref.x * rn, qy = ref.y * rn;                 // v * conj(ref / |ref|), then FP32
                        const float2 vf = make_float2((float)(vfin.x * qx + vfin.y * qy), (float)(vfin.y * qx - vfin.x * qy));
                        const float2 rf = make_float2((float)(ref.x * qx + ref.y * qy), (float)(ref.y * qx - ref.x * qy));
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
                            if (p == 0) xo[r] = o;
                        }
                        __syncthreads();
                        float sr = 0.f, si = 0.f;
                        tsum(sr, si);
                        sr += __shfl_xor_sync(FULLM, sr, 1); si += __shfl_xor_sync(FULLM, si, 1);
                        sr += __shfl_xor_sync(FULLM, sr, 2); si += __shfl_xor_sync(FULLM, si, 2);
                        double s4[4];
                        s4[0] = rowp ? (double)(zc.x * o.x + zc.y * o.y) : 0.0;        // z * conj(o), rows >= k0 (zc = 0 elsewhere)
                        s4[1] = rowp ? (double)(zc.y * o.x - zc.x * o.y) : 0.0;
                        s4[2] = rowp ? (double)sr : 0.0;
                        s4[3] = rowp ? (double)si : 0.0;
                        cta_sum_rows<4>(s4, red, slot, lane, warp);
                        const float invn = 1.0f / (float)(N - a.mini_stack_count + 1);
                        cmp = make_float2((float)s4[0] * invn, (float)s4[1] * invn);
                        const float fr = (float)s4[2], fi = (float)s4[3];
                        tc = sqrtf(fr * fr + fi * fi) / (float)((N * (N - 1)) >> 1);
                        solved = true;
                    };
                    bool got = false;
                    int it = 0;
                    if (gpath) {
                        // ---- (a) B = D^-1/2 Z in double, [date][SHP], rows SBP (odd) apart; SHP columns beyond S are zero ----
                        for (int idx = tid; idx < SB * npad; idx += NT) {
                            const int sh = idx / npad, t = idx - sh * npad;
                            if (t < N) {
                                double2 bv = make_double2(0.0, 0.0);
                                if (sh < S) { const float2 z = zs[idx]; const double w = rpw[t]; bv = make_double2((double)z.x * w, (double)z.y * w); }
                                Bd[t * SBP + sh] = bv;
                            }
                        }
                        __syncthreads();
                        // ---- (b) G = B^H B, then G^4 by two squarings (G is Hermitian: G^2 = G^H G, the same product).
                        // dst = src^H src for src [nrows][stride]: 4x4 tiles of the upper triangle, eight lanes per tile each
                        // taking every eighth row; the eight partial tiles are folded by recursive halving (lane q ends with
                        // entries 2q, 2q+1).  Every thread of the CTA works here -- throughput work -- where a power iteration
                        // on G itself is a chain of ~22 latency-bound steps on five warps: with G^4 applied four times per step it takes ~5.
                        // trace G = N, so a unit vector times G^16 stays below 1e32.
                        // LP lanes per tile: 8 for the N rows of B, 4 for the squarings (S rows: the folding shuffles, not
                        // the products, are what a pass costs there)
                        auto gram = [&](auto lp_tag, const double2* src, const int nrows, const int stride, double2* G) {
                            constexpr int LP = decltype(lp_tag)::value;
                            const int nbg = SB >> 2, ntg = nbg * (nbg + 1) / 2, q = lane & (LP - 1);
                            for (int t0 = 0; t0 < ntg; t0 += NT / LP) {
                                const int tile = t0 + tid / LP;
                                const bool act = tile < ntg;
                                int GI = 0, GJ = 0;
                                if (act) { int t = tile; while (t >= nbg - GI) { t -= nbg - GI; ++GI; } GJ = GI + t; }
                                double2 g[16];
#pragma unroll
                                for (int e = 0; e < 16; ++e) g[e] = make_double2(0.0, 0.0);
                                if (act) {
                                    for (int t = q; t < nrows; t += LP) {
                                        const double2* bi = src + t * stride + 4 * GI;
                                        const double2* bj = src + t * stride + 4 * GJ;
                                        double2 vi[4], vj[4];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) { vi[u] = bi[u]; vj[u] = bj[u]; }
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
#pragma unroll
                                            for (int v = 0; v < 4; ++v) {                        // conj(b_i) b_j
                                                g[4 * u + v].x = fma(vi[u].x, vj[v].x, g[4 * u + v].x);
                                                g[4 * u + v].x = fma(vi[u].y, vj[v].y, g[4 * u + v].x);
                                                g[4 * u + v].y = fma(vi[u].x, vj[v].y, g[4 * u + v].y);
                                                g[4 * u + v].y = fma(-vi[u].y, vj[v].x, g[4 * u + v].y);
                                            }
                                    }
                                }
                                // lane q of the tile's LP lanes ends with entries (16 / LP) q ... of the 4x4 tile
                                constexpr int X1 = LP / 2, X2 = LP / 4;                  // partner distances of the first two rounds
                                const bool h1 = (q & X1) != 0, h2 = (q & X2) != 0, h3 = (q & 1) != 0;
                                double2 g8[8], g4[4], g2[2];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const double2 lo = g[e], hi = g[e + 8];
                                    const double2 snd = h1 ? lo : hi, kp = h1 ? hi : lo;
                                    g8[e] = make_double2(kp.x + __shfl_xor_sync(FULLM, snd.x, X1), kp.y + __shfl_xor_sync(FULLM, snd.y, X1));
                                }
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const double2 lo = g8[e], hi = g8[e + 4];
                                    const double2 snd = h2 ? lo : hi, kp = h2 ? hi : lo;
                                    g4[e] = make_double2(kp.x + __shfl_xor_sync(FULLM, snd.x, X2), kp.y + __shfl_xor_sync(FULLM, snd.y, X2));
                                }
                                if constexpr (LP == 8) {
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        const double2 lo = g4[e], hi = g4[e + 2];
                                        const double2 snd = h3 ? lo : hi, kp = h3 ? hi : lo;
                                        g2[e] = make_double2(kp.x + __shfl_xor_sync(FULLM, snd.x, 1), kp.y + __shfl_xor_sync(FULLM, snd.y, 1));
                                    }
                                }
                                if (act) {
                                    constexpr int PER = 16 / LP;
#pragma unroll
                                    for (int e = 0; e < PER; ++e) {
                                        const int id = PER * q + e, row = 4 * GI + (id >> 2), col = 4 * GJ + (id & 3);
                                        const double2 val = (LP == 8) ? g2[e & 1] : g4[e];
                                        G[row * SBP + col] = val;
                                        if (GI != GJ) G[col * SBP + row] = make_double2(val.x, -val.y);
                                    }
                                }
                            }
                        };
                        double2* G = reinterpret_cast<double2*>(A);                                  // [SB][SBP] (odd row stride, as B: conflict-free column blocks), the |C| buffer
                        double2* H = reinterpret_cast<double2*>(smem + L.r0_bytes - tail_bytes);       // the same, over the staged samples
                        gram(std::integral_constant<int, 8>{}, Bd, N, SBP, G);
                        __syncthreads();
#if CTA_NSQ >= 1
                        gram(std::integral_constant<int, 4>{}, G, SB, SBP, H);          // G^2
                        __syncthreads();
#endif
#if CTA_NSQ >= 2
                        gram(std::integral_constant<int, 4>{}, H, SB, SBP, G);          // G^4
                        __syncthreads();
#endif
                        CPH_MARK(7)
                        // ---- (c) dominant eigenvector u of G^16 = (G^4)^4: plain power iteration on S rows x 4 lanes, the row strips
                        // (<= 12 entries) in registers, the first nwg warps only (named barrier), Rayleigh quotient and
                        // residual in every step.  A residual <= 1e-9 on G^16 bounds the one on G (same eigenvectors,
                        // 1 - (l_i / l_1)^16 >= 1 - l_i / l_1).
                        const int nwg = (4 * SB + 31) >> 5;
                        if (warp < nwg) {
                            const bool rowg = (r < SB) && (p == 0);
                            double2 gk[KG];
#pragma unroll
                            for (int k = 0; k < KG; ++k)
                                gk[k] = (r < SB && 4 * k + p < SB) ? ((CTA_NSQ == 1) ? H : G)[r * SBP + 4 * k + p] : make_double2(0.0, 0.0);
                            double2 x = make_double2(0.0, 0.0);
                            if (r < SB) { const double2 b0 = Bd[k0 * SBP + r]; x = make_double2(b0.x, -b0.y); }    // u0 = B^H e_k0
                            double s4[4];
                            s4[0] = rowg ? x.x * x.x + x.y * x.y : 0.0;
                            grp_sum_rows<1>(reinterpret_cast<double(&)[1]>(s4[0]), red, slot, lane, warp, nwg);
                            if (s4[0] > 0.0) {
                                const double sc0 = fast_rsqrt(s4[0]);
                                x.x *= sc0; x.y *= sc0;
                                int cur = 0;
                                if (p == 0 && r >= SB && r < 4 * KG) { xd[r] = make_double2(0.0, 0.0); xd[L.nx + r] = make_double2(0.0, 0.0); }
                                if (rowg) xd[r] = x;
                                asm volatile("bar.sync 1, %0;" ::"r"(nwg * 32) : "memory");
                                double lam = 1.0;
                                for (; it < 150; ++it) {
                                    // G^4 applied four times per step (a product is ~10x cheaper than another squaring):
                                    // three plain applications (unit vector times at most N^12) ...
#pragma unroll 1
                                    for (int rep = 0; rep < 3; ++rep) {
                                        const double2* xw = xd + cur * L.nx + p;
                                        double2 yw = (SB <= 32) ? strip_dot<8>(gk, xw) : strip_dot<KG>(gk, xw);
                                        yw.x += __shfl_xor_sync(FULLM, yw.x, 1); yw.y += __shfl_xor_sync(FULLM, yw.y, 1);
                                        yw.x += __shfl_xor_sync(FULLM, yw.x, 2); yw.y += __shfl_xor_sync(FULLM, yw.y, 2);
                                        x = yw;
                                        cur ^= 1;
                                        if (rowg) xd[cur * L.nx + r] = x;
                                        asm volatile("bar.sync 1, %0;" ::"r"(nwg * 32) : "memory");
                                    }
                                    // ... and the fourth with the Rayleigh quotient and the residual
                                    const double2* xv = xd + cur * L.nx + p;
                                    // straight-line code (a branch per strip entry would serialise the loads): 8 entries, or
                                    // all 12 -- the strip and the vector are zero beyond SB
                                    double2 y = (SB <= 32) ? strip_dot<8>(gk, xv) : strip_dot<KG>(gk, xv);
                                    y.x += __shfl_xor_sync(FULLM, y.x, 1); y.y += __shfl_xor_sync(FULLM, y.y, 1);
                                    y.x += __shfl_xor_sync(FULLM, y.x, 2); y.y += __shfl_xor_sync(FULLM, y.y, 2);
                                    // x.y, x.x, y.y and the residual against the previous Rayleigh quotient:
                                    // |y - lam x|^2 = |y - lam' x|^2 - (lam - lam')^2 x.x
                                    const double lam_prev = lam;
                                    const double rx = y.x - lam_prev * x.x, ry = y.y - lam_prev * x.y;
                                    s4[0] = rowg ? x.x * y.x + x.y * y.y : 0.0;
                                    s4[1] = rowg ? x.x * x.x + x.y * x.y : 0.0;
                                    s4[2] = rowg ? y.x * y.x + y.y * y.y : 0.0;
                                    s4[3] = rowg ? rx * rx + ry * ry : 0.0;
                                    grp_sum_rows<4>(s4, red, slot, lane, warp, nwg);
                                    const double xx = s4[1];
                                    lam = s4[0] * fast_rcp(xx);
                                    const double dl = lam - lam_prev;
                                    const double r2 = fmax(s4[3] - dl * dl * xx, 0.0);
                                    const double sc = fast_rsqrt(s4[2]);
                                    x = make_double2(y.x * sc, y.y * sc);
                                    cur ^= 1;
                                    if (rowg) xd[cur * L.nx + r] = x;
                                    if (r2 <= 1.0e-18 * lam * lam * xx && fabs(dl) <= 1.0e-3 * lam) { got = true; ++it; break; }
                                    asm volatile("bar.sync 1, %0;" ::"r"(nwg * 32) : "memory");
                                }
                                if (lane == 0 && warp == 0) { misc[3] = got ? 1 : 0; misc[4] = cur; misc[5] = it; misc[6] = slot; }
                            } else if (lane == 0 && warp == 0) { misc[3] = 0; misc[4] = 0; misc[5] = 0; misc[6] = slot; }
                        }
                        __syncthreads();
                        got = misc[3] != 0;
                        it = misc[5];
                        slot = misc[6];                   // the reduction buffers rotate in step in every warp
                        st_it += it;
                        CPH_MARK(5)
                        if (got) {
                            // ---- (d) v = B u, normalised ----
                            const double2* uv = xd + misc[4] * L.nx + p;
                            double2 v = make_double2(0.0, 0.0);
                            if (r < N) {
                                const double2* br = Bd + r * SBP + p;
                                for (int k = 0; 4 * k < SB; ++k) {
                                    const double2 b = br[4 * k], u = uv[4 * k];
                                    v.x = fma(b.x, u.x, v.x); v.x = fma(-b.y, u.y, v.x);
                                    v.y = fma(b.x, u.y, v.y); v.y = fma(b.y, u.x, v.y);
                                }
                            }
                            v.x += __shfl_xor_sync(FULLM, v.x, 1); v.y += __shfl_xor_sync(FULLM, v.y, 1);
                            v.x += __shfl_xor_sync(FULLM, v.x, 2); v.y += __shfl_xor_sync(FULLM, v.y, 2);
                            double n1[1] = {rowp ? v.x * v.x + v.y * v.y : 0.0};
                            cta_sum_rows<1>(n1, red, slot, lane, warp);        // barrier: every thread is done reading u
                            const double sc = 1.0 / sqrt(n1[0]);
                            CPH_MARK(7)
                            finish(make_double2(v.x * sc, v.y * sc), [&](float& sr, float& si) {
                                if (TI >= 0) {
#pragma unroll
                                    for (int u = 0; u < 4; ++u)
#pragma unroll
                                        for (int w = 0; w < 4; ++w) {
                                            const int ti = 4 * TI + u, tj = 4 * TJ + w;
                                            if (ti < tj && tj < N) {
                                                const float2 e = E[(4 * u + w) * ntp + tid];
                                                const float2 oi = xo[ti], oj = xo[tj];
                                                const float tx = e.x * oi.x + e.y * oi.y, ty = e.y * oi.x - e.x * oi.y;
                                                sr += tx * oj.x - ty * oj.y;
                                                si += tx * oj.y + ty * oj.x;
                                            }
                                        }
                                }
                            });
                        }
                    } else {
                        // ---- power iteration on C itself (more SHPs than half the dates) ----
                        auto centry = [&](int j) -> double2 {                    // C[r][j] from the stored triangle
                            if (j >= r) return Cd[r * ld + j];
                            const double2 m = Cd[j * ld + r];
                            return make_double2(m.x, -m.y);
                        };
                        double2 c[KR];
#pragma unroll
                        for (int k = 0; k < KR; ++k) {
                            const int j = 4 * k + p;
                            c[k] = (r < N && j < N) ? centry(j) : make_double2(0.0, 0.0);
                        }
                        // tail columns: straight from shared memory (both triangles are stored there); rows beyond N read row 0
                        // and are never used
                        const double2* ctail = Cd + (r < N ? r : 0) * ld + p;
                        double2 x = (r < N) ? centry(k0) : make_double2(0.0, 0.0), xp = make_double2(0.0, 0.0);
                        double2 vfin = make_double2(0.0, 0.0);
                        double s4[4];
                        s4[0] = rowp ? x.x * x.x + x.y * x.y : 0.0;
                        cta_sum_rows<1>(reinterpret_cast<double(&)[1]>(s4[0]), red, slot, lane, warp);
                        const double nrm = s4[0];
                        if (nrm > 0.0) {
                            double sc = 1.0 / sqrt(nrm);
                            x.x *= sc; x.y *= sc;
                            int cur = 0;
                            if (rowp) xd[r] = x;
                            __syncthreads();
                            double lam = 1.0, beta = 0.0, rho_prev = -1.0;
                            int next_chk = 2;
                            constexpr int gap = 2;
                            for (; it < 400; ++it) {
                                const double2* xv = xd + cur * L.nx + p;
                                double yr[4] = {0.0, 0.0, 0.0, 0.0}, yi[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                                for (int k = 0; k < CP; ++k) {
                                    const double2 xj = xv[4 * k];
                                    double2 ck;
                                    if (k < KR) ck = c[k < KR ? k : 0];
                                    else ck = (4 * k + p < N) ? ctail[4 * k] : make_double2(0.0, 0.0);
                                    yr[k & 3] = fma(ck.x, xj.x, yr[k & 3]); yr[k & 3] = fma(-ck.y, xj.y, yr[k & 3]);
                                    yi[k & 3] = fma(ck.x, xj.y, yi[k & 3]); yi[k & 3] = fma(ck.y, xj.x, yi[k & 3]);
                                }
                                double2 y = make_double2((yr[0] + yr[1]) + (yr[2] + yr[3]), (yi[0] + yi[1]) + (yi[2] + yi[3]));
                                y.x += __shfl_xor_sync(FULLM, y.x, 1); y.y += __shfl_xor_sync(FULLM, y.y, 1);
                                y.x += __shfl_xor_sync(FULLM, y.x, 2); y.y += __shfl_xor_sync(FULLM, y.y, 2);
                                if (it == next_chk) {
                                    // one reduction per test: x.y, x.x, y.y and the residual against the previous Rayleigh
                                    // quotient; |y - lam x|^2 = |y - lam' x|^2 - (lam - lam')^2 x.x  (the residual of the
                                    // Rayleigh quotient is orthogonal to x), free of cancellation once lam has settled
                                    const double lam_prev = lam;
                                    const double rx = y.x - lam_prev * x.x, ry = y.y - lam_prev * x.y;
                                    s4[0] = rowp ? x.x * y.x + x.y * y.y : 0.0;
                                    s4[1] = rowp ? x.x * x.x + x.y * x.y : 0.0;
                                    s4[2] = rowp ? y.x * y.x + y.y * y.y : 0.0;
                                    s4[3] = rowp ? rx * rx + ry * ry : 0.0;
                                    cta_sum_rows<4>(s4, red, slot, lane, warp);
                                    const double xx = s4[1];
                                    lam = s4[0] / xx;
                                    const double dl = lam - lam_prev;
                                    const double r2 = fmax(s4[3] - dl * dl * xx, 0.0);
                                    const double rho2 = r2 / (lam * lam * xx);
                                    if (rho2 <= 1.0e-18 && fabs(dl) <= 1.0e-3 * lam) {     // one more plain step, then done
                                        sc = 1.0 / sqrt(s4[2]);
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
                                    s4[0] = rowp ? y.x * y.x + y.y * y.y : 0.0;
                                    cta_sum_rows<1>(reinterpret_cast<double(&)[1]>(s4[0]), red, slot, lane, warp);
                                    sc = 1.0 / sqrt(s4[0]);
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
                                finish(vfin, [&](float& sr, float& si) {
                                    if (r < N) {
                                        const float2* ov = xo + p;
#pragma unroll
                                        for (int k = 0; k < CP; ++k) {
                                            const int j = 4 * k + p;
                                            if (j > r && j < N) {
                                                double2 ck;
                                                if (k < KR) ck = c[k < KR ? k : 0];
                                                else ck = ctail[4 * k];
                                                const float cx = (float)ck.x, cy = (float)ck.y;
                                                const float m2 = cx * cx + cy * cy;
                                                float ex = 1.f, ey = 0.f;
                                                if (m2 > 0.f) { const float im = rsqrtf(m2); ex = cx * im; ey = cy * im; }
                                                const float2 oj = ov[4 * k];
                                                const float tx = ex * o.x + ey * o.y, ty = ey * o.x - ex * o.y;   // e * conj(o_r) * o_j
                                                sr += tx * oj.x - ty * oj.y;
                                                si += tx * oj.y + ty * oj.x;
                                            }
                                        }
                                    }
                                });
                            }
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
        } else {
            publish_next();
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
