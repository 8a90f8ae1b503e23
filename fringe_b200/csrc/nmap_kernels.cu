// SHP selection kernels (sm_100a).
//
// What the reference does (src/nmap/nmap.cpp:370-473): amplitude of every date, per-pixel
// ascending sort, then for every pair of pixels inside a (2Ny+1)x(2Nx+1) window a two-sample
// KS or AD test on the two sorted vectors; pairs with p-value >= threshold set a bit in each
// other's window bitmask and bump each other's count.
//
// How it is done here:
//   k_amp_sort   one thread per pixel; the N amplitudes of a pixel live in a shared-memory
//                column (bank = thread), are insertion-sorted there and written rank-major
//                ([rank][pixel]) so later tile loads are contiguous row segments.
//   k_nmap<M,PS> one CTA per tile of pixels; the sorted vectors of tile + forward halo are staged in
//                shared memory by TMA: one cp.async.bulk.tensor box (tile + halo, one rank) per rank
//                from the 3-D tensor [rank][line][column] into planes PS words apart, all arriving on
//                one mbarrier; out-of-image parts of a box are zero-filled by the copy engine, and a
//                zero top rank *is* the validity flag (k_amp_sort writes zeros for invalid pixels).
//                One thread per pixel then walks the forward half of its window on integer keys.
//                KS2: the p-value threshold is turned into an integer bound k on
//                     max_v |#{a<=v} - #{b<=v}| on the host (fringe_ks2_critical_count); that
//                     bound holds iff b[i-k] <= a[i] and a[i-k] <= b[i] for all i >= k, so the
//                     device test is 2(N-k) exact integer compares, no merge (ks_bad4).
//                AD2: the inner sum of AD2unique.hpp:287-303 only takes values from a
//                     (2N-1)x(N+1) table of doubles T[j][|2m-(j+1)|]; the table is built on
//                     the host with the reference's expression and the device adds the same
//                     doubles in the same order, so the sum is bit-identical and the
//                     threshold is a single comparison against a host-computed bound.
//   Each unordered pair is tested once (forward half-plane, as nmap.cpp:404-409) and sets both
//   pixels' bits with atomic ORs into a pre-zeroed mask: order-independent, hence deterministic
//   and equal to the race-free reading of the reference loop; counts = popcounts (k_count).
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"

namespace fringe {

// ---------------------------------------------------------------------------------------
// amplitude (nmap.cpp:375: std::abs(complex<float>) / alpha, stored as float) + validity
// (:376) + sort (:389-397).  glibc's hypotf is (float)sqrt((double)re*re + (double)im*im)
// (both products exact in double), reproduced with IEEE double ops so the floats agree
// bit-for-bit with the CPU path.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_amp_sort(const float2* __restrict__ slc,
                                                  const uint8_t* __restrict__ mask,
                                                  const double* __restrict__ alpha, long npix,
                                                  long p0, long pcount, int cols, int apitch, long aplane,
                                                  int bands, float* __restrict__ amp,
                                                  uint8_t* __restrict__ valid) {
    extern __shared__ float s_col[];   // [bands][blockDim.x]
    const int tid = threadIdx.x;
    const int nt = blockDim.x;
    const long p = p0 + (long)blockIdx.x * nt + tid;     // pixels [p0, p0+pcount) of the block
    if (p >= p0 + pcount) return;
    bool ok = mask ? (mask[p] != 0) : true;
    for (int b = 0; b < bands; ++b) {
        const float2 z = __ldg(&slc[(long)b * npix + p]);
        float h;
        if (isinf(z.x) || isinf(z.y)) h = CUDART_INF_F;
        else h = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)z.x, (double)z.x),
                                             __dmul_rn((double)z.y, (double)z.y)));
        // no calibration constants: (float)((double)h / 1.0) == h, skip the double division
        const float v = alpha ? (float)__ddiv_rn((double)h, alpha[b]) : h;
        ok = ok && (v != 0.f) && !isnan(v);
        s_col[b * nt + tid] = v;
    }
    if (ok) {
        for (int i = 1; i < bands; ++i) {
            const float key = s_col[i * nt + tid];
            int j = i - 1;
            while (j >= 0) {
                const float c = s_col[j * nt + tid];
                if (!(c > key)) break;
                s_col[(j + 1) * nt + tid] = c;
                --j;
            }
            s_col[(j + 1) * nt + tid] = key;
        }
    }
    const long ap = (p / cols) * apitch + (p % cols);     // rows of the rank planes are apitch floats apart (TMA: 16-byte rows)
    for (int b = 0; b < bands; ++b) amp[(long)b * aplane + ap] = ok ? s_col[b * nt + tid] : 0.f;
    valid[p] = ok ? 1 : 0;
}

// Batcher's odd-even merge sort as a compare-exchange network on NB register-resident keys (NB a power of two): every
// index is a compile-time constant after unrolling, so the keys never leave the registers, there is no divergence and no
// shared memory.  191 exchanges (two integer min / max each) for NB = 32, where the insertion sort above executes ~7 000
// instructions per pixel with two thirds of the lanes active.  Amplitudes of a valid pixel are positive and NaN-free, so
// their bit patterns order like unsigned integers; unused slots hold 0xFFFFFFFF and stay at the top.
// odd-even merge of the two sorted halves of v[LO .. LO + N) taken with stride R (Batcher)
template <int NB, int LO, int N, int R>
struct OddEvenMerge {
    static __device__ __forceinline__ void run(uint32_t (&v)[NB]) {
        constexpr int M = R * 2;
        if constexpr (M < N) {
            OddEvenMerge<NB, LO, N, M>::run(v);
            OddEvenMerge<NB, LO + R, N, M>::run(v);
#pragma unroll
            for (int i = LO + R; i + R < LO + N; i += M) {
                const uint32_t lo = min(v[i], v[i + R]), hi = max(v[i], v[i + R]);
                v[i] = lo; v[i + R] = hi;
            }
        } else {
            const uint32_t lo = min(v[LO], v[LO + R]), hi = max(v[LO], v[LO + R]);
            v[LO] = lo; v[LO + R] = hi;
        }
    }
};
template <int NB, int LO, int N>
struct OddEvenSort {
    static __device__ __forceinline__ void run(uint32_t (&v)[NB]) {
        if constexpr (N > 1) {
            OddEvenSort<NB, LO, N / 2>::run(v);
            OddEvenSort<NB, LO + N / 2, N / 2>::run(v);
            OddEvenMerge<NB, LO, N, 1>::run(v);
        }
    }
};
template <int NB>
__device__ __forceinline__ void sort_network(uint32_t (&v)[NB]) { OddEvenSort<NB, 0, NB>::run(v); }

// amplitude + validity + sort for bands <= NB: one thread per pixel, everything in registers.  PIXEL_MAJOR: the
// amplitudes are handed over as [pixel][band] floats with the reference's validity mask (nmapProcessBlock).
template <int NB, bool PIXEL_MAJOR>
__global__ void __launch_bounds__(128) k_amp_sort_net(const float2* __restrict__ slc, const float* __restrict__ amp_in,
                                                      const uint8_t* __restrict__ mask, const double* __restrict__ alpha,
                                                      long npix, long p0, long pcount, int cols, int apitch, long aplane,
                                                      int bands, float* __restrict__ amp, uint8_t* __restrict__ valid) {
    // amplitudes in a rolled loop through a shared-memory column (bank = thread), then into registers: unrolling the
    // exact hypot (a double-precision square root with an out-of-line slow path) NB times costs 80 registers and spills
    __shared__ uint32_t s_col[NB][128];
    const int tid = threadIdx.x;
    const long p = p0 + (long)blockIdx.x * blockDim.x + tid;
    if (p >= p0 + pcount) return;
    bool ok = mask ? (mask[p] != 0) : true;
#pragma unroll 4
    for (int b = 0; b < bands; ++b) {
        float x;
        if (PIXEL_MAJOR) {
            x = amp_in[p * bands + b];
        } else {
            const float2 z = __ldg(&slc[(long)b * npix + p]);
            float h;
            if (isinf(z.x) || isinf(z.y)) h = CUDART_INF_F;
            else h = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)z.x, (double)z.x), __dmul_rn((double)z.y, (double)z.y)));
            // no calibration constants: (float)((double)h / 1.0) == h, skip the double division
            x = alpha ? (float)__ddiv_rn((double)h, alpha[b]) : h;
            ok = ok && (x != 0.f) && !isnan(x);
        }
        s_col[b][tid] = __float_as_uint(x);
    }
    uint32_t v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = (b < bands) ? s_col[b][tid] : 0xFFFFFFFFu;
    if (ok) sort_network<NB>(v);
    const long ap = (p / cols) * apitch + (p % cols);
#pragma unroll
    for (int b = 0; b < NB; ++b)
        if (b < bands) amp[(long)b * aplane + ap] = ok ? __uint_as_float(v[b]) : 0.f;
    valid[p] = ok ? 1 : 0;
}

template <bool PIXEL_MAJOR>
static bool launch_sort_net(const float2* slc, const float* amp_in, const uint8_t* mask, const double* alpha, long npix, long p0,
                            long pcount, int cols, int lines, int bands, float* amp, uint8_t* valid, cudaStream_t st) {
    const int apitch = nmap_amp_pitch(cols);
    const long aplane = (long)apitch * lines;
    const unsigned nblk = (unsigned)((pcount + 127) / 128);
#define FRINGE_SORT_CASE(NB_)                                                                                              \
    if (bands <= NB_) {                                                                                                    \
        k_amp_sort_net<NB_, PIXEL_MAJOR><<<nblk, 128, 0, st>>>(slc, amp_in, mask, alpha, npix, p0, pcount, cols, apitch, aplane, \
                                                              bands, amp, valid);                                          \
        return true;                                                                                                       \
    }
    FRINGE_SORT_CASE(8)
    FRINGE_SORT_CASE(16)
    FRINGE_SORT_CASE(32)
    FRINGE_SORT_CASE(64)
#undef FRINGE_SORT_CASE
    return false;                       // more than 64 bands: the shared-memory insertion sort
}

cudaError_t launch_amp_sort(const float2* slc, const uint8_t* mask, const double* alpha, int cols,
                            int lines, int bands, float* amp, uint8_t* valid, int row0, int nrows,
                            cudaStream_t st) {
    const long npix = (long)cols * lines;
    const long p0 = (long)row0 * cols, pcount = (long)nrows * cols;
    if (pcount <= 0) return cudaSuccess;
    if (launch_sort_net<false>(slc, nullptr, mask, alpha, npix, p0, pcount, cols, lines, bands, amp, valid, st)) return cudaGetLastError();
    int nt = 128;
    while (nt > 32 && (size_t)nt * bands * sizeof(float) > 160 * 1024) nt >>= 1;
    const size_t smem = (size_t)nt * bands * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_amp_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    const long nblk = (pcount + nt - 1) / nt;
    const int apitch = nmap_amp_pitch(cols);
    k_amp_sort<<<(unsigned)nblk, nt, smem, st>>>(slc, mask, alpha, npix, p0, pcount, cols, apitch, (long)apitch * lines, bands, amp, valid);
    return cudaGetLastError();
}

// Sort of amplitudes handed over pixel-major ([pixel][band], the reference's arma::fmat amp(N, cols * H)),
// as nmapProcessBlock receives them (src/nmap/nmap_cuda.h:13-17, runSortAmp nmap_cuda.cu:243-255): same
// shared-memory insertion sort, output rank-major like k_amp_sort.  msk is the reference's zeromask.
__global__ void __launch_bounds__(128) k_amp_in_sort(const float* __restrict__ amp_in, const uint8_t* __restrict__ msk,
                                                     long npix, int cols, int apitch, long aplane, int bands,
                                                     float* __restrict__ amp, uint8_t* __restrict__ valid) {
    extern __shared__ float s_col[];   // [bands][blockDim.x]
    const int tid = threadIdx.x, nt = blockDim.x;
    const long p = (long)blockIdx.x * nt + tid;
    if (p >= npix) return;
    const bool ok = msk[p] != 0;
    for (int b = 0; b < bands; ++b) s_col[b * nt + tid] = amp_in[p * bands + b];
    if (ok) {
        for (int i = 1; i < bands; ++i) {
            const float key = s_col[i * nt + tid];
            int j = i - 1;
            while (j >= 0) {
                const float c = s_col[j * nt + tid];
                if (!(c > key)) break;
                s_col[(j + 1) * nt + tid] = c;
                --j;
            }
            s_col[(j + 1) * nt + tid] = key;
        }
    }
    const long ap = (p / cols) * apitch + (p % cols);
    for (int b = 0; b < bands; ++b) amp[(long)b * aplane + ap] = ok ? s_col[b * nt + tid] : 0.f;
    valid[p] = ok ? 1 : 0;
}

cudaError_t launch_amp_in_sort(const float* amp_in, const uint8_t* msk, int cols, int lines, int bands, float* amp, uint8_t* valid,
                               cudaStream_t st) {
    const long npix = (long)cols * lines;
    const int apitch = nmap_amp_pitch(cols);
    if (npix <= 0) return cudaSuccess;
    if (launch_sort_net<true>(nullptr, amp_in, msk, nullptr, npix, 0, npix, cols, lines, bands, amp, valid, st)) return cudaGetLastError();
    int nt = 128;
    while (nt > 32 && (size_t)nt * bands * sizeof(float) > 160 * 1024) nt >>= 1;
    cudaError_t e = cudaFuncSetAttribute(k_amp_in_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    k_amp_in_sort<<<(unsigned)((npix + nt - 1) / nt), nt, (size_t)nt * bands * sizeof(float), st>>>(amp_in, msk, npix, cols, apitch,
                                                                                                  (long)apitch * lines, bands, amp, valid);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// pair tests
// ---------------------------------------------------------------------------------------
struct NmapKernelArgs {
    const float* amp;
    const uint8_t* valid;
    int cols, lines, bands, Nx, Ny, nulong;
    int apitch;               // floats between rows of a rank plane of amp
    int row0, row1;           // output rows [row0, row1) of the block are produced by this launch
    int kcrit;
    double scrit;
    const double* ad_table;   // [(2N-1)][N+1]
    int table_in_smem;
    int32_t* count;
    uint32_t* wts;
};

// The sorted amplitudes are non-negative, NaN-free floats, so their bit patterns order like
// unsigned integers.  The tests below therefore run on uint32 keys: comparisons are integer
// compares and ties are equal bit patterns.  The AD2 merge also keeps a sentinel rank 0xFFFFFFFF
// -- strictly above every real key including +inf -- so it needs no index guards and its loop
// body is a single shared-memory load from a selected address (no divergent branches).
//
// KS2sample.hpp:91-144 without the merge walk.  The reference's statistic is
//   D * n = max_v |#{a <= v} - #{b <= v}|   (v over the pooled values; ties consumed on both sides)
// and the host has already turned the p-value threshold into "D * n <= k".  For ascending a, b:
//   #{a <= v} - #{b <= v} <= k for all v  <=>  (i + 1) - #{b <= a[i]} <= k for all i
//                                         <=>  b[i - k] <= a[i]          for all i >= k
// (the binding v is a[i] itself; for tied a's the last of them is the binding index and the
// earlier ones are weaker), and the mirrored inequality gives a[i - k] <= b[i].  So the pair is
// accepted iff  b[i-k] <= a[i] and a[i-k] <= b[i]  for every i in [k, n): 2 (n - k) integer
// compares instead of a 2n-step merge.  Keys are rank-major with a compile-time plane stride PS
// ([rank][region pixel], as the TMA boxes land), so rank i of a pixel is an immediate offset from
// the pixel's rank-0 address and the lanes of a warp (consecutive pixels) never collide on a bank.
// pa points at rank 0 of the centre pixel.
// Four neighbours of one pixel are tested at once: the centre pixel's two keys of a step are loaded once
// and shared by the four pairs -- 10 shared-memory loads per step for four pairs instead of 16 (the kernel is bound by
// shared-memory bandwidth).  d[j] = offset of neighbour j from the centre pixel in region pixels (the same for every
// thread of the CTA).  Returns bit j set when pair j violates an inequality.
template <int PS>
__device__ __forceinline__ uint32_t ks_bad4(const uint32_t* __restrict__ pa, const int (&d)[4], int n, int k) {
    const uint32_t* __restrict__ lo = pa;
    const uint32_t* __restrict__ hi = pa + k * PS;
    const int m = n - k;
    bool bad0 = false, bad1 = false, bad2 = false, bad3 = false;
    int i = 0;
#pragma unroll 1
    for (; i + 2 <= m; i += 2) {
        uint32_t a_hi[2], a_lo[2], b_hi[4][2], b_lo[4][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a_hi[u] = hi[u * PS]; a_lo[u] = lo[u * PS];
#pragma unroll
            for (int j = 0; j < 4; ++j) { b_hi[j][u] = hi[d[j] + u * PS]; b_lo[j][u] = lo[d[j] + u * PS]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            bad0 |= (b_lo[0][u] > a_hi[u]) | (a_lo[u] > b_hi[0][u]);
            bad1 |= (b_lo[1][u] > a_hi[u]) | (a_lo[u] > b_hi[1][u]);
            bad2 |= (b_lo[2][u] > a_hi[u]) | (a_lo[u] > b_hi[2][u]);
            bad3 |= (b_lo[3][u] > a_hi[u]) | (a_lo[u] > b_hi[3][u]);
        }
        lo += 2 * PS; hi += 2 * PS;
    }
    for (; i < m; ++i) {
        const uint32_t ah = hi[0], al = lo[0];
        bad0 |= (lo[d[0]] > ah) | (al > hi[d[0]]);
        bad1 |= (lo[d[1]] > ah) | (al > hi[d[1]]);
        bad2 |= (lo[d[2]] > ah) | (al > hi[d[2]]);
        bad3 |= (lo[d[3]] > ah) | (al > hi[d[3]]);
        lo += PS; hi += PS;
    }
    return (bad0 ? 1u : 0u) | (bad1 ? 2u : 0u) | (bad2 ? 4u : 0u) | (bad3 ? 8u : 0u);
}

// AD2unique.hpp:211-303: merge (on equality the element of B goes first) and table sum; the pixel at ia0 must be the one
// that comes first in raster order.
// Four merges at once, for four neighbours of one pixel: the walk of one pair is a chain of dependent shared-memory loads
// (compare -> select -> load -> compare ...), and at 21x21 x 30 bands the tile leaves room for only 8 warps per SM, so a
// single walk per thread leaves the SM waiting on that chain.  Four independent walks per thread fill it.  Each walk
// adds the reference's terms in the reference's order: the sums are bit-identical.
constexpr int AD_BATCH = 4;          // walks per thread; C5 strip: 1 -> 54.1 ms, 4 -> 42.9 ms, 8 -> 43.9 ms
__device__ __forceinline__ void ad_inner_sums(const uint32_t* __restrict__ s_key, uint32_t ia0, const uint32_t (&ibq)[AD_BATCH], int n,
                                              uint32_t stride, const double* __restrict__ T, double (&S)[AD_BATCH]) {
    uint32_t ia[AD_BATCH], ib[AD_BATCH], va[AD_BATCH], vb[AD_BATCH];
    int m2[AD_BATCH];
#pragma unroll
    for (int q = 0; q < AD_BATCH; ++q) { ia[q] = ia0; ib[q] = ibq[q]; va[q] = s_key[ia0]; vb[q] = s_key[ibq[q]]; m2[q] = 0; S[q] = 0.0; }
    const double* Tj = T;
    for (int j = 0; j < 2 * n - 1; ++j) {
#pragma unroll
        for (int q = 0; q < AD_BATCH; ++q) {
            const bool ta = va[q] < vb[q];
            ia[q] += ta ? stride : 0u;
            ib[q] += ta ? 0u : stride;
            m2[q] += ta ? 1 : -1;
            const uint32_t nxt = s_key[ta ? ia[q] : ib[q]];
            va[q] = ta ? nxt : va[q];
            vb[q] = ta ? vb[q] : nxt;
            S[q] = __dadd_rn(S[q], Tj[abs(m2[q])]);
        }
        Tj += n + 1;
    }
}

// ---- TMA / mbarrier (sm_90+ PTX) --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one box of the 3-D tensor (x = column, y = line, z = rank) into shared memory; completion counted on bar
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// Forward half-plane pair tests (nmap.cpp:404-473): the thread of pixel p tests the neighbours q
// that follow it in raster order inside the window; an accepted pair sets bit (dy,dx) of p and
// bit (-dy,-dx) of q, exactly like the reference's symmetric update, but with atomic ORs into the
// (pre-zeroed) global mask so the result does not depend on scheduling.  Neighbour counts are
// the popcounts of the finished masks (k_count).
// PS = words between the rank planes in shared memory (>= region pixels, a multiple of 32).
template <int METHOD, int PS>
__global__ void k_nmap(const NmapKernelArgs a, const __grid_constant__ CUtensorMap amp_map) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    const int N = a.bands, Nx = a.Nx, Ny = a.Ny;
    const int TW = blockDim.x, TH = blockDim.y;
    const int RW = (TW + 2 * Nx + 3) & ~3, RH = TH + Ny;         // tile + forward halo; box rows are whole 16-byte units
    const int tid = threadIdx.y * TW + threadIdx.x, nthr = TW * TH;

    // carve: [uint32 keys (N+1) planes of PS words] [double table (optional)]
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_raw);
    double* s_tab = reinterpret_cast<double*>(s_key + (size_t)(N + 1) * PS);
    const int tab_elems = (METHOD == 1 && a.table_in_smem) ? (2 * N - 1) * (N + 1) : 0;

    // A box must start on a 16-byte boundary of its row (measured: an unaligned innermost coordinate faults, negative
    // and out-of-image boxes are fine and come back zero-filled -- scripts/probes/tma_probe.cu), so the tile grid is
    // shifted left by (-Nx) mod 4 columns: tile origin - Nx is then a multiple of 4 for every CTA.
    const int xshift = (4 - (Nx & 3)) & 3;
    const int x0 = blockIdx.x * TW - xshift - Nx, y0 = a.row0 + blockIdx.y * TH;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, (uint32_t)(N * RW * RH * (int)sizeof(float)));
        for (int k = 0; k < N; ++k) tma_load_3d(s_key + (size_t)k * PS, &amp_map, x0, y0, k, &s_bar);
    }
    // meanwhile: the sentinel rank of the AD2 merge and the term table
    if (METHOD == 1) for (int rp = tid; rp < RW * RH; rp += nthr) s_key[(size_t)N * PS + rp] = 0xFFFFFFFFu;
    for (int i = tid; i < tab_elems; i += nthr) s_tab[i] = a.ad_table[i];
    __syncthreads();
    mbar_wait(&s_bar, 0);

    const int gx = blockIdx.x * TW - xshift + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx < 0 || gx >= a.cols || gy >= a.row1) return;
    const int rp = threadIdx.y * RW + threadIdx.x + Nx;
    const uint32_t* s_top = s_key + (size_t)(N - 1) * PS;       // largest amplitude: zero <=> invalid or outside the image
    if (!s_top[rp]) return;                         // mask words stay zero
    const long p = (long)gy * a.cols + gx;
    uint32_t* wp = a.wts + p * a.nulong;
    const int WX = 2 * Nx + 1, W = WX * (2 * Ny + 1), center = Ny * WX + Nx;
    uint32_t word = 1u << (center & 31);            // a valid pixel is always its own neighbour
    const int kc = min(max(a.kcrit, 0), N);         // k >= N accepts every pair, k < 0 none
    if (METHOD == 0) {
        // KS2: four window positions per pass (ks_bad4).  Positions beyond the window and invalid neighbours are tested
        // against the pixel itself (offset 0) and their result is dropped.
        int dy = 0, dx = 1;
        for (int f0 = center + 1; f0 < W; f0 += 4) {
            int d[4], ddy[4], ddx[4];
            uint32_t live = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (dx > Nx) { dx = -Nx; ++dy; }
                ddy[j] = dy; ddx[j] = dx;
                const bool in_window = (f0 + j < W);
                d[j] = in_window ? dy * RW + dx : 0;
                if (in_window && s_top[rp + d[j]]) live |= 1u << j;
                ++dx;
            }
            uint32_t bad = 0xFu;
            if (a.kcrit >= 0) bad = ks_bad4<PS>(s_key + rp, d, N, kc);
            const uint32_t ok = live & ~bad;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = f0 + j;
                if (f < W && (f & 31) == 0) { atomicOr(&wp[(f - 1) >> 5], word); word = 0u; }
                if ((ok >> j) & 1u) {
                    word |= (1u << (f & 31));
                    const int fm = W - 1 - f;               // bit of (-dy,-dx) in q's mask
                    atomicOr(&a.wts[(p + (long)ddy[j] * a.cols + ddx[j]) * a.nulong + (fm >> 5)], 1u << (fm & 31));
                }
            }
        }
    } else {
        // (the walk is instantiated once per address space of the term table, so that its loads are plain LDS / LDG
        // instead of generic loads)
        auto walk_window = [&](const double* __restrict__ Tq) {
        // AD2: four window positions per pass (ad_inner_sums); positions beyond the window and invalid neighbours walk
        // against the pixel itself and their result is dropped
        int dy = 0, dx = 1;
        for (int f0 = center + 1; f0 < W; f0 += AD_BATCH) {
            uint32_t rq[AD_BATCH];
            int ddy[AD_BATCH], ddx[AD_BATCH];
            uint32_t live = 0u;
#pragma unroll
            for (int j = 0; j < AD_BATCH; ++j) {
                if (dx > Nx) { dx = -Nx; ++dy; }
                ddy[j] = dy; ddx[j] = dx;
                const bool in_window = (f0 + j < W);
                rq[j] = in_window ? rp + dy * RW + dx : rp;
                if (in_window && s_top[rq[j]]) live |= 1u << j;
                ++dx;
            }
            double S[AD_BATCH];
            #pragma unroll
            for (int j = 0; j < AD_BATCH; ++j) S[j] = 0.0;
            if (live != 0u) ad_inner_sums(s_key, rp, rq, N, PS, Tq, S);
#pragma unroll
            for (int j = 0; j < AD_BATCH; ++j) {
                const int f = f0 + j;
                if (f < W && (f & 31) == 0) { atomicOr(&wp[(f - 1) >> 5], word); word = 0u; }
                if (((live >> j) & 1u) && S[j] <= a.scrit) {
                    word |= (1u << (f & 31));
                    const int fm = W - 1 - f;               // bit of (-dy,-dx) in q's mask
                    atomicOr(&a.wts[(p + (long)ddy[j] * a.cols + ddx[j]) * a.nulong + (fm >> 5)], 1u << (fm & 31));
                }
            }
        }
        };
        if (a.table_in_smem) walk_window(s_tab); else walk_window(a.ad_table);
    }
    atomicOr(&wp[(W - 1) >> 5], word);
}

// The same pair tests straight from global memory, for band counts x windows whose tile does not fit
// shared memory (e.g. 200 dates with the 59 x 19 window of the sequential workflow): one thread per
// pixel, keys read rank-major (coalesced across the threads of a warp, served by L1 / L2).  Slower than
// the staged kernel, identical results.
template <int METHOD>
__global__ void __launch_bounds__(128) k_nmap_global(const NmapKernelArgs a) {
    const int N = a.bands, Nx = a.Nx, Ny = a.Ny;
    const long plane = (long)a.apitch * a.lines;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = a.row0 + blockIdx.y;
    if (gx >= a.cols || gy >= a.row1) return;
    const long p = (long)gy * a.cols + gx;
    if (!a.valid[p]) return;
    const long pa = (long)gy * a.apitch + gx;
    const uint32_t* __restrict__ key = reinterpret_cast<const uint32_t*>(a.amp);
    uint32_t* wp = a.wts + p * a.nulong;
    const int WX = 2 * Nx + 1, W = WX * (2 * Ny + 1), center = Ny * WX + Nx;
    uint32_t word = 1u << (center & 31);
    const int kc = min(max(a.kcrit, 0), N);
    int dy = 0, dx = 1;
    for (int f = center + 1; f < W; ++f) {
        if (dx > Nx) { dx = -Nx; ++dy; }
        if ((f & 31) == 0) { atomicOr(&wp[(f - 1) >> 5], word); word = 0u; }
        const int qy = gy + dy, qx = gx + dx;
        if (qy < a.lines && qx >= 0 && qx < a.cols) {
            const long q = (long)qy * a.cols + qx;
            const long qa = (long)qy * a.apitch + qx;
            if (a.valid[q]) {
                bool similar;
                if (METHOD == 0) {
                    bool bad = a.kcrit < 0;
                    for (int i = kc; i < N && !bad; ++i) {
                        const uint32_t a_hi = __ldg(&key[(long)i * plane + pa]), a_lo = __ldg(&key[(long)(i - kc) * plane + pa]);
                        const uint32_t b_hi = __ldg(&key[(long)i * plane + qa]), b_lo = __ldg(&key[(long)(i - kc) * plane + qa]);
                        bad = (b_lo > a_hi) | (a_lo > b_hi);
                    }
                    similar = !bad;
                } else {
                    // AD2unique.hpp:211-303 merge (ties: the element of B first), table sum in the reference's order
                    int ia = 0, ib = 0, m2 = 0;
                    uint32_t va = __ldg(&key[pa]), vb = __ldg(&key[qa]);
                    double S = 0.0;
                    const double* Tj = a.ad_table;
                    for (int j = 0; j < 2 * N - 1; ++j) {
                        const bool ta = (ib >= N) || (ia < N && va < vb);
                        if (ta) { ++ia; ++m2; va = (ia < N) ? __ldg(&key[(long)ia * plane + pa]) : 0xFFFFFFFFu; }
                        else { ++ib; --m2; vb = (ib < N) ? __ldg(&key[(long)ib * plane + qa]) : 0xFFFFFFFFu; }
                        S = __dadd_rn(S, Tj[abs(m2)]);
                        Tj += N + 1;
                    }
                    similar = S <= a.scrit;
                }
                if (similar) {
                    word |= (1u << (f & 31));
                    const int fm = W - 1 - f;
                    atomicOr(&a.wts[q * a.nulong + (fm >> 5)], 1u << (fm & 31));
                }
            }
        }
        ++dx;
    }
    atomicOr(&wp[(W - 1) >> 5], word);
}

// neighbour count = number of set bits (count(pp) is incremented exactly once per bit set,
// nmap.cpp:447-468)
__global__ void __launch_bounds__(256) k_count(const uint32_t* __restrict__ wts, int nulong, long p0,
                                               long pend, int32_t* __restrict__ count) {
    const long p = p0 + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pend) return;
    int c = 0;
    for (int w = 0; w < nulong; ++w) c += __popc(wts[p * nulong + w]);
    count[p] = c;
}

// rows of the rank planes of amp are padded to whole 16-byte units (a TMA tensor map needs 16-byte strides)
int nmap_amp_pitch(int cols) { return (cols + 3) & ~3; }

// shared-memory plane strides (words) the staged kernel is instantiated for
static const int kPlaneStrides[] = {448, 704, 960, 1984};

static size_t nmap_smem(int bands, int ps, bool table) {
    size_t b = (size_t)(bands + 1) * ps * sizeof(uint32_t);
    if (table) b += (size_t)(2 * bands - 1) * (bands + 1) * sizeof(double);
    return (b + 15) & ~(size_t)15;
}

bool nmap_plan(int bands, int Nx, int Ny, int method, NmapGeometry* g) {
    const size_t budget = 200 * 1024;
    const size_t tab = (size_t)(2 * bands - 1) * (bands + 1) * sizeof(double);
    static const int shapes[][2] = {{32, 8}, {32, 4}, {32, 2}, {32, 1}, {16, 2}, {16, 1}, {8, 1}};
    for (auto& s : shapes) {
        const int rw = (s[0] + 2 * Nx + 3) & ~3, rh = s[1] + Ny;
        if (rw > 256 || rh > 256) continue;                       // TMA box limits
        int ps = 0;
        for (int c : kPlaneStrides) if (rw * rh <= c) { ps = c; break; }
        if (!ps) continue;
        for (int t = (method == 1 ? 1 : 0); t >= 0; --t) {
            const bool table = (t == 1) && tab <= 64 * 1024;
            const size_t need = nmap_smem(bands, ps, table);
            if (need <= budget) {
                g->tile_w = s[0]; g->tile_h = s[1]; g->plane_stride = ps; g->smem_bytes = need; g->table_in_smem = table;
                return true;
            }
        }
    }
    // no tile fits: the global-memory kernel (tile_w = 0 marks it)
    g->tile_w = 0; g->tile_h = 0; g->plane_stride = 0; g->smem_bytes = 0; g->table_in_smem = false;
    return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// amp as a 3-D tensor (column, line, rank) of 32-bit keys; one box = tile + halo of one rank
static cudaError_t make_amp_map(const float* amp, int cols, int lines, int bands, int box_w, int box_h, CUtensorMap* map) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t pitch = (cuuint64_t)nmap_amp_pitch(cols);
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)lines, (cuuint64_t)bands};
    const cuuint64_t strides[2] = {pitch * sizeof(float), pitch * (cuuint64_t)lines * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<float*>(amp), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int METHOD, int PS>
static cudaError_t launch_nmap_staged(const NmapKernelArgs& a, const CUtensorMap& map, dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_nmap<METHOD, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_nmap<METHOD, PS><<<grid, block, smem, st>>>(a, map);
    return cudaGetLastError();
}

cudaError_t launch_nmap(const float* amp, const uint8_t* valid, int cols, int lines, int bands,
                        int Nx, int Ny, int method, int kcrit, double scrit,
                        const double* ad_table, const NmapGeometry& g, int32_t* count,
                        uint32_t* wts, int row0, int nrows, cudaStream_t st) {
    if (nrows <= 0) return cudaSuccess;
    NmapKernelArgs a;
    a.row0 = row0; a.row1 = row0 + nrows;
    a.amp = amp; a.valid = valid; a.cols = cols; a.lines = lines; a.bands = bands;
    a.Nx = Nx; a.Ny = Ny; a.nulong = ((2 * Ny + 1) * (2 * Nx + 1) + 31) / 32;
    a.apitch = nmap_amp_pitch(cols);
    a.kcrit = kcrit; a.scrit = scrit; a.ad_table = ad_table; a.table_in_smem = g.table_in_smem ? 1 : 0;
    a.count = count; a.wts = wts;
    if (g.tile_w == 0) {                                   // tile does not fit shared memory
        dim3 gblock(128), ggrid((cols + 127) / 128, nrows);
        if (method == 0) k_nmap_global<0><<<ggrid, gblock, 0, st>>>(a);
        else k_nmap_global<1><<<ggrid, gblock, 0, st>>>(a);
        return cudaGetLastError();
    }
    CUtensorMap map;
    cudaError_t e = make_amp_map(amp, cols, lines, bands, (g.tile_w + 2 * Nx + 3) & ~3, g.tile_h + Ny, &map);
    if (e != cudaSuccess) return e;
    dim3 block(g.tile_w, g.tile_h);
    const int xshift = (4 - (Nx & 3)) & 3;                  // see k_nmap: tile origins shifted so that boxes start 16-byte aligned
    dim3 grid((cols + xshift + g.tile_w - 1) / g.tile_w, (nrows + g.tile_h - 1) / g.tile_h);
#define FRINGE_NMAP_CASE(PS_)                                                                              \
    case PS_:                                                                                              \
        return method == 0 ? launch_nmap_staged<0, PS_>(a, map, grid, block, g.smem_bytes, st)              \
                           : launch_nmap_staged<1, PS_>(a, map, grid, block, g.smem_bytes, st);
    switch (g.plane_stride) {
        FRINGE_NMAP_CASE(448)
        FRINGE_NMAP_CASE(704)
        FRINGE_NMAP_CASE(960)
        FRINGE_NMAP_CASE(1984)
    }
#undef FRINGE_NMAP_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_count(const uint32_t* wts, int cols, int nulong, int row0, int nrows, int32_t* count,
                         cudaStream_t st) {
    if (nrows <= 0) return cudaSuccess;
    const long p0 = (long)row0 * cols, pend = p0 + (long)nrows * cols;
    k_count<<<(unsigned)((pend - p0 + 255) / 256), 256, 0, st>>>(wts, nulong, p0, pend, count);
    return cudaGetLastError();
}

}  // namespace fringe
