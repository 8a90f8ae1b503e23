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
//   k_nmap<M>    one CTA per tile of pixels; the sorted vectors of tile + forward halo are
//                staged in shared memory once as integer keys; one thread per pixel walks the
//                forward half of its window.
//                KS2: the p-value threshold is turned into an integer bound k on
//                     max_v |#{a<=v} - #{b<=v}| on the host (fringe_ks2_critical_count); that
//                     bound holds iff b[i-k] <= a[i] and a[i-k] <= b[i] for all i >= k, so the
//                     device test is 2(N-k) exact integer compares, no merge (ks_within).
//                AD2: the inner sum of AD2unique.hpp:287-303 only takes values from a
//                     (2N-1)x(N+1) table of doubles T[j][|2m-(j+1)|]; the table is built on
//                     the host with the reference's expression and the device adds the same
//                     doubles in the same order, so the sum is bit-identical and the
//                     threshold is a single comparison against a host-computed bound.
//   Each unordered pair is tested once (forward half-plane, as nmap.cpp:404-409) and sets both
//   pixels' bits with atomic ORs into a pre-zeroed mask: order-independent, hence deterministic
//   and equal to the race-free reading of the reference loop; counts = popcounts (k_count).
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
                                                  long p0, long pcount,
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
    for (int b = 0; b < bands; ++b) amp[(long)b * npix + p] = ok ? s_col[b * nt + tid] : 0.f;
    valid[p] = ok ? 1 : 0;
}

cudaError_t launch_amp_sort(const float2* slc, const uint8_t* mask, const double* alpha, int cols,
                            int lines, int bands, float* amp, uint8_t* valid, int row0, int nrows,
                            cudaStream_t st) {
    const long npix = (long)cols * lines;
    const long p0 = (long)row0 * cols, pcount = (long)nrows * cols;
    if (pcount <= 0) return cudaSuccess;
    int nt = 128;
    while (nt > 32 && (size_t)nt * bands * sizeof(float) > 160 * 1024) nt >>= 1;
    const size_t smem = (size_t)nt * bands * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_amp_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    const long nblk = (pcount + nt - 1) / nt;
    k_amp_sort<<<(unsigned)nblk, nt, smem, st>>>(slc, mask, alpha, npix, p0, pcount, bands, amp, valid);
    return cudaGetLastError();
}

// Sort of amplitudes handed over pixel-major ([pixel][band], the reference's arma::fmat amp(N, cols * H)),
// as nmapProcessBlock receives them (src/nmap/nmap_cuda.h:13-17, runSortAmp nmap_cuda.cu:243-255): same
// shared-memory insertion sort, output rank-major like k_amp_sort.  msk is the reference's zeromask.
__global__ void __launch_bounds__(128) k_amp_in_sort(const float* __restrict__ amp_in, const uint8_t* __restrict__ msk,
                                                     long npix, int bands, float* __restrict__ amp,
                                                     uint8_t* __restrict__ valid) {
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
    for (int b = 0; b < bands; ++b) amp[(long)b * npix + p] = ok ? s_col[b * nt + tid] : 0.f;
    valid[p] = ok ? 1 : 0;
}

cudaError_t launch_amp_in_sort(const float* amp_in, const uint8_t* msk, long npix, int bands, float* amp, uint8_t* valid,
                               cudaStream_t st) {
    if (npix <= 0) return cudaSuccess;
    int nt = 128;
    while (nt > 32 && (size_t)nt * bands * sizeof(float) > 160 * 1024) nt >>= 1;
    cudaError_t e = cudaFuncSetAttribute(k_amp_in_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    k_amp_in_sort<<<(unsigned)((npix + nt - 1) / nt), nt, (size_t)nt * bands * sizeof(float), st>>>(amp_in, msk, npix, bands, amp, valid);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// pair tests
// ---------------------------------------------------------------------------------------
struct NmapKernelArgs {
    const float* amp;
    const uint8_t* valid;
    int cols, lines, bands, Nx, Ny, nulong;
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
// compares on keys each pixel keeps contiguously ([pixel][rank], odd stride), instead of a 2n-step
// merge.  pa / pb point at rank 0 of the two pixels.
__device__ __forceinline__ bool ks_within(const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                                          int n, int k) {
    const uint32_t* __restrict__ ah = pa + k;
    const uint32_t* __restrict__ bh = pb + k;
    const int m = n - k;
    bool bad = false;                     // no short circuit: all loads of a step issue together
    int i = 0;
#pragma unroll 1
    for (; i + 4 <= m; i += 4) {
        uint32_t a_hi[4], a_lo[4], b_hi[4], b_lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a_hi[u] = ah[i + u]; a_lo[u] = pa[i + u]; b_hi[u] = bh[i + u]; b_lo[u] = pb[i + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) bad |= (b_lo[u] > a_hi[u]) | (a_lo[u] > b_hi[u]);
    }
    for (; i < m; ++i) bad |= (pb[i] > ah[i]) | (pa[i] > bh[i]);
    return !bad;
}

// AD2unique.hpp:211-303: merge (on equality the element of B goes first) and table sum.
// The pixel at ia must be the one that comes first in raster order.
__device__ __forceinline__ double ad_inner_sum(const uint32_t* __restrict__ s_key, uint32_t ia,
                                               uint32_t ib, int n, uint32_t stride,
                                               const double* __restrict__ T) {
    uint32_t va = s_key[ia], vb = s_key[ib];
    double S = 0.0;
    int m2 = 0;                          // 2 * (#a consumed) - (j+1)
    const double* Tj = T;
#pragma unroll 2
    for (int j = 0; j < 2 * n - 1; ++j) {
        const bool ta = va < vb;
        ia += ta ? stride : 0u;
        ib += ta ? 0u : stride;
        m2 += ta ? 1 : -1;
        const uint32_t nxt = s_key[ta ? ia : ib];
        va = ta ? nxt : va;
        vb = ta ? vb : nxt;
        S = __dadd_rn(S, Tj[abs(m2)]);
        Tj += n + 1;
    }
    return S;
}

// Forward half-plane pair tests (nmap.cpp:404-473): the thread of pixel p tests the neighbours q
// that follow it in raster order inside the window; an accepted pair sets bit (dy,dx) of p and
// bit (-dy,-dx) of q, exactly like the reference's symmetric update, but with atomic ORs into the
// (pre-zeroed) global mask so the result does not depend on scheduling.  Neighbour counts are
// the popcounts of the finished masks (k_count).
template <int METHOD>
__global__ void k_nmap(const NmapKernelArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int N = a.bands, Nx = a.Nx, Ny = a.Ny;
    const int TW = blockDim.x, TH = blockDim.y;
    const int RW = TW + 2 * Nx, RH = TH + Ny, RP = RW * RH;       // tile + forward halo
    const int tid = threadIdx.y * TW + threadIdx.x, nthr = TW * TH;
    const long npix = (long)a.cols * a.lines;

    // carve: [double table (optional)] [uint32 keys (N+1)*RP] [uint8 valid RP]
    // key layout: AD2 [rank][region pixel] + sentinel rank; KS2 [region pixel][rank], odd stride KS
    double* s_tab = reinterpret_cast<double*>(s_raw);
    const int tab_elems = (METHOD == 1 && a.table_in_smem) ? (2 * N - 1) * (N + 1) : 0;
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_tab + tab_elems);
    uint8_t* s_valid = reinterpret_cast<uint8_t*>(s_key + (size_t)(N + 1) * RP);
    const int KS = N | 1;

    const int x0 = blockIdx.x * TW - Nx, y0 = a.row0 + blockIdx.y * TH;
    for (int rp = tid; rp < RP; rp += nthr) {
        const int gy = y0 + rp / RW, gx = x0 + rp % RW;
        const bool inb = (gy < a.lines) && (gx >= 0) && (gx < a.cols);
        s_valid[rp] = inb ? a.valid[(long)gy * a.cols + gx] : 0;
        if (METHOD == 1) s_key[(size_t)N * RP + rp] = 0xFFFFFFFFu;
    }
    for (int idx = tid; idx < N * RP; idx += nthr) {
        const int k = idx / RP, rp = idx - k * RP;
        const int gy = y0 + rp / RW, gx = x0 + rp % RW;
        const bool inb = (gy < a.lines) && (gx >= 0) && (gx < a.cols);
        const uint32_t key = inb ? __float_as_uint(__ldg(&a.amp[(long)k * npix + (long)gy * a.cols + gx])) : 0u;
        if (METHOD == 1) s_key[idx] = key;
        else s_key[rp * KS + k] = key;
    }
    for (int i = tid; i < tab_elems; i += nthr) s_tab[i] = a.ad_table[i];
    __syncthreads();

    const int gx = blockIdx.x * TW + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= a.cols || gy >= a.row1) return;
    const int rp = threadIdx.y * RW + threadIdx.x + Nx;
    if (!s_valid[rp]) return;                       // mask words stay zero
    const long p = (long)gy * a.cols + gx;
    uint32_t* wp = a.wts + p * a.nulong;
    const double* T = (METHOD == 1) ? (a.table_in_smem ? s_tab : a.ad_table) : nullptr;
    const int WX = 2 * Nx + 1, W = WX * (2 * Ny + 1), center = Ny * WX + Nx;
    uint32_t word = 1u << (center & 31);            // a valid pixel is always its own neighbour
    const int kc = min(max(a.kcrit, 0), N);         // k >= N accepts every pair, k < 0 none
    int dy = 0, dx = 1;
    for (int f = center + 1; f < W; ++f) {
        if (dx > Nx) { dx = -Nx; ++dy; }
        if ((f & 31) == 0) { atomicOr(&wp[(f - 1) >> 5], word); word = 0u; }
        const int rq = rp + dy * RW + dx;
        if (s_valid[rq]) {
            bool similar;
            if (METHOD == 0) similar = (a.kcrit >= 0) && ks_within(s_key + rp * KS, s_key + rq * KS, N, kc);
            else similar = ad_inner_sum(s_key, rp, rq, N, RP, T) <= a.scrit;
            if (similar) {
                word |= (1u << (f & 31));
                const int fm = W - 1 - f;               // bit of (-dy,-dx) in q's mask
                atomicOr(&a.wts[(p + (long)dy * a.cols + dx) * a.nulong + (fm >> 5)], 1u << (fm & 31));
            }
        }
        ++dx;
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
    const long npix = (long)a.cols * a.lines;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = a.row0 + blockIdx.y;
    if (gx >= a.cols || gy >= a.row1) return;
    const long p = (long)gy * a.cols + gx;
    if (!a.valid[p]) return;
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
            if (a.valid[q]) {
                bool similar;
                if (METHOD == 0) {
                    bool bad = a.kcrit < 0;
                    for (int i = kc; i < N && !bad; ++i) {
                        const uint32_t a_hi = __ldg(&key[(long)i * npix + p]), a_lo = __ldg(&key[(long)(i - kc) * npix + p]);
                        const uint32_t b_hi = __ldg(&key[(long)i * npix + q]), b_lo = __ldg(&key[(long)(i - kc) * npix + q]);
                        bad = (b_lo > a_hi) | (a_lo > b_hi);
                    }
                    similar = !bad;
                } else {
                    // AD2unique.hpp:211-303 merge (ties: the element of B first), table sum in the reference's order
                    int ia = 0, ib = 0, m2 = 0;
                    uint32_t va = __ldg(&key[p]), vb = __ldg(&key[q]);
                    double S = 0.0;
                    const double* Tj = a.ad_table;
                    for (int j = 0; j < 2 * N - 1; ++j) {
                        const bool ta = (ib >= N) || (ia < N && va < vb);
                        if (ta) { ++ia; ++m2; va = (ia < N) ? __ldg(&key[(long)ia * npix + p]) : 0xFFFFFFFFu; }
                        else { ++ib; --m2; vb = (ib < N) ? __ldg(&key[(long)ib * npix + q]) : 0xFFFFFFFFu; }
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

static size_t nmap_smem(int bands, int Nx, int Ny, int tw, int th, bool table) {
    const size_t RP = (size_t)(tw + 2 * Nx) * (th + Ny);
    size_t b = (size_t)(bands + 1) * RP * sizeof(float) + RP;
    if (table) b += (size_t)(2 * bands - 1) * (bands + 1) * sizeof(double);
    return (b + 15) & ~(size_t)15;
}

bool nmap_plan(int bands, int Nx, int Ny, int method, NmapGeometry* g) {
    const size_t budget = 200 * 1024;
    const size_t tab = (size_t)(2 * bands - 1) * (bands + 1) * sizeof(double);
    static const int shapes[][2] = {{32, 8}, {32, 4}, {32, 2}, {32, 1}, {16, 2}, {16, 1}, {8, 1}};
    for (auto& s : shapes) {
        for (int t = (method == 1 ? 1 : 0); t >= 0; --t) {
            const bool table = (t == 1) && tab <= 64 * 1024;
            const size_t need = nmap_smem(bands, Nx, Ny, s[0], s[1], table);
            if (need <= budget) {
                g->tile_w = s[0]; g->tile_h = s[1]; g->smem_bytes = need; g->table_in_smem = table;
                return true;
            }
        }
    }
    // no tile fits: the global-memory kernel (tile_w = 0 marks it)
    g->tile_w = 0; g->tile_h = 0; g->smem_bytes = 0; g->table_in_smem = false;
    return true;
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
    a.kcrit = kcrit; a.scrit = scrit; a.ad_table = ad_table; a.table_in_smem = g.table_in_smem ? 1 : 0;
    a.count = count; a.wts = wts;
    cudaError_t e;
    if (g.tile_w == 0) {                                   // tile does not fit shared memory
        dim3 gblock(128), ggrid((cols + 127) / 128, nrows);
        if (method == 0) k_nmap_global<0><<<ggrid, gblock, 0, st>>>(a);
        else k_nmap_global<1><<<ggrid, gblock, 0, st>>>(a);
        return cudaGetLastError();
    }
    dim3 block(g.tile_w, g.tile_h);
    dim3 grid((cols + g.tile_w - 1) / g.tile_w, (nrows + g.tile_h - 1) / g.tile_h);
    if (method == 0) {
        e = cudaFuncSetAttribute(k_nmap<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
        if (e != cudaSuccess) return e;
        k_nmap<0><<<grid, block, g.smem_bytes, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(k_nmap<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
        if (e != cudaSuccess) return e;
        k_nmap<1><<<grid, block, g.smem_bytes, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_count(const uint32_t* wts, int cols, int nulong, int row0, int nrows, int32_t* count,
                         cudaStream_t st) {
    if (nrows <= 0) return cudaSuccess;
    const long p0 = (long)row0 * cols, pend = p0 + (long)nrows * cols;
    k_count<<<(unsigned)((pend - p0 + 255) / 256), 256, 0, st>>>(wts, nulong, p0, pend, count);
    return cudaGetLastError();
}

}  // namespace fringe
