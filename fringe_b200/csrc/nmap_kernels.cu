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
//   k_nmap<M>    one CTA per tile of output pixels; the sorted vectors of tile+halo are staged
//                in shared memory once ([rank][region pixel], + one +inf sentinel rank); one
//                thread per output pixel walks its window and runs a branch-free merge per
//                neighbour.
//                KS2: the p-value threshold is turned into an integer bound on
//                     max_v |#{a<=v} - #{b<=v}| on the host (fringe_ks2_critical_count), so
//                     the device test is exact integer arithmetic.
//                AD2: the inner sum of AD2unique.hpp:287-303 only takes values from a
//                     (2N-1)x(N+1) table of doubles T[j][|2m-(j+1)|]; the table is built on
//                     the host with the reference's expression and the device adds the same
//                     doubles in the same order, so the sum is bit-identical and the
//                     threshold is a single comparison against a host-computed bound.
//   Each pixel tests its whole window and writes only its own words: no atomics, results
//   are deterministic and equal to the race-free reading of the reference loop.
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
        const double al = alpha ? alpha[b] : 1.0;
        const float v = (float)__ddiv_rn((double)h, al);
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

// KS2sample.hpp:91-144 as an integer walk.  A/B point at rank 0 of the two pixels, consecutive
// ranks are `stride` floats apart, rank `n` holds +inf.  Returns max |#b - #a| sampled only
// where every element equal to the last consumed value has been consumed on both sides.
__device__ __forceinline__ int ks_max_count_diff(const float* __restrict__ A,
                                                 const float* __restrict__ B, int n, int stride) {
    int ia = 0, ib = 0, kmax = 0;
    float va = A[0], vb = B[0];
    for (int s = 0; s < 2 * n; ++s) {
        const bool ta = (ib >= n) || ((ia < n) && (va <= vb));
        const float x = ta ? va : vb;
        ia += ta ? 1 : 0;
        ib += ta ? 0 : 1;
        const float nxt = ta ? A[ia * stride] : B[ib * stride];
        va = ta ? nxt : va;
        vb = ta ? vb : nxt;
        const int k = abs(ib - ia);
        kmax = ((va > x) && (vb > x)) ? max(kmax, k) : kmax;
    }
    return kmax;
}

// AD2unique.hpp:211-303: merge (on equality the element of B goes first) and table sum.
// A must be the pixel that comes first in raster order.
__device__ __forceinline__ double ad_inner_sum(const float* __restrict__ A,
                                               const float* __restrict__ B, int n, int stride,
                                               const double* __restrict__ T) {
    int ia = 0, ib = 0;
    float va = A[0], vb = B[0];
    double S = 0.0;
    const int L = 2 * n;
    for (int j = 0; j < L - 1; ++j) {
        const bool ta = (ib >= n) || ((ia < n) && (va < vb));
        ia += ta ? 1 : 0;
        ib += ta ? 0 : 1;
        const float nxt = ta ? A[ia * stride] : B[ib * stride];
        va = ta ? nxt : va;
        vb = ta ? vb : nxt;
        const int u = abs(2 * ia - (j + 1));
        S = __dadd_rn(S, T[j * (n + 1) + u]);
    }
    return S;
}

template <int METHOD>
__global__ void k_nmap(const NmapKernelArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int N = a.bands, Nx = a.Nx, Ny = a.Ny;
    const int TW = blockDim.x, TH = blockDim.y;
    const int RW = TW + 2 * Nx, RH = TH + 2 * Ny, RP = RW * RH;
    const int tid = threadIdx.y * TW + threadIdx.x, nthr = TW * TH;
    const long npix = (long)a.cols * a.lines;

    // carve: [double table (optional)] [float amps (N+1)*RP] [uint8 valid RP]
    double* s_tab = reinterpret_cast<double*>(s_raw);
    const int tab_elems = (METHOD == 1 && a.table_in_smem) ? (2 * N - 1) * (N + 1) : 0;
    float* s_amp = reinterpret_cast<float*>(s_tab + tab_elems);
    uint8_t* s_valid = reinterpret_cast<uint8_t*>(s_amp + (size_t)(N + 1) * RP);

    const int x0 = blockIdx.x * TW - Nx, y0 = a.row0 + blockIdx.y * TH - Ny;
    for (int rp = tid; rp < RP; rp += nthr) {
        const int gy = y0 + rp / RW, gx = x0 + rp % RW;
        const bool inb = (gy >= 0) && (gy < a.lines) && (gx >= 0) && (gx < a.cols);
        s_valid[rp] = inb ? a.valid[(long)gy * a.cols + gx] : 0;
        s_amp[(size_t)N * RP + rp] = CUDART_INF_F;
    }
    for (int idx = tid; idx < N * RP; idx += nthr) {
        const int k = idx / RP, rp = idx - k * RP;
        const int gy = y0 + rp / RW, gx = x0 + rp % RW;
        const bool inb = (gy >= 0) && (gy < a.lines) && (gx >= 0) && (gx < a.cols);
        s_amp[idx] = inb ? __ldg(&a.amp[(long)k * npix + (long)gy * a.cols + gx]) : 0.f;
    }
    for (int i = tid; i < tab_elems; i += nthr) s_tab[i] = a.ad_table[i];
    __syncthreads();

    const int gx = blockIdx.x * TW + threadIdx.x, gy = a.row0 + blockIdx.y * TH + threadIdx.y;
    if (gx >= a.cols || gy >= a.row1) return;
    const long p = (long)gy * a.cols + gx;
    const int rp = (threadIdx.y + Ny) * RW + threadIdx.x + Nx;
    uint32_t* wp = a.wts + p * a.nulong;
    if (!s_valid[rp]) {
        a.count[p] = 0;
        for (int w = 0; w < a.nulong; ++w) wp[w] = 0u;
        return;
    }
    const double* T = (METHOD == 1) ? (a.table_in_smem ? s_tab : a.ad_table) : nullptr;
    const int WX = 2 * Nx + 1, W = WX * (2 * Ny + 1), center = Ny * WX + Nx;
    uint32_t word = 0u;
    int cnt = 0;
    int dy = -Ny, dx = -Nx;
    for (int f = 0; f < W; ++f) {
        const int rq = rp + dy * RW + dx;
        bool similar = false;
        if (s_valid[rq]) {
            if (f == center) similar = true;
            else if (METHOD == 0) {
                similar = ks_max_count_diff(s_amp + rp, s_amp + rq, N, RP) <= a.kcrit;
            } else {
                const float* first = s_amp + (f < center ? rq : rp);
                const float* second = s_amp + (f < center ? rp : rq);
                similar = ad_inner_sum(first, second, N, RP, T) <= a.scrit;
            }
        }
        if (similar) { word |= (1u << (f & 31)); ++cnt; }
        if ((f & 31) == 31 || f == W - 1) { wp[f >> 5] = word; word = 0u; }
        if (++dx > Nx) { dx = -Nx; ++dy; }
    }
    a.count[p] = cnt;
}

static size_t nmap_smem(int bands, int Nx, int Ny, int tw, int th, bool table) {
    const size_t RP = (size_t)(tw + 2 * Nx) * (th + 2 * Ny);
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
    return false;
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
    dim3 block(g.tile_w, g.tile_h);
    dim3 grid((cols + g.tile_w - 1) / g.tile_w, (nrows + g.tile_h - 1) / g.tile_h);
    cudaError_t e;
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

}  // namespace fringe
