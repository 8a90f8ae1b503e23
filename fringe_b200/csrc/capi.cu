// C ABI of libfringe_b200.so (see include/fringe_b200.h): context, workspaces, host-side
// threshold arithmetic, and the host / device variants of the two block entry points.
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry
// point returns FRINGE_ERR_NO_DEVICE.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fringe_b200.h"
#include "../../include/fringe_b200_prof.h"
#include "common.cuh"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct fringe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int prof_generic = 0;                 // fringe_prof_force_generic: A/B comparisons only (bit 0; bits 8.. launch-shape overrides)
    // workspaces reused across blocks
    DevBuf amp, valid, zpix, zscale, adtab, alpha, stats, scratch, worklist;
    DevBuf in_slc, in_mask, in_wts, o_count, o_wts, o_out, o_tcorr, o_comp;
    DevBuf seq_stack[2], seq_comp, seq_mini, seq_datum;     // fringe_sequential_block
    // cached AD2 table key
    int adtab_bands = -1;
    // CUDA events bracketing the most recent launch of each kernel (on its launching stream)
    cudaEvent_t ev[FRINGE_KERNEL_COUNT][2] = {};
    bool ev_valid[FRINGE_KERNEL_COUNT] = {};
    // copy streams + event pool of the pipelined host entry points
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t pool_event(size_t i) {
        while (pool.size() <= i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); pool.push_back(e); }
        return pool[i];
    }
};

namespace {

int fail(fringe_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
int cuda_fail(fringe_ctx* c, cudaError_t e, const char* where) {
    const int code = (e == cudaErrorMemoryAllocation) ? FRINGE_ERR_MEMORY : FRINGE_ERR_CUDA;
    return fail(c, code, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);   \
    } while (0)

// ---- threshold arithmetic (host, double) ------------------------------------------------
// Kolmogorov distribution series with the constants of src/nmap/KS2sample.hpp:12-19,:42-77.
double kolmogorov_q(double z) {
    const double u = std::fabs(z);
    if (u < 0.2) return 1.0;
    if (u < 0.755) {
        const double v = 1.0 / (u * u);
        return 1 - 2.50662827 * (std::exp(-1.2337005501361697 * v) + std::exp(-11.103304951225528 * v) +
                                 std::exp(-30.842513753404244 * v)) / u;
    }
    if (u < 6.8116) {
        const double e[4] = {-2, -8, -18, -32};
        double r[4] = {0, 0, 0, 0};
        const double v = u * u;
        const int nj = std::max(1, (int)std::round(3.0 / u));
        for (int j = 0; j < nj; ++j) r[j] = std::exp(e[j] * v);
        return 2 * (r[0] - r[1] + r[2] - r[3]);
    }
    return 0.0;
}
double ks_prob_of_count(int k, int n) {
    // KS2sample.hpp:139-141 with rdmax = k/N (the reference accumulates k steps of 1/N)
    const double rn = n;
    double d = 0.0;
    const double step = 1.0 / rn;
    for (int i = 0; i < k; ++i) d += step;
    return kolmogorov_q(d * std::sqrt(rn * rn / (rn + rn)));
}

// sigma_N for two samples of equal size, src/nmap/AD2unique.hpp:162-208 (same summation order)
double ad_sigma(int n) {
    const int N = 2 * n;
    const double H = 1.0 / (1.0 * n) + 1.0 / (1.0 * n);
    double h = 0.0, g = 0.0;
    if (N < 2000) {
        std::vector<double> inv(N, 0.0);
        for (int i = 1; i < N; ++i) { inv[i] = 1.0 / i; h += inv[i]; }
        for (int i = 1; i < N - 1; ++i) {
            const double t = inv[N - i];
            for (int j = i + 1; j < N; ++j) g += t * inv[j];
        }
    } else {
        h = std::log(double(N - 1)) + 0.5772156649015328606065120900824024;
        g = (M_PI) * (M_PI) / 6.0;
    }
    const double k = 2.0, k2 = std::pow(k, 2);
    const double a = (4 * g - 6) * (k - 1) + (10 - 6 * g) * H;
    const double b = (2 * g - 4) * k2 + 8 * h * k + (2 * g - 14 * h - 4) * H - 8 * h + 4 * g - 6;
    const double c = (6 * h + 2 * g - 2) * k2 + (4 * h - 4 * g + 6) * k + (2 * h - 6) * H + 4 * h;
    const double d = (2 * h + 6) * k2 - 4 * h * k;
    double s = 0.0;
    s += a * std::pow(double(N), 3) + b * std::pow(double(N), 2) + c * N + d;
    s /= (double(N - 1) * double(N - 2) * double(N - 3));
    return std::sqrt(s);
}
// p-value of the standardised statistic, AD2unique.hpp:125-160; knots are column 0 of the
// reference's table (:9-44), probabilities :47-49.
double ad_pvalue(double tx) {
    static const double knot[35] = {
        -1.1954, -1.1786, -1.166, -1.1407, -1.1253, -1.0777, -1.0489, -0.9978, -0.9417, -0.8981,
        -0.8598, -0.7258, -0.5966, -0.4572, -0.2966, -0.1009, 0.1571, 0.5357, 1.2255, 1.5262,
        1.9633, 2.7314, 3.7825, 4.1241, 4.6044, 5.409, 6.4954, 6.8279, 7.2755, 8.1885, 9.3061,
        9.6132, 10.0989, 10.8825, 11.8537};
    static const double prob[35] = {
        .00001, .00005, .0001, .0005, .001, .005, .01, .025, .05, .075, .1, .2, .3, .4, .5, .6, .7,
        .8, .9, .925, .95, .975, .99, .9925, .995, .9975, .999, .99925, .9995, .99975, .9999,
        .999925, .99995, .999975, .99999};
    int i1 = -1;
    for (int i = 0; i < 35; ++i) { i1 = i - 1; if (tx <= knot[i]) break; }
    int i2 = i1 + 1;
    if (i1 < 0) { i1 = 0; i2 = 1; }
    if (i2 >= 35) { i1 = 33; i2 = 34; }
    const double lp1 = std::log((1.0 - prob[i1]) / prob[i1]), lp2 = std::log((1.0 - prob[i2]) / prob[i2]);
    const double lp0 = (lp1 - lp2) * (tx - knot[i2]) / (knot[i1] - knot[i2]) + lp2;
    return std::exp(lp0) / (1. + std::exp(lp0));
}
// AD2unique.hpp:305-348 for equal inner sums S (a-side == b-side, see nmap_kernels.cu)
double ad_prob_of_sum(double S, int n, double sigma) {
    double akn2 = 0.0;
    akn2 = akn2 + S / (n * 1.0);
    akn2 = akn2 + S / (n * 1.0);
    akn2 = akn2 / (double)(2 * n);
    double a2 = akn2 - 1;
    a2 /= sigma;
    return ad_pvalue(a2);
}
void ad_build_table(int n, std::vector<double>& T) {
    const int L = 2 * n;
    T.assign((size_t)(L - 1) * (n + 1), 0.0);
    for (int j = 0; j < L - 1; ++j) {
        const double bj = j + 1;
        for (int u = 0; u <= n; ++u) {
            const double tmp = (double)n * (double)u;        // |2N*m - N*(j+1)| = N*|2m-(j+1)|
            T[(size_t)j * (n + 1) + u] = tmp * tmp / (bj * ((double)L - bj));
        }
    }
}

int check_geometry(fringe_ctx* ctx, int cols, int lines, int bands, int Nx, int Ny) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (cols <= 0 || lines <= 0 || bands <= 0 || Nx < 0 || Ny < 0)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "non-positive geometry");
    return FRINGE_OK;
}

}  // namespace

extern "C" {

int fringe_abi_version(void) { return FRINGE_ABI_VERSION; }

const char* fringe_status_string(int s) {
    switch (s) {
        case FRINGE_OK: return "ok";
        case FRINGE_ERR_METHOD: return "unknown method";
        case FRINGE_ERR_ARGUMENT: return "invalid argument";
        case FRINGE_ERR_UNSUPPORTED: return "unsupported configuration";
        case FRINGE_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case FRINGE_ERR_CUDA: return "CUDA error";
        case FRINGE_ERR_MEMORY: return "out of memory";
    }
    return "unknown status";
}

int fringe_device_count(int* count) {
    if (!count) return FRINGE_ERR_ARGUMENT;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { *count = 0; cudaGetLastError(); return FRINGE_ERR_NO_DEVICE; }
    *count = n;
    return n > 0 ? FRINGE_OK : FRINGE_ERR_NO_DEVICE;
}

int fringe_create(int device, fringe_ctx** out) {
    if (!out) return FRINGE_ERR_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (fringe_device_count(&n) != FRINGE_OK || device < 0 || device >= n) return FRINGE_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return FRINGE_ERR_CUDA;
    fringe_ctx* c = new fringe_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return FRINGE_ERR_CUDA; }
    for (int k = 0; k < FRINGE_KERNEL_COUNT; ++k)
        for (int j = 0; j < 2; ++j) cudaEventCreate(&c->ev[k][j]);
    cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    *out = c;
    return FRINGE_OK;
}

int fringe_destroy(fringe_ctx* c) {
    if (!c) return FRINGE_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* all[] = {&c->amp, &c->valid, &c->zpix, &c->zscale, &c->adtab, &c->alpha, &c->stats, &c->scratch, &c->worklist, &c->in_slc, &c->in_mask,
                     &c->in_wts, &c->o_count, &c->o_wts, &c->o_out, &c->o_tcorr, &c->o_comp, &c->seq_stack[0], &c->seq_stack[1],
                     &c->seq_comp, &c->seq_mini, &c->seq_datum};
    for (DevBuf* b : all) b->release();
    for (int k = 0; k < FRINGE_KERNEL_COUNT; ++k)
        for (int j = 0; j < 2; ++j) if (c->ev[k][j]) cudaEventDestroy(c->ev[k][j]);
    for (cudaEvent_t e : c->pool) cudaEventDestroy(e);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    cudaStreamDestroy(c->stream);
    delete c;
    return FRINGE_OK;
}

const char* fringe_last_error(const fringe_ctx* c) { return c ? c->err.c_str() : "null context"; }

int fringe_synchronize(fringe_ctx* ctx) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return FRINGE_OK;
}

int64_t fringe_launch_count(const fringe_ctx* c) { return c ? c->launches : 0; }

int fringe_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return FRINGE_ERR_ARGUMENT;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return e == cudaErrorMemoryAllocation ? FRINGE_ERR_MEMORY : FRINGE_ERR_NO_DEVICE; }
    return FRINGE_OK;
}
int fringe_host_free(void* ptr) {
    if (!ptr) return FRINGE_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? FRINGE_OK : FRINGE_ERR_CUDA;
}

int fringe_nulong(int Nx, int Ny) { return ((2 * Ny + 1) * (2 * Nx + 1) + 31) / 32; }

int fringe_ks2_critical_count(int bands, double pvalue, int* kcrit, double* margin) {
    if (bands <= 0 || !kcrit) return FRINGE_ERR_ARGUMENT;
    int k = -1;
    for (int i = 0; i <= bands; ++i) {
        if (ks_prob_of_count(i, bands) >= pvalue) k = i; else break;
    }
    *kcrit = k;
    if (margin) {
        double m = DBL_MAX;
        if (k >= 0) m = std::min(m, std::fabs(ks_prob_of_count(k, bands) - pvalue));
        if (k < bands) m = std::min(m, std::fabs(ks_prob_of_count(k + 1, bands) - pvalue));
        *margin = m;
    }
    return FRINGE_OK;
}

int fringe_ad2_sigma(int bands, double* sigma) {
    if (bands < 2 || !sigma) return FRINGE_ERR_ARGUMENT;
    *sigma = ad_sigma(bands);
    return FRINGE_OK;
}

int fringe_ad2_critical_sum(int bands, double pvalue, double* scrit) {
    if (bands < 2 || !scrit) return FRINGE_ERR_ARGUMENT;
    const double sg = ad_sigma(bands);
    if (!(ad_prob_of_sum(0.0, bands, sg) >= pvalue)) { *scrit = -1.0; return FRINGE_OK; }
    if (ad_prob_of_sum(DBL_MAX, bands, sg) >= pvalue) { *scrit = DBL_MAX; return FRINGE_OK; }
    // bisection over the (ordered) bit patterns of non-negative doubles
    uint64_t lo = 0, hi;
    double dmax = DBL_MAX;
    std::memcpy(&hi, &dmax, 8);
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        double s;
        std::memcpy(&s, &mid, 8);
        if (ad_prob_of_sum(s, bands, sg) >= pvalue) lo = mid; else hi = mid;
    }
    std::memcpy(scrit, &lo, 8);
    return FRINGE_OK;
}

int fringe_evd_max_bands(int method, int variant) { return fringe::evd_max_bands(method, variant); }

// ---------------------------------------------------------------------------------------
// nmap
// ---------------------------------------------------------------------------------------
namespace {

struct NmapPlan {
    fringe::NmapGeometry g;
    int kcrit = 0;
    double scrit = 0.0;
};

// validation + thresholds + workspaces shared by the host and device variants
int nmap_prepare(fringe_ctx* ctx, int cols, int lines, int bands, int Nx, int Ny, int method, double pvalue,
                 cudaStream_t st, NmapPlan* plan) {
    int rc = check_geometry(ctx, cols, lines, bands, Nx, Ny);
    if (rc) return rc;
    if (method != FRINGE_NMAP_KS2 && method != FRINGE_NMAP_AD2) return fail(ctx, FRINGE_ERR_METHOD, "method must be KS2 or AD2");
    if (method == FRINGE_NMAP_AD2 && bands < 2) return fail(ctx, FRINGE_ERR_UNSUPPORTED, "AD2 needs >= 2 bands");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)cols * lines;
    if (!fringe::nmap_plan(bands, Nx, Ny, method, &plan->g))
        return fail(ctx, FRINGE_ERR_UNSUPPORTED, "no pair-test plan");      // not reached: the plan falls back to the global-memory kernel
    if (method == FRINGE_NMAP_KS2) {
        fringe_ks2_critical_count(bands, pvalue, &plan->kcrit, nullptr);
    } else {
        fringe_ad2_critical_sum(bands, pvalue, &plan->scrit);
        if (ctx->adtab_bands != bands) {
            std::vector<double> T;
            ad_build_table(bands, T);
            CU(ctx->adtab.ensure(T.size() * sizeof(double)));
            CU(cudaMemcpyAsync(ctx->adtab.p, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));     // T is a local
            ctx->adtab_bands = bands;
        }
    }
    CU(ctx->amp.ensure((size_t)fringe::nmap_amp_pitch(cols) * lines * bands * sizeof(float)));
    CU(ctx->valid.ensure(npix));
    return FRINGE_OK;
}

// amplitude+sort for input rows [s0, s0+sn), pair tests producing output rows [r0, r0+rn)
int nmap_launch_rows(fringe_ctx* ctx, const NmapPlan& plan, const float* slc, const uint8_t* mask,
                     const double* alpha, int cols, int lines, int bands, int Nx, int Ny, int method,
                     int32_t* count, uint32_t* wts, int s0, int sn, int r0, int rn, cudaStream_t st) {
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_AMP_SORT][0], st));
    CU(fringe::launch_amp_sort((const float2*)slc, mask, alpha, cols, lines, bands, (float*)ctx->amp.p,
                               (uint8_t*)ctx->valid.p, s0, sn, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_AMP_SORT][1], st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_NMAP][0], st));
    CU(fringe::launch_nmap((const float*)ctx->amp.p, (const uint8_t*)ctx->valid.p, cols, lines, bands, Nx, Ny,
                           method, plan.kcrit, plan.scrit, (const double*)ctx->adtab.p, plan.g, count, wts,
                           r0, rn, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_NMAP][1], st));
    // forward tests of rows < r0+rn have now delivered every bit of rows [r0, r0+rn)
    CU(fringe::launch_count(wts, cols, fringe_nulong(Nx, Ny), r0, rn, count, st));
    ctx->ev_valid[FRINGE_KERNEL_AMP_SORT] = ctx->ev_valid[FRINGE_KERNEL_NMAP] = true;
    ctx->launches += (sn > 0 ? 1 : 0) + (rn > 0 ? 2 : 0);
    return FRINGE_OK;
}

// rows per pipeline stage: ~384 MB of input per stage, at least 8 rows (20 / 40 / 80 / 160 rows of the
// 30 x 20000 bench geometry measured 308 / 293 / 287 / 293 ms per end-to-end step)
int chunk_rows(int cols, int bands, int total_rows) {
    const double row_bytes = (double)cols * bands * 8.0;
    int r = (int)(384.0e6 / row_bytes);
    if (r < 8) r = 8;
    if (r > total_rows) r = total_rows;
    return r;
}

// Row boundaries of the pipeline stages over [begin, end): full stages of `step` rows in the middle,
// a ramp of quarter and half stages at both ends, so that the first upload and the last download -- the
// two transfers nothing overlaps with -- are short.
std::vector<int> chunk_plan(int begin, int end, int step) {
    std::vector<int> cuts{begin};
    const int total = end - begin;
    if (total <= 0) return cuts;
    const int q = std::max(1, step / 4), h = std::max(1, step / 2);
    if (total < 3 * step) {                                  // too short for a ramp: plain stages
        for (int r = begin + step; r < end; r += step) cuts.push_back(r);
        cuts.push_back(end);
        return cuts;
    }
    int r = begin;
    r += q; cuts.push_back(r);
    r += h; cuts.push_back(r);
    const int tail = q + h;
    while (end - r - tail > step) { r += step; cuts.push_back(r); }
    if (end - r > tail) { r = end - tail; cuts.push_back(r); }
    r = end - q; cuts.push_back(r);
    cuts.push_back(end);
    return cuts;
}

}  // namespace

int fringe_nmap_block_device(fringe_ctx* ctx, const float* slc, const uint8_t* mask, const double* alpha,
                             int cols, int lines, int bands, int Nx, int Ny, int method, double pvalue,
                             int32_t* count, uint32_t* wts, void* stream) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !count || !wts) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    NmapPlan plan;
    int rc = nmap_prepare(ctx, cols, lines, bands, Nx, Ny, method, pvalue, st, &plan);
    if (rc) return rc;
    CU(cudaMemsetAsync(wts, 0, (size_t)cols * lines * fringe_nulong(Nx, Ny) * sizeof(uint32_t), st));
    return nmap_launch_rows(ctx, plan, slc, mask, alpha, cols, lines, bands, Nx, Ny, method, count, wts,
                            0, lines, 0, lines, st);
}

// Host variant: the block is cut into row chunks and pipelined over three streams -- upload of
// chunk k+1, kernels of chunk k and download of chunk k-1 overlap (pinned host memory needed for
// real overlap; pageable memory still works, serialised by the driver).
int fringe_nmap_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, const double* alpha, int cols,
                      int lines, int bands, int Nx, int Ny, int method, double pvalue, int32_t* count,
                      uint32_t* wts) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !count || !wts) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    cudaStream_t st = ctx->stream;
    NmapPlan plan;
    int rc = nmap_prepare(ctx, cols, lines, bands, Nx, Ny, method, pvalue, st, &plan);
    if (rc) return rc;
    const size_t npix = (size_t)cols * lines;
    const int nu = fringe_nulong(Nx, Ny);
    CU(ctx->in_slc.ensure(npix * bands * sizeof(float2)));
    CU(ctx->o_count.ensure(npix * sizeof(int32_t)));
    CU(ctx->o_wts.ensure(npix * nu * sizeof(uint32_t)));
    const uint8_t* dmask = nullptr;
    if (mask) { CU(ctx->in_mask.ensure(npix)); dmask = (const uint8_t*)ctx->in_mask.p; }
    const double* dalpha = nullptr;
    if (alpha) {
        CU(ctx->alpha.ensure(bands * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->alpha.p, alpha, bands * sizeof(double), cudaMemcpyHostToDevice, st));
        dalpha = (const double*)ctx->alpha.p;
    }
    CU(cudaMemsetAsync(ctx->o_wts.p, 0, npix * nu * sizeof(uint32_t), st));
    const int step = chunk_rows(cols, bands, lines);
    int uploaded = 0;                  // input rows [0, uploaded) are on the device and sorted
    size_t ev = 0;
    for (int r0 = 0; r0 < lines; r0 += step) {
        const int r1 = std::min(lines, r0 + step);
        const int need = std::min(lines, r1 + Ny);          // pair tests of rows < r1 read rows < r1+Ny
        const int s0 = uploaded, sn = need - uploaded;
        if (sn > 0) {
            const size_t off = (size_t)s0 * cols, cnt = (size_t)sn * cols;
            CU(cudaMemcpy2DAsync((float2*)ctx->in_slc.p + off, npix * sizeof(float2), (const float2*)slc + off,
                                 npix * sizeof(float2), cnt * sizeof(float2), bands, cudaMemcpyHostToDevice, ctx->s_in));
            if (mask) CU(cudaMemcpyAsync((uint8_t*)ctx->in_mask.p + off, mask + off, cnt, cudaMemcpyHostToDevice, ctx->s_in));
            uploaded = need;
        }
        cudaEvent_t e_in = ctx->pool_event(ev++), e_done = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_in, ctx->s_in));
        CU(cudaStreamWaitEvent(st, e_in, 0));
        rc = nmap_launch_rows(ctx, plan, (const float*)ctx->in_slc.p, dmask, dalpha, cols, lines, bands, Nx, Ny,
                              method, (int32_t*)ctx->o_count.p, (uint32_t*)ctx->o_wts.p, s0, sn, r0, r1 - r0, st);
        if (rc) return rc;
        CU(cudaEventRecord(e_done, st));
        CU(cudaStreamWaitEvent(ctx->s_out, e_done, 0));
        const size_t off = (size_t)r0 * cols, cnt = (size_t)(r1 - r0) * cols;
        CU(cudaMemcpyAsync(count + off, (int32_t*)ctx->o_count.p + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaMemcpyAsync(wts + off * nu, (uint32_t*)ctx->o_wts.p + off * nu, cnt * nu * sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Link-time drop-in for the reference's one C-style kernel boundary (src/nmap/nmap_cuda.h:13-17, called
// at src/nmap/nmap.cpp:475-485 under BUILD_NMAP_WITH_CUDA): same names, same C++ linkage, same argument
// meaning -- amplitudes pixel-major [pixel][band] as computed by nmap.cpp:370-381, the validity mask,
// host pointers in and out, synchronous, KS2 (the reference's device path ignores `method`,
// nmap_cuda.cu:310).  lockGPU / unlockGPU create / destroy the context on device 0 like
// nmap_cuda.cu:368-376.  Errors are reported on stderr and, as in cudaUtils.h:13-21, end the process.
// ---------------------------------------------------------------------------------------
namespace { fringe_ctx* g_shim_ctx = nullptr; }

void lockGPU() {
    if (g_shim_ctx) return;
    const int rc = fringe_create(0, &g_shim_ctx);
    if (rc != FRINGE_OK) { std::fprintf(stderr, "lockGPU: %s\n", fringe_status_string(rc)); std::exit(1); }
}

void unlockGPU() {
    if (g_shim_ctx) fringe_destroy(g_shim_ctx);
    g_shim_ctx = nullptr;
}

void nmapProcessBlock(float* amp, unsigned char* msk, int cols, int lines, int bands, int* cnt, unsigned int* wmask,
                      int wtslen, double pval, int Nx, int Ny) {
    if (!g_shim_ctx) lockGPU();
    fringe_ctx* ctx = g_shim_ctx;
    auto die = [&](const char* what, int rc) {
        std::fprintf(stderr, "nmapProcessBlock: %s (%s: %s)\n", what, fringe_status_string(rc), fringe_last_error(ctx));
        std::exit(1);
    };
    if (!amp || !msk || !cnt || !wmask || wtslen != fringe_nulong(Nx, Ny)) die("bad arguments", FRINGE_ERR_ARGUMENT);
    cudaStream_t st = ctx->stream;
    NmapPlan plan;
    int rc = nmap_prepare(ctx, cols, lines, bands, Nx, Ny, FRINGE_NMAP_KS2, pval, st, &plan);
    if (rc) die("prepare", rc);
    const size_t npix = (size_t)cols * lines;
    auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) { cuda_fail(ctx, e, what); die(what, FRINGE_ERR_CUDA); } };
    cu(ctx->in_slc.ensure(npix * bands * sizeof(float)), "alloc amp");
    cu(ctx->in_mask.ensure(npix), "alloc mask");
    cu(ctx->o_count.ensure(npix * sizeof(int32_t)), "alloc count");
    cu(ctx->o_wts.ensure(npix * wtslen * sizeof(uint32_t)), "alloc wts");
    cu(cudaMemcpyAsync(ctx->in_slc.p, amp, npix * bands * sizeof(float), cudaMemcpyHostToDevice, st), "upload amp");
    cu(cudaMemcpyAsync(ctx->in_mask.p, msk, npix, cudaMemcpyHostToDevice, st), "upload mask");
    cu(cudaMemsetAsync(ctx->o_wts.p, 0, npix * wtslen * sizeof(uint32_t), st), "clear wts");
    cu(fringe::launch_amp_in_sort((const float*)ctx->in_slc.p, (const uint8_t*)ctx->in_mask.p, cols, lines, bands,
                                  (float*)ctx->amp.p, (uint8_t*)ctx->valid.p, st), "sort");
    cu(fringe::launch_nmap((const float*)ctx->amp.p, (const uint8_t*)ctx->valid.p, cols, lines, bands, Nx, Ny, FRINGE_NMAP_KS2,
                           plan.kcrit, plan.scrit, (const double*)ctx->adtab.p, plan.g, (int32_t*)ctx->o_count.p,
                           (uint32_t*)ctx->o_wts.p, 0, lines, st), "pair tests");
    cu(fringe::launch_count((const uint32_t*)ctx->o_wts.p, cols, wtslen, 0, lines, (int32_t*)ctx->o_count.p, st), "count");
    ctx->launches += 3;
    cu(cudaMemcpyAsync(cnt, ctx->o_count.p, npix * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "download count");
    cu(cudaMemcpyAsync(wmask, ctx->o_wts.p, npix * wtslen * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "download wts");
    cu(cudaStreamSynchronize(st), "synchronize");
}

extern "C" {

// ---------------------------------------------------------------------------------------
// evd / phase_link
// ---------------------------------------------------------------------------------------
static int check_evd(fringe_ctx* ctx, int cols, int lines, int bands, int Nx, int Ny, int first_line,
                     int n_lines, int method, int bandwidth, int mini_stack_count, int variant) {
    int rc = check_geometry(ctx, cols, lines, bands, Nx, Ny);
    if (rc) return rc;
    if (method != FRINGE_EVD_EVD && method != FRINGE_EVD_MLE && method != FRINGE_EVD_STBAS)
        return fail(ctx, FRINGE_ERR_METHOD, "method must be EVD, MLE or STBAS");
    if (variant != FRINGE_VARIANT_EVD && variant != FRINGE_VARIANT_PHASE_LINK)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "unknown variant");
    if (first_line < 0 || n_lines < 0 || first_line + n_lines > lines)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "line range outside block");
    if (mini_stack_count < 1 || mini_stack_count > bands)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "miniStackCount outside [1, bands]");
    if (method == FRINGE_EVD_STBAS && (bandwidth <= 0 || bandwidth >= bands - 1))
        return fail(ctx, FRINGE_ERR_ARGUMENT, "STBAS bandwidth must be in (0, bands-1)");   // evd.cpp:74-91
    if (bands > fringe::evd_max_bands(method, variant))
        return fail(ctx, FRINGE_ERR_UNSUPPORTED, "too many bands for the evd kernels");
    return FRINGE_OK;
}

namespace {

struct EvdPlan { int NP = 0; int zblock = 0; bool generic = false; int prof_bits = 0; };

int evd_prepare(fringe_ctx* ctx, int cols, int lines, int bands, int method, int variant, cudaStream_t st,
                EvdPlan* plan) {
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)cols * lines;
    plan->NP = (bands + 1) & ~1;
    plan->generic = (ctx->prof_generic & 1) != 0;
    plan->prof_bits = ctx->prof_generic;
    if (!plan->generic && variant == FRINGE_VARIANT_EVD && method != FRINGE_EVD_MLE && fringe::evd_mma_order(bands) > 0) {
        plan->NP = 32;                        // 64 words per pixel: FP16 hi and lo parts of 32 bands
        plan->zblock = -1;
        CU(ctx->zscale.ensure(32 * sizeof(float)));
        CU(cudaMemsetAsync(ctx->zscale.p, 0, 32 * sizeof(float), st));        // every band's scale still open
    }
    // one extra, all-zero sample vector behind the image: the register-blocked kernel points
    // exhausted / out-of-block SHP slots at it instead of branching
    CU(ctx->zpix.ensure((npix + 1) * plan->NP * sizeof(float2)));
    CU(cudaMemsetAsync((float2*)ctx->zpix.p + npix * plan->NP, 0, plan->NP * sizeof(float2), st));
    CU(ctx->stats.ensure(16 * sizeof(unsigned long long)));      // 8 counters + 8 phase clocks
    CU(cudaMemsetAsync(ctx->stats.p, 0, 16 * sizeof(unsigned long long), st));
    return FRINGE_OK;
}

// re-layout of input rows [t0, t0+tn), then the solve for output rows [first_line, first_line+n_lines)
int evd_launch_rows(fringe_ctx* ctx, const EvdPlan& plan, const float* slc, const uint32_t* wts, int cols,
                    int lines, int bands, int Nx, int Ny, int t0, int tn, int first_line, int n_lines,
                    int method, int bandwidth, int mini_stack_count, int variant, int min_neighbors,
                    float* out, float* tcorr, float* comp, cudaStream_t st) {
    const size_t npix = (size_t)cols * lines;
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_TRANSPOSE][0], st));
    if (plan.zblock < 0) {
        if (tn > 0) {          // per-band scales: fixed by the first rows of this block call that hold data, a no-op afterwards
            CU(fringe::launch_band_scale((const float2*)slc, (long)npix, (long)t0 * cols, (long)tn * cols, bands,
                                         (float*)ctx->zscale.p, st));
            ctx->launches += 1;
        }
        CU(fringe::launch_transpose_mma((const float2*)slc, (long)npix, (long)t0 * cols, (long)tn * cols, bands,
                                        (const float*)ctx->zscale.p, (float2*)ctx->zpix.p, st));
    } else
        CU(fringe::launch_transpose((const float2*)slc, (long)npix, (long)t0 * cols, (long)tn * cols, bands, plan.NP,
                                    plan.zblock, (float2*)ctx->zpix.p, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_TRANSPOSE][1], st));
    fringe::EvdArgs a;
    a.zpix = (const float2*)ctx->zpix.p; a.slc = (const float2*)slc; a.wts = wts;
    a.cols = cols; a.lines = lines; a.bands = bands; a.NP = plan.NP;
    a.Nx = Nx; a.Ny = Ny; a.nulong = fringe_nulong(Nx, Ny);
    a.first_line = first_line; a.n_lines = n_lines;
    a.method = method; a.bandwidth = bandwidth; a.mini_stack_count = mini_stack_count;
    a.variant = variant; a.min_neighbors = min_neighbors;
    a.out = (float2*)out; a.tcorr = tcorr; a.comp = (float2*)comp;
    a.stats = (unsigned long long*)ctx->stats.p;
    a.zblock = plan.zblock; a.tile_pairs = 0; a.scratch = nullptr;
    a.force_generic = plan.prof_bits;
    if (plan.zblock >= 0) {
        int gw; long gg; size_t gs; bool use_scratch;
        fringe::evd_generic_plan(a, &gw, &gg, &gs, &use_scratch);
        if (use_scratch) {
            const bool dp = (method == FRINGE_EVD_MLE) || (variant == FRINGE_VARIANT_PHASE_LINK);
            CU(ctx->scratch.ensure((size_t)gg * gw * fringe::evd_generic_workspace_bytes(bands, dp)));
            a.scratch = (unsigned char*)ctx->scratch.p;
        }
        if (fringe::evd_cta_order(bands, Nx, Ny, method, variant) > 0) {     // CTA-per-pixel kernel: counters + deferred pixels
            CU(ctx->worklist.ensure((2 + (size_t)std::max(n_lines, 1) * cols) * sizeof(int)));
            a.worklist = (int*)ctx->worklist.p;
        }
    }
    int nl = 0;
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_EVD][0], st));
    if (n_lines > 0) CU(fringe::launch_evd(a, st, &nl));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_EVD][1], st));
    ctx->ev_valid[FRINGE_KERNEL_TRANSPOSE] = ctx->ev_valid[FRINGE_KERNEL_EVD] = true;
    ctx->launches += (tn > 0 ? 1 : 0) + nl;
    return FRINGE_OK;
}

}  // namespace

int fringe_evd_block_device(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols, int lines,
                            int bands, int Nx, int Ny, int first_line, int n_lines, int method, int bandwidth,
                            int mini_stack_count, int variant, int min_neighbors, float* out, float* tcorr,
                            float* comp, void* stream) {
    int rc = check_evd(ctx, cols, lines, bands, Nx, Ny, first_line, n_lines, method, bandwidth, mini_stack_count, variant);
    if (rc) return rc;
    if (!slc || !wts || !out || !tcorr || !comp) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (n_lines == 0) return FRINGE_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    EvdPlan plan;
    rc = evd_prepare(ctx, cols, lines, bands, method, variant, st, &plan);
    if (rc) return rc;
    // only the rows the requested lines can see need the re-layout
    const int t0 = std::max(0, first_line - Ny), t1 = std::min(lines, first_line + n_lines + Ny);
    return evd_launch_rows(ctx, plan, slc, wts, cols, lines, bands, Nx, Ny, t0, t1 - t0, first_line, n_lines, method,
                           bandwidth, mini_stack_count, variant, min_neighbors, out, tcorr, comp, st);
}

// Host variant, pipelined like fringe_nmap_block: row chunks flow through upload -> re-layout +
// solve -> download on three streams.
int fringe_evd_block(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols, int lines, int bands,
                     int Nx, int Ny, int first_line, int n_lines, int method, int bandwidth,
                     int mini_stack_count, int variant, int min_neighbors, float* out, float* tcorr,
                     float* comp) {
    int rc = check_evd(ctx, cols, lines, bands, Nx, Ny, first_line, n_lines, method, bandwidth, mini_stack_count, variant);
    if (rc) return rc;
    if (!slc || !wts || !out || !tcorr || !comp) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (n_lines == 0) return FRINGE_OK;
    cudaStream_t st = ctx->stream;
    EvdPlan plan;
    rc = evd_prepare(ctx, cols, lines, bands, method, variant, st, &plan);
    if (rc) return rc;
    const size_t npix = (size_t)cols * lines;
    const int nu = fringe_nulong(Nx, Ny);
    CU(ctx->in_slc.ensure(npix * bands * sizeof(float2)));
    CU(ctx->in_wts.ensure(npix * nu * sizeof(uint32_t)));
    CU(ctx->o_out.ensure(npix * bands * sizeof(float2)));
    CU(ctx->o_tcorr.ensure(npix * sizeof(float)));
    CU(ctx->o_comp.ensure(npix * sizeof(float2)));
    const int step = chunk_rows(cols, bands, n_lines);
    const int last = first_line + n_lines;
    int uploaded = std::max(0, first_line - Ny);      // input rows [.., uploaded) are on the device
    size_t ev = 0;
    for (int r0 = first_line; r0 < last; r0 += step) {
        const int r1 = std::min(last, r0 + step);
        const int need = std::min(lines, r1 + Ny);
        const int t0 = uploaded, tn = need - uploaded;
        if (tn > 0) {
            const size_t off = (size_t)t0 * cols, cnt = (size_t)tn * cols;
            CU(cudaMemcpy2DAsync((float2*)ctx->in_slc.p + off, npix * sizeof(float2), (const float2*)slc + off,
                                 npix * sizeof(float2), cnt * sizeof(float2), bands, cudaMemcpyHostToDevice, ctx->s_in));
            uploaded = need;
        }
        {   // the mask words of the output rows only
            const size_t off = (size_t)r0 * cols * nu, cnt = (size_t)(r1 - r0) * cols * nu;
            CU(cudaMemcpyAsync((uint32_t*)ctx->in_wts.p + off, wts + off, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->s_in));
        }
        cudaEvent_t e_in = ctx->pool_event(ev++), e_done = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_in, ctx->s_in));
        CU(cudaStreamWaitEvent(st, e_in, 0));
        rc = evd_launch_rows(ctx, plan, (const float*)ctx->in_slc.p, (const uint32_t*)ctx->in_wts.p, cols, lines, bands,
                             Nx, Ny, t0, tn, r0, r1 - r0, method, bandwidth, mini_stack_count, variant, min_neighbors,
                             (float*)ctx->o_out.p, (float*)ctx->o_tcorr.p, (float*)ctx->o_comp.p, st);
        if (rc) return rc;
        CU(cudaEventRecord(e_done, st));
        CU(cudaStreamWaitEvent(ctx->s_out, e_done, 0));
        const size_t off = (size_t)r0 * cols, cnt = (size_t)(r1 - r0) * cols;
        CU(cudaMemcpy2DAsync((float2*)out + off, npix * sizeof(float2), (float2*)ctx->o_out.p + off,
                             npix * sizeof(float2), cnt * sizeof(float2), bands, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaMemcpyAsync(tcorr + off, (float*)ctx->o_tcorr.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaMemcpyAsync((float2*)comp + off, (float2*)ctx->o_comp.p + off, cnt * sizeof(float2), cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

// Fused host variant: one upload of the block feeds both stages.  Row chunks flow through
// upload -> (amplitude sort, pair tests, count, re-layout, solve) -> download; the bit mask never
// leaves the device between the stages (it is still copied out when `wts` is given).  Same results,
// bit for bit, as fringe_nmap_block followed by fringe_evd_block on the same block.
int fringe_nmap_evd_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, const double* alpha, int cols,
                          int lines, int bands, int Nx, int Ny, int nmap_method, double pvalue, int first_line,
                          int n_lines, int evd_method, int bandwidth, int mini_stack_count, int variant,
                          int min_neighbors, int32_t* count, uint32_t* wts, float* out, float* tcorr,
                          float* comp) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    int rc = check_evd(ctx, cols, lines, bands, Nx, Ny, first_line, n_lines, evd_method, bandwidth, mini_stack_count, variant);
    if (rc) return rc;
    if (!slc || !out || !tcorr || !comp) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    cudaStream_t st = ctx->stream;
    NmapPlan nplan;
    rc = nmap_prepare(ctx, cols, lines, bands, Nx, Ny, nmap_method, pvalue, st, &nplan);
    if (rc) return rc;
    EvdPlan eplan;
    rc = evd_prepare(ctx, cols, lines, bands, evd_method, variant, st, &eplan);
    if (rc) return rc;
    const size_t npix = (size_t)cols * lines;
    const int nu = fringe_nulong(Nx, Ny);
    CU(ctx->in_slc.ensure(npix * bands * sizeof(float2)));
    CU(ctx->o_count.ensure(npix * sizeof(int32_t)));
    CU(ctx->o_wts.ensure(npix * nu * sizeof(uint32_t)));
    CU(ctx->o_out.ensure(npix * bands * sizeof(float2)));
    CU(ctx->o_tcorr.ensure(npix * sizeof(float)));
    CU(ctx->o_comp.ensure(npix * sizeof(float2)));
    const uint8_t* dmask = nullptr;
    if (mask) { CU(ctx->in_mask.ensure(npix)); dmask = (const uint8_t*)ctx->in_mask.p; }
    const double* dalpha = nullptr;
    if (alpha) {
        CU(ctx->alpha.ensure(bands * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->alpha.p, alpha, bands * sizeof(double), cudaMemcpyHostToDevice, st));
        dalpha = (const double*)ctx->alpha.p;
    }
    CU(cudaMemsetAsync(ctx->o_wts.p, 0, npix * nu * sizeof(uint32_t), st));
    const std::vector<int> cuts = chunk_plan(0, lines, chunk_rows(cols, bands, lines));
    const int last = first_line + n_lines;
    int uploaded = 0;
    size_t ev = 0;
    for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
        const int r0 = cuts[ci], r1 = cuts[ci + 1];
        const int need = std::min(lines, r1 + Ny);
        const int s0 = uploaded, sn = need - uploaded;
        if (sn > 0) {
            const size_t off = (size_t)s0 * cols, cnt = (size_t)sn * cols;
            CU(cudaMemcpy2DAsync((float2*)ctx->in_slc.p + off, npix * sizeof(float2), (const float2*)slc + off,
                                 npix * sizeof(float2), cnt * sizeof(float2), bands, cudaMemcpyHostToDevice, ctx->s_in));
            if (mask) CU(cudaMemcpyAsync((uint8_t*)ctx->in_mask.p + off, mask + off, cnt, cudaMemcpyHostToDevice, ctx->s_in));
            uploaded = need;
        }
        cudaEvent_t e_in = ctx->pool_event(ev++), e_done = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_in, ctx->s_in));
        CU(cudaStreamWaitEvent(st, e_in, 0));
        rc = nmap_launch_rows(ctx, nplan, (const float*)ctx->in_slc.p, dmask, dalpha, cols, lines, bands, Nx, Ny,
                              nmap_method, (int32_t*)ctx->o_count.p, (uint32_t*)ctx->o_wts.p, s0, sn, r0, r1 - r0, st);
        if (rc) return rc;
        // solve the part of this chunk that lies inside the requested lines; its mask rows are final
        const int e0 = std::max(r0, first_line), e1 = std::min(r1, last);
        rc = evd_launch_rows(ctx, eplan, (const float*)ctx->in_slc.p, (const uint32_t*)ctx->o_wts.p, cols, lines, bands,
                             Nx, Ny, s0, sn, e0, std::max(0, e1 - e0), evd_method, bandwidth, mini_stack_count, variant,
                             min_neighbors, (float*)ctx->o_out.p, (float*)ctx->o_tcorr.p, (float*)ctx->o_comp.p, st);
        if (rc) return rc;
        CU(cudaEventRecord(e_done, st));
        CU(cudaStreamWaitEvent(ctx->s_out, e_done, 0));
        const size_t off = (size_t)r0 * cols, cnt = (size_t)(r1 - r0) * cols;
        if (count) CU(cudaMemcpyAsync(count + off, (int32_t*)ctx->o_count.p + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out));
        if (wts) CU(cudaMemcpyAsync(wts + off * nu, (uint32_t*)ctx->o_wts.p + off * nu, cnt * nu * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, ctx->s_out));
        if (e1 > e0) {
            const size_t eoff = (size_t)e0 * cols, ecnt = (size_t)(e1 - e0) * cols;
            CU(cudaMemcpy2DAsync((float2*)out + eoff, npix * sizeof(float2), (float2*)ctx->o_out.p + eoff,
                                 npix * sizeof(float2), ecnt * sizeof(float2), bands, cudaMemcpyDeviceToHost, ctx->s_out));
            CU(cudaMemcpyAsync(tcorr + eoff, (float*)ctx->o_tcorr.p + eoff, ecnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_out));
            CU(cudaMemcpyAsync((float2*)comp + eoff, (float2*)ctx->o_comp.p + eoff, ecnt * sizeof(float2), cudaMemcpyDeviceToHost, ctx->s_out));
        }
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}


// ---------------------------------------------------------------------------------------
// sequential estimator on the device (src/sequential/sequential.py:190-254 + python/adjustMiniStacks.py)
// ---------------------------------------------------------------------------------------
int fringe_sequential_halo(int n_dates, int mini_stack_size, int Ny) {
    if (n_dates <= 0 || mini_stack_size <= 0 || Ny < 0) return 0;
    const int nmini = (n_dates + mini_stack_size - 1) / mini_stack_size;
    return (nmini + 1) * Ny;
}

int fringe_sequential_block(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols, int lines, int n_dates,
                            int Nx, int Ny, int first_line, int n_lines, int mini_stack_size, int method, int bandwidth,
                            float* out_mini, float* tcorr_mini, float* comp, float* out_datum, float* tcorr_datum,
                            float* adjusted) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !wts || !out_mini || !tcorr_mini || !comp || !out_datum || !tcorr_datum)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (mini_stack_size < 2 || n_dates < 2) return fail(ctx, FRINGE_ERR_ARGUMENT, "need >= 2 dates per ministack");
    const int s = mini_stack_size;
    const int nmini = (n_dates + s - 1) / s;
    if (n_dates - (nmini - 1) * s < 2 && nmini > 1)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "the last ministack would hold a single acquisition");
    const int max_bands = (nmini - 1) + s;                       // compressed SLCs + own acquisitions of the last ministack
    int rc = check_evd(ctx, cols, lines, max_bands, Nx, Ny, first_line, n_lines, method, bandwidth, 1, FRINGE_VARIANT_EVD);
    if (rc) return rc;
    if (nmini < 2) return fail(ctx, FRINGE_ERR_ARGUMENT, "fewer than two ministacks: nothing to connect (run fringe_evd_block)");
    if (nmini > fringe::evd_max_bands(method, FRINGE_VARIANT_EVD)) return fail(ctx, FRINGE_ERR_UNSUPPORTED, "too many ministacks for the datum connection");
    if (n_lines == 0) return FRINGE_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t npix = (size_t)cols * lines;
    const size_t nout = (size_t)cols * n_lines;                  // pixels of the delivered rows
    const size_t ooff = (size_t)first_line * cols;
    const int nu = fringe_nulong(Nx, Ny);
    const int last = first_line + n_lines;

    CU(ctx->in_wts.ensure(npix * nu * sizeof(uint32_t)));
    for (int b = 0; b < 2; ++b) CU(ctx->seq_stack[b].ensure(npix * max_bands * sizeof(float2)));
    CU(ctx->seq_comp.ensure(npix * nmini * sizeof(float2)));
    CU(ctx->seq_mini.ensure(nout * n_dates * sizeof(float2)));
    CU(ctx->seq_datum.ensure(npix * nmini * sizeof(float2)));
    CU(ctx->o_out.ensure(npix * max_bands * sizeof(float2)));
    CU(ctx->o_tcorr.ensure(npix * sizeof(float)));
    CU(ctx->o_comp.ensure(npix * nmini * sizeof(float)));        // temporal coherence of every ministack
    CU(cudaMemcpyAsync(ctx->in_wts.p, wts, npix * nu * sizeof(uint32_t), cudaMemcpyDefault, ctx->s_in));
    CU(cudaMemsetAsync(ctx->seq_comp.p, 0, npix * nmini * sizeof(float2), st));
    size_t ev = 0;
    cudaEvent_t buf_free[2] = {nullptr, nullptr};                // solve that last read a stack buffer has finished
    float2* d_out = (float2*)ctx->o_out.p;

    for (int k = 1; k <= nmini; ++k) {
        const int d0 = (k - 1) * s, d1 = std::min(n_dates, k * s);
        const int nk = (k - 1) + (d1 - d0);
        float2* buf = (float2*)ctx->seq_stack[k & 1].p;
        // rows on which this ministack's compressed SLC is needed: the delivered rows plus one window halo per later
        // stage (each later ministack and the datum connection look Ny lines further), clipped to the block
        const int h = (nmini - k + 1) * Ny;
        const int r0 = std::max(0, first_line - h), r1 = std::min(lines, last + h);
        const int t0 = std::max(0, r0 - Ny), t1 = std::min(lines, r1 + Ny);
        // own acquisitions -> planes [k-1, nk) of the buffer (only the rows the solve reads), overlapping the previous solve
        if (buf_free[k & 1]) CU(cudaStreamWaitEvent(ctx->s_in, buf_free[k & 1], 0));
        {
            const size_t off = (size_t)t0 * cols, cnt = (size_t)(t1 - t0) * cols;
            CU(cudaMemcpy2DAsync(buf + (size_t)(k - 1) * npix + off, npix * sizeof(float2), (const float2*)slc + (size_t)d0 * npix + off,
                                 npix * sizeof(float2), cnt * sizeof(float2), d1 - d0, cudaMemcpyDefault, ctx->s_in));
        }
        cudaEvent_t e_in = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_in, ctx->s_in));
        CU(cudaStreamWaitEvent(st, e_in, 0));
        // compressed SLCs of the earlier ministacks -> planes [0, k-1)
        if (k > 1) CU(cudaMemcpyAsync(buf, ctx->seq_comp.p, (size_t)(k - 1) * npix * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        EvdPlan plan;
        rc = evd_prepare(ctx, cols, lines, nk, method, FRINGE_VARIANT_EVD, st, &plan);
        if (rc) return rc;
        rc = evd_launch_rows(ctx, plan, (const float*)buf, (const uint32_t*)ctx->in_wts.p, cols, lines, nk, Nx, Ny, t0, t1 - t0, r0,
                             r1 - r0, method, bandwidth, k, FRINGE_VARIANT_EVD, 2, (float*)d_out, (float*)ctx->o_comp.p + (size_t)(k - 1) * npix,
                             (float*)((float2*)ctx->seq_comp.p + (size_t)(k - 1) * npix), st);
        if (rc) return rc;
        // keep the phasors of the own acquisitions (delivered rows only) for the adjustment; results to the host
        CU(cudaMemcpy2DAsync((float2*)ctx->seq_mini.p + (size_t)d0 * nout, nout * sizeof(float2), d_out + (size_t)(k - 1) * npix + ooff,
                             npix * sizeof(float2), nout * sizeof(float2), d1 - d0, cudaMemcpyDeviceToDevice, st));
        cudaEvent_t e_done = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_done, st));
        buf_free[k & 1] = e_done;
        CU(cudaStreamWaitEvent(ctx->s_out, e_done, 0));
        CU(cudaMemcpy2DAsync((float2*)out_mini + (size_t)d0 * npix + ooff, npix * sizeof(float2), (float2*)ctx->seq_mini.p + (size_t)d0 * nout,
                             nout * sizeof(float2), nout * sizeof(float2), d1 - d0, cudaMemcpyDefault, ctx->s_out));
        CU(cudaMemcpyAsync(tcorr_mini + (size_t)(k - 1) * npix + ooff, (float*)ctx->o_comp.p + (size_t)(k - 1) * npix + ooff,
                           nout * sizeof(float), cudaMemcpyDefault, ctx->s_out));
        CU(cudaMemcpyAsync((float2*)comp + (size_t)(k - 1) * npix + ooff, (float2*)ctx->seq_comp.p + (size_t)(k - 1) * npix + ooff,
                           nout * sizeof(float2), cudaMemcpyDefault, ctx->s_out));
    }
    // datum connection: all compressed SLCs, first band is the reference (sequential.py:247-254)
    {
        EvdPlan plan;
        rc = evd_prepare(ctx, cols, lines, nmini, method, FRINGE_VARIANT_EVD, st, &plan);
        if (rc) return rc;
        const int t0 = std::max(0, first_line - Ny), t1 = std::min(lines, last + Ny);
        rc = evd_launch_rows(ctx, plan, (const float*)ctx->seq_comp.p, (const uint32_t*)ctx->in_wts.p, cols, lines, nmini, Nx, Ny, t0,
                             t1 - t0, first_line, n_lines, method, bandwidth, 1, FRINGE_VARIANT_EVD, 2, (float*)ctx->seq_datum.p,
                             (float*)ctx->o_tcorr.p, (float*)d_out, st);      // the datum run's own compressed SLC is not used
        if (rc) return rc;
        CU(cudaMemcpy2DAsync((float2*)out_datum + ooff, npix * sizeof(float2), (float2*)ctx->seq_datum.p + ooff, npix * sizeof(float2),
                             nout * sizeof(float2), nmini, cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(tcorr_datum + ooff, (float*)ctx->o_tcorr.p + ooff, nout * sizeof(float), cudaMemcpyDefault, st));
    }
    // wrapped time series: ministack phasor x datum phasor of its ministack (adjustMiniStacks.py:180-199)
    if (adjusted) {
        for (int k = 1; k <= nmini; ++k) {
            const int d0 = (k - 1) * s, d1 = std::min(n_dates, k * s);
            for (int d = d0; d < d1; ++d) {
                float2* m = (float2*)ctx->seq_mini.p + (size_t)d * nout;
                // the datum plane has the block's row pitch; the kept ministack phasors are compact: multiply row-compact copies
                CU(fringe::launch_cmul(m, (const float2*)ctx->seq_datum.p + (size_t)(k - 1) * npix + ooff, m, (long)nout, st));
                ctx->launches += 1;
            }
        }
        CU(cudaMemcpy2DAsync((float2*)adjusted + ooff, npix * sizeof(float2), ctx->seq_mini.p, nout * sizeof(float2), nout * sizeof(float2),
                             n_dates, cudaMemcpyDefault, st));
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

// ---------------------------------------------------------------------------------------
// datum adjustment (python/adjustMiniStacks.py:180-199): out = a * b, complex64
// ---------------------------------------------------------------------------------------
int fringe_cmul_device(fringe_ctx* ctx, const float* a, const float* b, int64_t n, float* out, void* stream) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer or negative size");
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) != 0) return fail(ctx, FRINGE_ERR_ARGUMENT, "pointers must be 16-byte aligned");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_CMUL][0], st));
    CU(fringe::launch_cmul((const float2*)a, (const float2*)b, (float2*)out, (long)n, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_CMUL][1], st));
    ctx->ev_valid[FRINGE_KERNEL_CMUL] = true;
    ctx->launches += (n > 0) ? 1 : 0;
    return FRINGE_OK;
}

// Host variant: pieces of ~128 MB per operand flow through upload -> multiply -> download.
int fringe_cmul(fringe_ctx* ctx, const float* a, const float* b, int64_t n, float* out) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer or negative size");
    if (n == 0) return FRINGE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t piece = (size_t)16 << 20;                       // pixels per piece (even)
    const size_t nbuf = (std::min((size_t)n, 2 * piece) + 1) & ~(size_t)1;    // even: keeps the b half 16-byte aligned
    CU(ctx->in_slc.ensure(2 * nbuf * sizeof(float2)));            // a | b, double buffered halves
    CU(ctx->o_out.ensure(nbuf * sizeof(float2)));
    cudaStream_t st = ctx->stream;
    size_t ev = 0;
    int slot = 0;
    std::vector<cudaEvent_t> freed(2, nullptr);
    for (size_t off = 0; off < (size_t)n; off += piece, slot ^= 1) {
        const size_t cnt = std::min(piece, (size_t)n - off);
        float2* da = (float2*)ctx->in_slc.p + (size_t)slot * piece;
        float2* db = (float2*)ctx->in_slc.p + nbuf + (size_t)slot * piece;
        float2* dout = (float2*)ctx->o_out.p + (size_t)slot * piece;
        if (freed[slot]) CU(cudaStreamWaitEvent(ctx->s_in, freed[slot], 0));     // slot's previous download done
        CU(cudaMemcpyAsync(da, (const float2*)a + off, cnt * sizeof(float2), cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaMemcpyAsync(db, (const float2*)b + off, cnt * sizeof(float2), cudaMemcpyHostToDevice, ctx->s_in));
        cudaEvent_t e_in = ctx->pool_event(ev++), e_done = ctx->pool_event(ev++), e_out = ctx->pool_event(ev++);
        CU(cudaEventRecord(e_in, ctx->s_in));
        CU(cudaStreamWaitEvent(st, e_in, 0));
        int rc = fringe_cmul_device(ctx, (const float*)da, (const float*)db, (int64_t)cnt, (float*)dout, st);
        if (rc) return rc;
        CU(cudaEventRecord(e_done, st));
        CU(cudaStreamWaitEvent(ctx->s_out, e_done, 0));
        CU(cudaMemcpyAsync((float2*)out + off, dout, cnt * sizeof(float2), cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaEventRecord(e_out, ctx->s_out));
        freed[slot] = e_out;
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

// ---------------------------------------------------------------------------------------
// despeck (src/despeck/despeck.cpp:321-361, 387-432)
// ---------------------------------------------------------------------------------------
static int check_despeck(fringe_ctx* ctx, const void* z1, const void* wts, const void* out, int cols, int lines,
                         int Nx, int Ny, int first_line, int n_lines) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!z1 || !wts || !out) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (cols <= 0 || lines <= 0 || Nx < 0 || Ny < 0) return fail(ctx, FRINGE_ERR_ARGUMENT, "non-positive size");
    if (first_line < 0 || n_lines < 0 || first_line + n_lines > lines)
        return fail(ctx, FRINGE_ERR_ARGUMENT, "line range outside block");
    return FRINGE_OK;
}

int fringe_despeck_block_device(fringe_ctx* ctx, const float* z1, const float* z2, const uint32_t* wts, int cols,
                                int lines, int Nx, int Ny, int first_line, int n_lines, int compute_coherence,
                                float* out, void* stream) {
    int rc = check_despeck(ctx, z1, wts, out, cols, lines, Nx, Ny, first_line, n_lines);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t npix = (size_t)cols * lines;
    const int mode = z2 ? (compute_coherence ? 2 : 1) : (compute_coherence ? 3 : 0);
    CU(ctx->scratch.ensure(2 * npix * sizeof(float2)));
    float2* d1 = (float2*)ctx->scratch.p;
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_DESPECK][0], st));
    CU(fringe::launch_despeck((const float2*)z1, (const float2*)z2, wts, cols, lines, Nx, Ny, first_line, n_lines, mode,
                              d1, d1 + npix, (float2*)out, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_DESPECK][1], st));
    ctx->ev_valid[FRINGE_KERNEL_DESPECK] = true;
    ctx->launches += (n_lines > 0) ? 2 : 0;
    return FRINGE_OK;
}

int fringe_despeck_block(fringe_ctx* ctx, const float* z1, const float* z2, const uint32_t* wts, int cols, int lines,
                         int Nx, int Ny, int first_line, int n_lines, int compute_coherence, float* out) {
    int rc = check_despeck(ctx, z1, wts, out, cols, lines, Nx, Ny, first_line, n_lines);
    if (rc) return rc;
    if (n_lines == 0) return FRINGE_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t npix = (size_t)cols * lines;
    const int nu = fringe_nulong(Nx, Ny);
    CU(ctx->in_slc.ensure(2 * npix * sizeof(float2)));
    CU(ctx->in_wts.ensure(npix * nu * sizeof(uint32_t)));
    CU(ctx->o_out.ensure(npix * sizeof(float2)));
    float2* dz1 = (float2*)ctx->in_slc.p;
    float2* dz2 = z2 ? dz1 + npix : nullptr;
    CU(cudaMemcpyAsync(dz1, z1, npix * sizeof(float2), cudaMemcpyHostToDevice, st));
    if (z2) CU(cudaMemcpyAsync(dz2, z2, npix * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->in_wts.p, wts, npix * nu * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    rc = fringe_despeck_block_device(ctx, (const float*)dz1, (const float*)dz2, (const uint32_t*)ctx->in_wts.p, cols, lines,
                                     Nx, Ny, first_line, n_lines, compute_coherence, (float*)ctx->o_out.p, st);
    if (rc) return rc;
    const size_t off = (size_t)first_line * cols, cnt = (size_t)n_lines * cols;
    CU(cudaMemcpyAsync((float2*)out + off, (float2*)ctx->o_out.p + off, cnt * sizeof(float2), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

// ---------------------------------------------------------------------------------------
// ampdispersion (src/ampdispersion/ampdispersion.cpp:207-247)
// ---------------------------------------------------------------------------------------
int fringe_ampdispersion_block_device(fringe_ctx* ctx, const float* slc, const double* alpha, int cols, int lines,
                                      int bands, float* da, float* meanamp, void* stream) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !da || !meanamp) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (cols <= 0 || lines <= 0 || bands <= 0) return fail(ctx, FRINGE_ERR_ARGUMENT, "non-positive size");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_AMPDISP][0], st));
    CU(fringe::launch_ampdispersion((const float2*)slc, alpha, (long)cols * lines, bands, da, meanamp, st));
    CU(cudaEventRecord(ctx->ev[FRINGE_KERNEL_AMPDISP][1], st));
    ctx->ev_valid[FRINGE_KERNEL_AMPDISP] = true;
    ctx->launches += 1;
    return FRINGE_OK;
}

int fringe_ampdispersion_block(fringe_ctx* ctx, const float* slc, const double* alpha, int cols, int lines, int bands,
                               float* da, float* meanamp) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !da || !meanamp) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (cols <= 0 || lines <= 0 || bands <= 0) return fail(ctx, FRINGE_ERR_ARGUMENT, "non-positive size");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t npix = (size_t)cols * lines;
    CU(ctx->in_slc.ensure(npix * bands * sizeof(float2)));
    CU(ctx->o_tcorr.ensure(2 * npix * sizeof(float)));
    const double* dalpha = nullptr;
    if (alpha) {
        CU(ctx->alpha.ensure(bands * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->alpha.p, alpha, bands * sizeof(double), cudaMemcpyHostToDevice, st));
        dalpha = (const double*)ctx->alpha.p;
    }
    CU(cudaMemcpyAsync(ctx->in_slc.p, slc, npix * bands * sizeof(float2), cudaMemcpyHostToDevice, st));
    float* dda = (float*)ctx->o_tcorr.p;
    int rc = fringe_ampdispersion_block_device(ctx, (const float*)ctx->in_slc.p, dalpha, cols, lines, bands, dda, dda + npix, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(da, dda, npix * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(meanamp, dda + npix, npix * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

// ---------------------------------------------------------------------------------------
// calamp (src/calamp/calamp.cpp:207-226) and PS / DS integration (python/integratePS.py:97-159)
// ---------------------------------------------------------------------------------------
int fringe_calamp_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, int cols, int lines, int bands, double* sums,
                        double* counts) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (!slc || !sums || !counts) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer");
    if (cols <= 0 || lines <= 0 || bands <= 0) return fail(ctx, FRINGE_ERR_ARGUMENT, "non-positive size");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t npix = (size_t)cols * lines;
    CU(ctx->in_slc.ensure(npix * bands * sizeof(float2)));
    CU(ctx->scratch.ensure((size_t)bands * 2 * sizeof(double)));
    const uint8_t* dmask = nullptr;
    if (mask) { CU(ctx->in_mask.ensure(npix)); CU(cudaMemcpyAsync(ctx->in_mask.p, mask, npix, cudaMemcpyDefault, st)); dmask = (const uint8_t*)ctx->in_mask.p; }
    CU(cudaMemcpyAsync(ctx->in_slc.p, slc, npix * bands * sizeof(float2), cudaMemcpyDefault, st));
    CU(cudaMemsetAsync(ctx->scratch.p, 0, (size_t)bands * 2 * sizeof(double), st));
    CU(fringe::launch_calamp((const float2*)ctx->in_slc.p, dmask, (long)npix, bands, (double*)ctx->scratch.p, st));
    ctx->launches += 1;
    std::vector<double> h((size_t)bands * 2);
    CU(cudaMemcpyAsync(h.data(), ctx->scratch.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int b = 0; b < bands; ++b) { sums[b] += h[2 * b]; counts[b] += h[2 * b + 1]; }
    return FRINGE_OK;
}

int fringe_integrate_ps(fringe_ctx* ctx, const float* ds_i, const float* ds_j, const float* slc_i, const float* slc_j,
                        const uint8_t* ps, int64_t n, float* out) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (n < 0 || (n > 0 && (!ds_i || !ds_j || !slc_i || !slc_j || !ps || !out))) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer or negative size");
    if (n == 0) return FRINGE_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(ctx->in_slc.ensure((size_t)n * 4 * sizeof(float2)));
    CU(ctx->in_mask.ensure((size_t)n));
    CU(ctx->o_out.ensure((size_t)n * sizeof(float2)));
    float2* d = (float2*)ctx->in_slc.p;
    const float* src[4] = {ds_i, ds_j, slc_i, slc_j};
    for (int k = 0; k < 4; ++k) CU(cudaMemcpyAsync(d + (size_t)k * n, src[k], (size_t)n * sizeof(float2), cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(ctx->in_mask.p, ps, (size_t)n, cudaMemcpyDefault, st));
    CU(fringe::launch_integrate_ps(d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, (const uint8_t*)ctx->in_mask.p, (long)n, (float2*)ctx->o_out.p, st));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(out, ctx->o_out.p, (size_t)n * sizeof(float2), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

int fringe_ps_coherence(fringe_ctx* ctx, const float* tcorr, const uint8_t* ps, int64_t n, float ps_value, float* out) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    if (n < 0 || (n > 0 && (!tcorr || !ps || !out))) return fail(ctx, FRINGE_ERR_ARGUMENT, "null pointer or negative size");
    if (n == 0) return FRINGE_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(ctx->o_tcorr.ensure((size_t)n * 2 * sizeof(float)));
    CU(ctx->in_mask.ensure((size_t)n));
    float* d = (float*)ctx->o_tcorr.p;
    CU(cudaMemcpyAsync(d, tcorr, (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(ctx->in_mask.p, ps, (size_t)n, cudaMemcpyDefault, st));
    CU(fringe::launch_ps_coherence(d, (const uint8_t*)ctx->in_mask.p, (long)n, ps_value, d + n, st));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(out, d + n, (size_t)n * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    return FRINGE_OK;
}

int fringe_last_kernel_ms(fringe_ctx* ctx, int kernel, float* ms) {
    if (!ctx || !ms || kernel < 0 || kernel >= FRINGE_KERNEL_COUNT) return FRINGE_ERR_ARGUMENT;
    if (!ctx->ev_valid[kernel]) return fail(ctx, FRINGE_ERR_ARGUMENT, "kernel has not been launched on this context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev[kernel][1]));
    CU(cudaEventElapsedTime(ms, ctx->ev[kernel][0], ctx->ev[kernel][1]));
    return FRINGE_OK;
}

int fringe_evd_phase_cycles(fringe_ctx* ctx, int64_t cycles[8]) {
    if (!ctx || !cycles) return FRINGE_ERR_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[16] = {0};
    if (ctx->stats.p) CU(cudaMemcpy(h, ctx->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) cycles[i] = (int64_t)h[8 + i];
    return FRINGE_OK;
}

int fringe_evd_stats(fringe_ctx* ctx, int64_t stats[8]) {
    if (!ctx || !stats) return FRINGE_ERR_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    unsigned long long h[8] = {0};
    if (ctx->stats.p) CU(cudaMemcpy(h, ctx->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) stats[i] = (int64_t)h[i];
    return FRINGE_OK;
}

int fringe_prof_force_generic(fringe_ctx* ctx, int on) {
    if (!ctx) return FRINGE_ERR_ARGUMENT;
    ctx->prof_generic = on;
    return FRINGE_OK;
}

}  // extern "C"
