// Probe: the per-pixel Gram product of k_evd_mma in the tcgen05 formulation.
//
// One pixel = W W^T with W = [hi | lo] parts of the 64 real rows (32 bands x re/im) of its S SHPs: a 64 x 64 x S product
// issued as tcgen05.mma.cta_group::1.kind::f16 (M = 64, N = 64, K = 16 per instruction), three products (hi*hi, hi*lo,
// lo*hi) accumulated in one 64-column TMEM slot, drained by four warps with tcgen05.ld.  The probe measures what the
// tensor side of that formulation sustains per SM -- operands already sitting in shared memory in the canonical K-major
// layout (all ones, so the result is 16 x #mma whatever the layout), no gather, no hand-off to an eigen solver:
//     mode 0   issue + commit only                      (tensor-pipe ceiling of the formulation)
//     mode 1   + the four drain warps read every accumulator back (tcgen05.ld 32x32b.x8, 64 columns per pixel)
// and checks the drained values.  Build and run on a B200:
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe scripts/probes/umma_gram_probe.cu && /tmp/umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: false if the phase did not complete in ~2^24 polls (the probe then reports a time-out instead of hanging)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); ++spin) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                                  // descriptor version of sm_100
    return d;                                                // layout type 0: no swizzle
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int SLOTS = 8;            // 64-column accumulators in flight (512 TMEM columns)
constexpr int KMAX = 64;            // SHPs the operand buffers hold

template <int KC>
__global__ void __launch_bounds__(160, 1) k_umma_gram(int pixels, int mode, float* out, int* status) {
    constexpr int kchunks = KC;
    __shared__ __align__(128) __half s_hi[64 * KMAX];
    __shared__ __align__(128) __half s_lo[64 * KMAX];
    __shared__ __align__(8) uint64_t full[SLOTS], empty[SLOTS];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * KMAX; i += blockDim.x) { s_hi[i] = __float2half(1.f); s_lo[i] = __float2half(1.f); }
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the operands -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // instruction descriptor: D = F32, A = B = F16, both K-major, N = 64, M = 64
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
    // canonical K-major layout without swizzle: core matrix = 8 rows x 16 bytes (128 B); the 8 row groups of a K half are
    // 128 B apart (SBO), the two K halves of one instruction 1024 B apart (LBO); 16 SHPs further = 2048 B
    bool ok = true;
    float acc = 0.f;
    int good = 0;
    if (warp == 4) {
        if (lane == 0) {
            // descriptors of the 3 x KC instructions of a pixel, built once (the operands of this probe do not move)
            uint64_t da[3 * KC], db[3 * KC];
#pragma unroll
            for (int prod = 0; prod < 3; ++prod)
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    da[prod * KC + kc] = smem_desc(smem_u32(prod == 2 ? s_lo : s_hi) + kc * 2048, 1024, 128);
                    db[prod * KC + kc] = smem_desc(smem_u32(prod == 1 ? s_lo : s_hi) + kc * 2048, 1024, 128);
                }
            for (int p = 0; p < pixels && ok; ++p) {
                const int slot = p % SLOTS;
                if (p >= SLOTS) ok = mbar_wait(&empty[slot], ((p / SLOTS) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + slot * 64;
#pragma unroll
                for (int i = 0; i < 3 * KC; ++i) umma_f16(d, da[i], db[i], idesc, i ? 1u : 0u);
                umma_commit(&full[slot]);
            }
        }
        __syncwarp();
    } else {
        for (int p = 0; p < pixels && ok; ++p) {
            const int slot = p % SLOTS;
            ok = mbar_wait(&full[slot], (p / SLOTS) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (mode == 1 && ok) {
                const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + slot * 64;
                const float want = 16.f * 3.f * (float)kchunks;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    uint32_t v[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                 : "r"(taddr + c));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float f = __uint_as_float(v[k]);
                        acc += f;
                        if (p == pixels - 1 && f == want) ++good;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
    if (!ok) atomicExch(status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    if (warp < 4) {
        atomicAdd(&out[0], acc * 1e-9f);
        if (blockIdx.x == 0) atomicAdd(status + 1, good);
    }
}

int main(int argc, char** argv) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int pixels = argc > 1 ? atoi(argv[1]) : 20000;
    float* out; int* status;
    cudaMalloc(&out, 4); cudaMalloc(&status, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kchunks = 2; kchunks <= 4; ++kchunks)
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e30f; int hst[2] = {0, 0};
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(out, 0, 4); cudaMemset(status, 0, 8);
                cudaEventRecord(e0);
                if (kchunks == 2) k_umma_gram<2><<<nsm, 160>>>(pixels, mode, out, status);
                else if (kchunks == 3) k_umma_gram<3><<<nsm, 160>>>(pixels, mode, out, status);
                else k_umma_gram<4><<<nsm, 160>>>(pixels, mode, out, status);
                cudaEventRecord(e1);
                cudaError_t e = cudaEventSynchronize(e1);
                if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
                cudaMemcpy(hst, status, 8, cudaMemcpyDeviceToHost);
                if (hst[0]) { printf("kchunks %d mode %d: mbarrier wait timed out\n", kchunks, mode); break; }
            }
            const double pxs = (double)nsm * pixels / (best * 1e-3);
            const double cyc = best * 1e-3 * 1.965e9 / pixels;
            printf("SHPs %2d  mode %d (%s): %.3f ms for %d px/SM -> %.3e px/s, %.0f SM cycles per pixel, %d mma per pixel (%.1f cycles each)%s\n",
                   16 * kchunks, mode, mode ? "issue + drain" : "issue only", best, pixels, pxs, cyc, 3 * kchunks, cyc / (3 * kchunks),
                   mode ? (hst[1] == 64 * 64 ? ", drained values correct (4096 of 4096 entries)" : ", DRAINED VALUES WRONG") : "");
            if (mode) printf("         (entries of the last pixel equal to 16 x #mma in CTA 0: %d)\n", hst[1]);
        }
    return 0;
}
