// Stand-alone probe of the TMA box load used by k_nmap: one variant per process (a faulting kernel poisons the context).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/probes/tma_probe.cu && for v in 0 1 2 ...; do /tmp/tma_probe $v; done
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DIMS>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, int use_gmap, int x0, int y0, int z0,
                        int box_words, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_raw);
    const CUtensorMap* m = use_gmap ? gmap : &map;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_words * 4) : "memory");
        if (DIMS == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(dst)), "l"(m), "r"(x0), "r"(y0), "r"(z0), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(dst)), "l"(m), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
    }
    long spins = 0;
    uint32_t done = 0;
    while (!done && spins < 2000000) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ++spins;
    }
    for (int i = threadIdx.x; i < box_words; i += blockDim.x) out[i] = done ? dst[i] : 0xDEADBEEFu;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    const int cols = 96, lines = 40, bands = 20;
    int pitch = 96;
    int bw = 44, bh = 10, x0 = -5, y0 = 16, z0 = 3, dims = 3;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT32;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    int use_gmap = 0, use_cluster = 0;
    switch (v) {
        case 0: break;
        case 1: x0 = 0; break;
        case 2: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
        case 3: dims = 2; break;
        case 4: bw = 64; break;
        case 5: l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
        case 6: use_cluster = 1; break;
        case 7: use_gmap = 1; break;
        case 8: bw = 32; x0 = 0; y0 = 0; z0 = 0; break;
        case 9: y0 = 0; break;
        case 10: x0 = -4; break;
        case 11: x0 = 3; break;
        case 12: x0 = -8; y0 = -2; break;
        case 13: x0 = 60; y0 = 36; break;      // box runs over the right and bottom edges
    }
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("variant %d: no entry point\n", v); return 1;
    }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    std::vector<uint32_t> h((size_t)pitch * lines * bands);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i + 1u;
    uint32_t *d = nullptr, *out = nullptr;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int box_words = bw * bh;
    cudaMalloc(&out, box_words * 4);
    CUtensorMap map;
    CUresult r;
    const cuuint32_t estr[3] = {1, 1, 1};
    if (dims == 3) {
        const cuuint64_t gd[3] = {(cuuint64_t)cols, (cuuint64_t)lines, (cuuint64_t)bands};
        const cuuint64_t gs[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * lines * 4};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
        r = enc(&map, dt, 3, d, gd, gs, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)lines * bands};
        const cuuint64_t gs[1] = {(cuuint64_t)pitch * 4};
        const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
        r = enc(&map, dt, 2, d, gd, gs, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        y0 = z0 * lines + y0;
    }
    if (r != CUDA_SUCCESS) { printf("variant %d: encode failed %d\n", v, (int)r); return 1; }
    CUtensorMap* gmap = nullptr;
    cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)box_words * 4 + 128;
    if (use_cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_probe<3>, map, (const CUtensorMap*)gmap, use_gmap, x0, y0, z0, box_words, out);
    } else if (dims == 3) k_probe<3><<<1, 128, smem>>>(map, gmap, use_gmap, x0, y0, z0, box_words, out);
    else k_probe<2><<<1, 128, smem>>>(map, gmap, use_gmap, x0, y0, z0, box_words, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: kernel error: %s\n", v, cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> o(box_words);
    cudaMemcpy(o.data(), out, box_words * 4, cudaMemcpyDeviceToHost);
    int bad = 0, timeout = (o[0] == 0xDEADBEEFu);
    for (int yy = 0; yy < bh && !timeout; ++yy)
        for (int xx = 0; xx < bw; ++xx) {
            const int gx = x0 + xx, gy = (dims == 3 ? y0 : y0 - z0 * lines) + yy;
            uint32_t want = 0;
            if (gx >= 0 && gx < cols && gy >= 0 && gy < lines) want = (uint32_t)(((size_t)z0 * lines + gy) * pitch + gx) + 1u;
            if (dims == 2 && gx >= 0 && gx < cols) { const long row = (long)y0 + yy; want = (row >= 0 && row < (long)lines * bands) ? (uint32_t)(row * pitch + gx) + 1u : 0; }
            if (o[yy * bw + xx] != want) ++bad;
        }
    printf("variant %d: %s (mismatches %d of %d)\n", v, timeout ? "mbarrier never completed" : (bad ? "WRONG DATA" : "ok"), bad, box_words);
    return 0;
}
