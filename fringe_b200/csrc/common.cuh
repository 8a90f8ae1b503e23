// Shared declarations between the kernel translation units and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fringe {

// ---- nmap_kernels.cu ------------------------------------------------------------------
// amp   : float [bands][lines][nmap_amp_pitch(cols)]  ascending per valid pixel, all zero for invalid ones (rank-major:
//         the window tile of one rank is one TMA box of the 3-D tensor; rows padded to whole 16-byte units)
// valid : uint8 [npix]
// rows [row0, row0+nrows) of the block are processed (plane stride stays lines*cols)
cudaError_t launch_amp_sort(const float2* slc, const uint8_t* mask, const double* alpha, int cols,
                            int lines, int bands, float* amp, uint8_t* valid, int row0, int nrows,
                            cudaStream_t st);

// the same from amplitudes handed over pixel-major [pixel][band] with the reference's validity mask (nmapProcessBlock)
cudaError_t launch_amp_in_sort(const float* amp_in, const uint8_t* msk, int cols, int lines, int bands, float* amp,
                               uint8_t* valid, cudaStream_t st);
int nmap_amp_pitch(int cols);

struct NmapGeometry {
    int tile_w, tile_h;      // output pixels per CTA (0 x 0: the global-memory kernel, no tile fits)
    int plane_stride;        // words between the rank planes of the staged tile (one of the instantiated strides)
    size_t smem_bytes;
    bool table_in_smem;      // AD2 term table staged in shared memory
};
// Chooses the CTA tile so that tile+halo of all ranks fits in shared memory; when even the smallest
// tile does not fit, the plan selects the (slower) global-memory kernel.  Always returns true.
bool nmap_plan(int bands, int Nx, int Ny, int method, NmapGeometry* g);

cudaError_t launch_nmap(const float* amp, const uint8_t* valid, int cols, int lines, int bands,
                        int Nx, int Ny, int method, int kcrit, double scrit,
                        const double* ad_table, const NmapGeometry& g, int32_t* count,
                        uint32_t* wts, int row0, int nrows, cudaStream_t st);

// count[p] = popcount of the finished mask words, rows [row0, row0+nrows)
cudaError_t launch_count(const uint32_t* wts, int cols, int nulong, int row0, int nrows, int32_t* count,
                         cudaStream_t st);

// ---- evd_kernels.cu -------------------------------------------------------------------
// band-major planes [bands][npix] -> pixel-major vectors [npix][bands_padded] (zero padded)
// pixels [first, first+count) of the block
// zblock = 0: zpix[pix][NP] complex; zblock = B: per pixel and per block of B samples, B real parts
// then B imaginary parts
cudaError_t launch_transpose(const float2* slc, long npix, long first, long count, int bands,
                             int bands_padded, int zblock, float2* zpix, cudaStream_t st);

struct EvdArgs {
    const float2* zpix;      // [npix][NP]
    const float2* slc;       // [bands][npix] (original planes; used for the compressed SLC)
    const uint32_t* wts;     // [npix][nulong]
    int cols, lines, bands, NP;
    int Nx, Ny, nulong;
    int first_line, n_lines;
    int method, bandwidth, mini_stack_count, variant, min_neighbors;
    float2* out;             // [bands][npix]
    float* tcorr;            // [npix]
    float2* comp;            // [npix]
    unsigned long long* stats;   // [4] device counters
    int force_generic;           // profiling (fringe_prof_force_generic): bit 0 bypasses the specialised kernels
    unsigned char* scratch;      // generic kernel, large bands: per-warp workspaces in global memory (else NULL)
    int tile_pairs;              // set by the launchers of the specialised kernels: columns per CTA segment
    int zblock;                  // 0: zpix is interleaved complex [pix][NP]; -1: the FP16 hi/lo layout of evd_mma.cu
    int* worklist = nullptr;     // [2 + pixels]: [0] pixel counter of k_evd_cta, [1] number of deferred pixels, [2..] their indices
    int list_mode = 0;           // k_evd<...>: solve the pixels of the work list instead of the line range
};
int evd_max_bands(int method, int variant);
// launch geometry of the generic kernel; *use_scratch = the per-warp workspace does not fit shared
// memory and EvdArgs::scratch must provide grid * warps * evd_generic_workspace_bytes() bytes
void evd_generic_plan(const EvdArgs& a, int* warps, long* grid, size_t* smem, bool* use_scratch);
size_t evd_generic_workspace_bytes(int bands, bool dp);
cudaError_t launch_evd(const EvdArgs& a, cudaStream_t st, int* n_launches);

// ---- evd_mma.cu -----------------------------------------------------------------------
// tensor-pipe variant (evd_mma.cu): its own pixel-major layout (zblock = -1, NP = 32: 64 words per pixel, FP16 hi and
// lo parts of the band-scaled samples) and eigen order evd_mma_order(bands) (0 = not eligible, bands <= 32)
int evd_mma_order(int bands);
// scale[b] = the power of two that brings band b's typical magnitude near 2^6 (sampled over pixels [first, first+count))
cudaError_t launch_band_scale(const float2* slc, long npix, long first, long count, int bands, float* scale, cudaStream_t st);
cudaError_t launch_transpose_mma(const float2* slc, long npix, long first, long count, int bands, const float* scale,
                                 float2* zpix, cudaStream_t st);
cudaError_t launch_evd_mma(const EvdArgs& a, cudaStream_t st);
// FP64 MLE / phase_link kernel with one Hermitian row per lane in registers (mle_kernels.cu): bands <= 32;
// same pixel-major layout as the generic kernel (zblock = 0).  evd_mle_order = 0: not covered.
int evd_mle_order(int bands);
cudaError_t launch_evd_mle(const EvdArgs& a, cudaStream_t st);
// phase_link with 32 < bands <= 104 (evd_cta.cu): one CTA per pixel, coherence matrix on chip; pixels whose |C| is
// positive definite (or whose iteration stalls) are appended to EvdArgs::worklist for k_evd<...> in list mode.
// evd_cta_order = 0: not covered (other variants / sizes, or the window's SHP list does not fit shared memory)
int evd_cta_order(int bands, int Nx, int Ny, int method, int variant);
cudaError_t launch_evd_cta(const EvdArgs& a, cudaStream_t st);

// out[i] = a[i] * b[i] (complex64, double arithmetic inside), n pixels; 16-byte aligned pointers
cudaError_t launch_cmul(const float2* a, const float2* b, float2* out, long n, cudaStream_t st);
// ampdispersion: slc [bands][npix], alpha [bands] device doubles or nullptr; da, meanamp [npix]
cudaError_t launch_ampdispersion(const float2* slc, const double* alpha, long npix, int bands, float* da, float* meanamp,
                                 cudaStream_t st);
// calamp: acc [bands][2] doubles (sum of amplitudes, number of valid pixels), accumulated; mask may be nullptr
cudaError_t launch_calamp(const float2* slc, const uint8_t* mask, long npix, int bands, double* acc, cudaStream_t st);
// PS / DS integration of one pair and the PS-aware coherence raster (python/integratePS.py)
cudaError_t launch_integrate_ps(const float2* ds_i, const float2* ds_j, const float2* slc_i, const float2* slc_j, const uint8_t* ps,
                                long n, float2* out, cudaStream_t st);
cudaError_t launch_ps_coherence(const float* tcorr, const uint8_t* ps, long n, float value, float* out, cudaStream_t st);
// despeck: mode 0 one band, 1 interferogram, 2 interferogram coherence, 3 one band with the coherence flag;
// d1, d2: npix float2 of scratch each (d2 only read in mode 2); out written for lines [first_line, +n_lines)
cudaError_t launch_despeck(const float2* z1, const float2* z2, const uint32_t* wts, int cols, int lines, int Nx,
                           int Ny, int first_line, int n_lines, int mode, float2* d1, float2* d2, float2* out,
                           cudaStream_t st);

}  // namespace fringe
