/* fringe_b200 -- C ABI of the B200-native phase-linking hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no C++ or
 * torch types.  Everything above it (the C++ block drivers nmap_process / evd_process,
 * the nmaplib / evdlib / phase_linklib Python modules, the CLIs) only moves rasters around;
 * everything below it is hand-written sm_100a CUDA.
 *
 * The one C-style kernel boundary the reference has is
 *     void nmapProcessBlock(float* amp, unsigned char* msk, int cols, int lines, int bands,
 *                           int* cnt, unsigned int* wmask, int wtslen, double pval,
 *                           int Nx, int Ny);                       (src/nmap/nmap_cuda.h:13-17)
 * plus lockGPU()/unlockGPU() (src/nmap/nmap_cuda.cu:368-376).  The entry points below keep
 * its conventions -- caller-owned host buffers, one block of image lines per call, the same
 * array layouts the reference drivers hold in memory -- and extend them to the whole path.
 * INTEGRATION.md shows the call sites in the reference that would bind to each of them.
 *
 * Array layouts (identical to what the reference block loops hold):
 *   slc    complex64 [bands][lines*cols]   one plane per date   (arma cpxdata, evd.cpp:193)
 *   mask   uint8     [lines*cols]          non-zero = use pixel (nmap.cpp:190, :323-343)
 *   count  int32     [lines*cols]                               (nmap.cpp:194)
 *   wts    uint32    [lines*cols][nulong]  BIP, bit layout of include/fringe/ulongmask.hpp:57-95
 *   out    complex64 [bands][lines*cols]   unit phasors         (arma evddata, evd.cpp:194)
 *   tcorr  float32   [lines*cols]          >=0 coherence, <0 sentinel (evd.cpp:608-725)
 *   comp   complex64 [lines*cols]          compressed SLC       (evd.cpp:755-762)
 *
 * All entry points return FRINGE_OK (0) or a positive error code; none of them calls exit().
 */
#ifndef FRINGE_B200_H
#define FRINGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRINGE_ABI_VERSION 2

/* status codes */
enum {
    FRINGE_OK = 0,
    FRINGE_ERR_METHOD = 1,       /* unknown method; same value nmap_process returns (nmap.cpp:41) */
    FRINGE_ERR_ARGUMENT = 2,     /* null pointer / non-positive size / inconsistent geometry */
    FRINGE_ERR_UNSUPPORTED = 3,  /* valid request the kernels do not cover (e.g. bands too large) */
    FRINGE_ERR_NO_DEVICE = 4,    /* no CUDA device / driver: there is no CPU fallback */
    FRINGE_ERR_CUDA = 5,         /* CUDA runtime error; text via fringe_last_error() */
    FRINGE_ERR_MEMORY = 6        /* device or pinned allocation failed */
};

/* SHP test selector (nmapOptions::method "KS2" / "AD2", src/nmap/nmap.cpp:27-42) */
enum { FRINGE_NMAP_KS2 = 0, FRINGE_NMAP_AD2 = 1 };
/* decomposition selector (evdOptions::method "EVD" / "MLE" / "STBAS", src/evd/evd.cpp:507-510) */
enum { FRINGE_EVD_EVD = 0, FRINGE_EVD_MLE = 1, FRINGE_EVD_STBAS = 2 };
/* which driver's per-pixel control flow: src/evd/evd.cpp:566-732 or
 * src/phase_link/phase_link.cpp:524-618 (MLE with EVD fallback, honours minNeighbors) */
enum { FRINGE_VARIANT_EVD = 0, FRINGE_VARIANT_PHASE_LINK = 1 };

typedef struct fringe_ctx fringe_ctx;

/* ---- context: one per GPU; owns a stream and reusable device workspaces ---------------
 * Replaces lockGPU()/unlockGPU() (nmap_cuda.cu:368-376), which pin device 0 and reset it. */
int fringe_abi_version(void);
int fringe_device_count(int* count);
int fringe_create(int device, fringe_ctx** ctx);
int fringe_destroy(fringe_ctx* ctx);
const char* fringe_last_error(const fringe_ctx* ctx);   /* never NULL */
const char* fringe_status_string(int status);
/* Block until everything queued on the context's own stream has finished. */
int fringe_synchronize(fringe_ctx* ctx);
/* Number of kernel launches issued through this context since creation. */
int64_t fringe_launch_count(const fringe_ctx* ctx);

/* pinned host memory for the block buffers (the reference uses pageable arma arrays) */
int fringe_host_alloc(void** ptr, size_t bytes);
int fringe_host_free(void* ptr);

/* ---- host-side helpers (pure integer / double arithmetic, no device needed) ----------- */
/* ceil((2Ny+1)(2Nx+1)/32), nmap.cpp:65 */
int fringe_nulong(int Nx, int Ny);
/* Largest integer k with KolmogorovProb(k/N * sqrt(N/2)) >= pvalue (KS2sample.hpp:42-77,:139-141):
 * a pair is accepted iff max_v |#{a<=v} - #{b<=v}| <= k.  *margin receives the smaller of the
 * two distances |p(k)-pvalue|, |p(k+1)-pvalue| so callers can assert the decision is not
 * rounding-sensitive. */
int fringe_ks2_critical_count(int bands, double pvalue, int* kcrit, double* margin);
/* Largest value S of the Anderson-Darling inner sum (AD2unique.hpp:287-303) for which the
 * reference's p-value chain (:309-348, :125-160) still returns >= pvalue. */
int fringe_ad2_critical_sum(int bands, double pvalue, double* scrit);
/* sigma_N of AD2unique::getSigmaN(N,N) (AD2unique.hpp:162-208) */
int fringe_ad2_sigma(int bands, double* sigma);

/* ---- SHP selection ------------------------------------------------------------------
 * One block of `lines` image lines, all columns.  Replaces nmapProcessBlock and the CPU loops
 * src/nmap/nmap.cpp:370-473: amplitude / calibration / validity, per-pixel sort, pair tests,
 * bitmask and neighbour count.  Unlike nmapProcessBlock it takes the complex samples (the
 * amplitude is computed on the device) and honours `method` (the reference's GPU path always
 * runs KS2, nmap_cuda.cu:310).
 *   mask  may be NULL (all pixels usable);  alpha may be NULL (all 1.0) else [bands] doubles,
 *   already normalised by band 0 as nmap.cpp:225-230 does.
 * Host variant: pointers are host memory (pinned preferred); copies in, runs, copies out,
 * returns when the results are in `count` / `wts`. */
int fringe_nmap_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, const double* alpha,
                      int cols, int lines, int bands, int Nx, int Ny, int method, double pvalue,
                      int32_t* count, uint32_t* wts);
/* Device variant: every pointer is device memory on the context's GPU; work is queued on
 * `stream` (a cudaStream_t, NULL = the context's stream) and the call returns immediately. */
int fringe_nmap_block_device(fringe_ctx* ctx, const float* slc, const uint8_t* mask,
                             const double* alpha, int cols, int lines, int bands, int Nx, int Ny,
                             int method, double pvalue, int32_t* count, uint32_t* wts,
                             void* stream);

/* ---- covariance + eigen solve + phase referencing + compression + temporal coherence --
 * One block of `lines` lines; results are produced for lines [first_line, first_line+n_lines)
 * only (the reference's firstlinetowrite / linestowrite, evd.cpp:414-443); the rest of the
 * output arrays is left untouched.  Replaces the pixel loops src/evd/evd.cpp:512-788
 * (variant EVD) and src/phase_link/phase_link.cpp:479-666 (variant PHASE_LINK).
 *   bandwidth        STBAS only (evd.cpp:74-91 range rules apply)
 *   mini_stack_count 1-based index of the first non-compressed band (evd.cpp:740,757)
 *   min_neighbors    used by variant PHASE_LINK only (evd.cpp hard-codes 2, :566) */
int fringe_evd_block(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols, int lines,
                     int bands, int Nx, int Ny, int first_line, int n_lines, int method,
                     int bandwidth, int mini_stack_count, int variant, int min_neighbors,
                     float* out, float* tcorr, float* comp);
int fringe_evd_block_device(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols,
                            int lines, int bands, int Nx, int Ny, int first_line, int n_lines,
                            int method, int bandwidth, int mini_stack_count, int variant,
                            int min_neighbors, float* out, float* tcorr, float* comp,
                            void* stream);

/* ---- both stages on one upload -----------------------------------------------------------
 * fringe_nmap_block followed by fringe_evd_block on the same block, with the block uploaded once
 * and the bit mask kept on the device between the stages: what nmap.py -> evd.py do through the
 * weights file (src/nmap/nmap.cpp:487-559 writes it, src/evd/evd.cpp:470-480 reads it back).
 * Results are bit-identical to the two separate calls.  `count` and `wts` cover all `lines` and
 * may each be NULL when the caller does not want them copied back; out / tcorr / comp as in
 * fringe_evd_block.  Host pointers (pinned preferred). */
int fringe_nmap_evd_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, const double* alpha,
                          int cols, int lines, int bands, int Nx, int Ny, int nmap_method,
                          double pvalue, int first_line, int n_lines, int evd_method, int bandwidth,
                          int mini_stack_count, int variant, int min_neighbors, int32_t* count,
                          uint32_t* wts, float* out, float* tcorr, float* comp);

/* ---- sequential (ministack) estimator, device resident -------------------------------------
 * What src/sequential/sequential.py:190-254 does through files -- ministack k links the compressed SLCs of
 * ministacks 1..k-1 followed by its own `mini_stack_size` acquisitions with miniStackCount = k, then a
 * "datum connection" run links all compressed SLCs -- followed by the wrapped-phase adjustment of
 * python/adjustMiniStacks.py:180-199, on one block of lines with the stack uploaded exactly once: the
 * compressed SLCs never leave the device between the stages.
 *   slc         complex64 [n_dates][lines*cols]   all acquisitions, in time order
 *   wts         uint32    [lines*cols][nulong]    SHP mask of the full stack (nmap)
 *   out_mini    complex64 [n_dates][lines*cols]   phasor of every date from its own ministack (<ministack>/EVD/<date>.slc)
 *   tcorr_mini  float32   [n_mini][lines*cols]    temporal coherence of every ministack
 *   comp        complex64 [n_mini][lines*cols]    compressed SLCs (compressedSlc/<last date>/<last date>.slc)
 *   out_datum   complex64 [n_mini][lines*cols]    datum-connection phasors (Datum_connection/EVD/<last date>.slc)
 *   tcorr_datum float32   [lines*cols]
 *   adjusted    complex64 [n_dates][lines*cols]   out_mini(date) * out_datum(its ministack); may be NULL
 * n_mini = ceil(n_dates / mini_stack_size).  Results are delivered for lines [first_line, first_line + n_lines).
 * A compressed SLC is needed Ny lines beyond the rows of every later stage, so ministack k is solved on the
 * delivered rows +- (n_mini - k + 1) * Ny (clipped to the block): results are identical to the file-based chain
 * when the block carries fringe_sequential_halo() extra lines on each side that is not an image edge; the redundant
 * rows are recomputed instead of exchanged, which is what lets row tiles run on different GPUs without any
 * communication.  `method` as in fringe_evd_block (the reference's chain runs the evd binding's default, MLE).
 * Pointers may be host memory (pinned preferred) or device memory of the context's GPU (unified addressing tells
 * them apart): with device pointers nothing crosses the host link.  Returns when the results are in place. */
int fringe_sequential_halo(int n_dates, int mini_stack_size, int Ny);
int fringe_sequential_block(fringe_ctx* ctx, const float* slc, const uint32_t* wts, int cols, int lines, int n_dates,
                            int Nx, int Ny, int first_line, int n_lines, int mini_stack_size, int method, int bandwidth,
                            float* out_mini, float* tcorr_mini, float* comp, float* out_datum, float* tcorr_datum,
                            float* adjusted);

/* ---- datum adjustment of the sequential estimator ----------------------------------------
 * out[i] = a[i] * b[i] on n complex64 pixels: the wrapped time series
 * adjusted(date) = ministack phasor(date) * datum phasor(ministack), which
 * python/adjustMiniStacks.py:180-199 delegates to GDAL's "mul" VRT pixel function (complex product
 * evaluated in double, stored as CFloat32 -- the arithmetic reproduced here).  `a`, `b`, `out` may
 * alias.  Host variant: host pointers (pinned preferred), returns when `out` is filled.  Device
 * variant: 16-byte aligned device pointers, work queued on `stream`. */
int fringe_cmul(fringe_ctx* ctx, const float* a, const float* b, int64_t n, float* out);
int fringe_cmul_device(fringe_ctx* ctx, const float* a, const float* b, int64_t n, float* out,
                       void* stream);

/* ---- amplitude dispersion -----------------------------------------------------------------
 * Mean calibrated amplitude and amplitude dispersion (sigma / mean over the dates with non-zero
 * amplitude; -1 where fewer than two dates are valid or sigma is not positive) of every pixel of
 * a block; replaces the loops src/ampdispersion/ampdispersion.cpp:207-247.  alpha: [bands]
 * calibration constants already normalised by the reference band (:119-127), NULL = all 1.
 * da, meanamp: float32 [lines*cols], the values the reference writes to its two Float32 rasters. */
int fringe_ampdispersion_block(fringe_ctx* ctx, const float* slc, const double* alpha, int cols, int lines,
                               int bands, float* da, float* meanamp);
int fringe_ampdispersion_block_device(fringe_ctx* ctx, const float* slc, const double* alpha, int cols,
                                      int lines, int bands, float* da, float* meanamp, void* stream);

/* ---- amplitude calibration (calamp) -------------------------------------------------------
 * Adds, for every band of one block, the sum of the amplitudes of its valid pixels to sums[band] and their number
 * to counts[band] (valid: amplitude neither 0 nor NaN and mask > 0; mask may be NULL); the calibration constant of
 * src/calamp/calamp.cpp:228-243 is sums / counts once all blocks are in.  The caller zeroes sums / counts. */
int fringe_calamp_block(fringe_ctx* ctx, const float* slc, const uint8_t* mask, int cols, int lines, int bands,
                        double* sums, double* counts);

/* ---- PS / DS integration (python/integratePS.py:97-130, :134-159) ----------------------------
 * out = ds_j * conj(ds_i), except where ps == 1: there exp(1j * angle(slc_j * conj(slc_i))).  n complex64 pixels each;
 * host or device pointers.  fringe_ps_coherence: out = ps == 1 ? ps_value : tcorr (the reference uses 0.95). */
int fringe_integrate_ps(fringe_ctx* ctx, const float* ds_i, const float* ds_j, const float* slc_i, const float* slc_j,
                        const uint8_t* ps, int64_t n, float* out);
int fringe_ps_coherence(fringe_ctx* ctx, const float* tcorr, const uint8_t* ps, int64_t n, float ps_value, float* out);

/* ---- despeck: SHP-weighted average ----------------------------------------------------------
 * One block of `lines` lines; replaces the preparation and pixel loops of
 * src/despeck/despeck.cpp:321-361 and :387-432.  z1, z2: the two bands ([lines*cols] complex64) the
 * reference reads with ibands[0] / ibands[1]; z2 == NULL = single band (its amplitude is averaged).
 *   z2 != NULL, compute_coherence == 0   average of z1 * conj(z2) over the SHPs
 *   z2 != NULL, compute_coherence != 0   that sum / (sqrt(sum |z1|^2) * sqrt(sum |z2|^2))
 *   z2 == NULL, compute_coherence != 0   zeros, as the reference yields (its weight sum has no imaginary part)
 * `out` [lines*cols] complex64 is written for lines [first_line, first_line + n_lines) only; pixels
 * whose own mask bit is clear get 0.  Results are bit-identical to the CPU loop (float sums in window
 * raster order). */
int fringe_despeck_block(fringe_ctx* ctx, const float* z1, const float* z2, const uint32_t* wts, int cols,
                         int lines, int Nx, int Ny, int first_line, int n_lines, int compute_coherence,
                         float* out);
int fringe_despeck_block_device(fringe_ctx* ctx, const float* z1, const float* z2, const uint32_t* wts,
                                int cols, int lines, int Nx, int Ny, int first_line, int n_lines,
                                int compute_coherence, float* out, void* stream);

/* Largest `bands` the evd kernels accept for the given method. */
int fringe_evd_max_bands(int method, int variant);

#ifdef __cplusplus
}
#endif
#endif /* FRINGE_B200_H */
