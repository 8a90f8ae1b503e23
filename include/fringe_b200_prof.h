/* fringe_b200 -- profiling and measurement hooks.  NOT part of the drop-in boundary (that is
 * include/fringe_b200.h): bench.py, scripts/ and the tests use these to time kernels, read solver
 * statistics and measure the roofline denominators.
 *
 * Two groups:
 *   (a) context-bound hooks, implemented in libfringe_b200.so because they read the context's event
 *       pairs and device counters;
 *   (b) stand-alone microbenchmarks in their own library, libfringe_b200_prof.so (fringe_prof_*).
 */
#ifndef FRINGE_B200_PROF_H
#define FRINGE_B200_PROF_H

#include "fringe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (a) context-bound -------------------------------------------------------------------
 * Device time of the most recent launch of one kernel, from CUDA events recorded on the
 * stream it was launched on (synchronises on the closing event). */
enum {
    FRINGE_KERNEL_AMP_SORT = 0,   /* amplitude + per-pixel sort */
    FRINGE_KERNEL_NMAP = 1,       /* window pair tests */
    FRINGE_KERNEL_TRANSPOSE = 2,  /* band-major -> pixel-major re-layout */
    FRINGE_KERNEL_EVD = 3,        /* covariance + eigen + post-processing */
    FRINGE_KERNEL_CMUL = 4,       /* datum adjustment product */
    FRINGE_KERNEL_DESPECK = 5,    /* despeck preparation + SHP-weighted average */
    FRINGE_KERNEL_AMPDISP = 6,    /* amplitude dispersion */
    FRINGE_KERNEL_COUNT = 7
};
int fringe_last_kernel_ms(fringe_ctx* ctx, int kernel, float* ms);

/* Per-pixel solver statistics of the most recent evd call on this context:
 * stats[0] pixels solved, [1] FP32 power iterations (EVD) or inverse-iteration solves (MLE),
 * [2] pixels that took the FP64 path, [3] pixels that hit an iteration cap or took the certified
 * fall-back, [4] Cholesky factorisations of the MLE eigen solver and its gates, [5] pixels of the tensor-pipe EVD
 * kernel whose Gram product was recomputed on FP32 FMAs (a sample outside the FP16 range), [6..7] spare.
 * Synchronises the device. */
int fringe_evd_stats(fringe_ctx* ctx, int64_t stats[8]);
/* Per-phase warp cycles of the most recent tensor-pipe evd launch (summed over warps):
 * [0] SHP lists, [1] covariance accumulation, [2] normalisation + hand-off to shared memory,
 * [3] row load + start vector, [4] power iteration, [5] epilogue, [6..7] spare.  All zero unless
 * the library was built with -DFRINGE_PHASE_CLOCKS (python -m fringe_b200.build --phase-clocks);
 * a profiling build, not for timing. */
int fringe_evd_phase_cycles(fringe_ctx* ctx, int64_t cycles[8]);
/* A/B comparisons only: route MLE / phase_link and EVD calls on this context to the generic
 * any-N kernel (k_evd) instead of the specialised ones.  Off by default; nothing in the product
 * sets it. */
int fringe_prof_force_generic(fringe_ctx* ctx, int on);

/* ---- (b) stand-alone microbenchmarks (libfringe_b200_prof.so) -----------------------------
 * FP32 FMA throughput of the device measured with a register-resident FMA loop; the roofline
 * denominator for the covariance + eigen kernel (MEASURED_PEAKS.json carries no FP32 figure). */
int fringe_prof_fp32_peak(int device, double* tflops);
/* FP32 rate of a register-resident 6x6 complex block update (the covariance inner step with loads
 * and address arithmetic removed): [0] interleaved and [1] de-interleaved scalar FFMA, [2] packed
 * fma.rn.f32x2. */
int fringe_prof_block_fma_rate(int device, double tflops[3]);
/* Dense TF32 TFLOP/s of the warp-level mma.sync.m16n8k8 path (12 independent accumulator tiles per
 * warp). */
int fringe_prof_mma_tf32_rate(int device, double* tflops);
/* The same on mma.sync.m16n8k16 FP16 with FP32 accumulate (the two-term FP16 split of the Gram product). */
int fringe_prof_mma_f16_rate(int device, double* tflops);
/* ... and on mma.sync.m16n8k8 FP16 (half the k extent per instruction; the Gram kernel's shape). */
int fringe_prof_mma_f16_k8_rate(int device, double* tflops);
/* FP64 FMA throughput (register-resident DFMA loop): denominator for the MLE kernel. */
int fringe_prof_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* FRINGE_B200_PROF_H */
