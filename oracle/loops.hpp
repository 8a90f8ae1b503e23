// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FRInGE phase-linking hot path.
//
// Pixel loops of the reference drivers restated on plain arrays (no GDAL, no
// Armadillo).  The per-pair / per-matrix workers come from a traits class `I`:
//   * oracle/restated.hpp            -> liboracle.so          (kind "port")
//   * the reference's own headers    -> _ref/libfringe_ref.so (kind "reference")
// so the same loop text is compiled against both and the two are compared bit-for-bit
// in tests/test_oracle_pins.py.
//
// Deviations from the reference, all deliberate:
//   * the nmap pair loop is race-free: every pixel tests its whole window and sets only
//     its own bits (the reference sets both ends of a pair from different OpenMP threads,
//     src/nmap/nmap.cpp:464-468).  The decision for a pair is symmetric, so the answer is
//     the race-free one the reference intends.
//   * STBAS temporal coherence clamps the band limit to nbands (the reference reads past
//     the matrix when ti+BW+1 > nbands, src/evd/evd.cpp:777-779).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

typedef std::complex<float> cfloat;
typedef std::complex<double> cdouble;

enum { NMAP_KS2 = 0, NMAP_AD2 = 1 };
enum { EVD_EVD = 0, EVD_MLE = 1, EVD_STBAS = 2 };
enum { VARIANT_EVD = 0, VARIANT_PHASE_LINK = 1 };

// ---------------------------------------------------------------------------------
// Amplitude + validity + per-pixel sort.  src/nmap/nmap.cpp:370-381 and :389-397.
//   slc   [bands][lines*cols] complex64 (one plane per date, as RasterIO delivers it)
//   mask  [lines*cols] or NULL (non-zero = use pixel)
//   alpha [bands] amplitude calibration (already normalised by band 0) or NULL (=1)
//   amp   [lines*cols][bands] float, ascending per valid pixel
//   valid [lines*cols]
// ---------------------------------------------------------------------------------
inline void amplitude_sort(const cfloat* slc, const uint8_t* mask, const double* alpha, int cols,
                           int lines, int bands, float* amp, uint8_t* valid) {
    const long npix = (long)cols * lines;
    for (long p = 0; p < npix; ++p) valid[p] = mask ? (mask[p] != 0) : 1;
    std::memset(amp, 0, sizeof(float) * npix * bands);
    for (int b = 0; b < bands; ++b) {
        const double al = alpha ? alpha[b] : 1.0;
#pragma omp parallel for
        for (long p = 0; p < npix; ++p) {
            float val;
            val = std::abs(slc[(long)b * npix + p]) / al;
            valid[p] = (valid[p] != 0) && (val != 0.) && (!std::isnan(val));
            if (valid[p]) amp[p * bands + b] = val;
        }
    }
#pragma omp parallel for
    for (long p = 0; p < npix; ++p) {
        if (!valid[p]) continue;
        std::sort(amp + p * bands, amp + (p + 1) * bands);
    }
}

// ---------------------------------------------------------------------------------
// SHP selection over one block of lines.  src/nmap/nmap.cpp:404-473.
//   count [lines*cols] int32, wts [lines*cols][nulong] uint32 (BIP, as written to disk)
// Pair (p,q): the vector of the pixel that comes first in raster order is passed as the
// first argument of the test, as in the reference (cenpix = pp, refpix = qq, qq > pp).
// ---------------------------------------------------------------------------------
template <class I>
int nmap_block(const cfloat* slc, const uint8_t* mask, const double* alpha, int cols, int lines,
               int bands, int Nx, int Ny, int method, double thresh, int32_t* count,
               uint32_t* wts, float* amp_out) {
    if (method != NMAP_KS2 && method != NMAP_AD2) return 1;
    const long npix = (long)cols * lines;
    const int nulong = (int)std::ceil(((2 * Ny + 1) * (2 * Nx + 1)) / 32.0);
    std::vector<float> amp_store;
    float* amp = amp_out;
    if (!amp) { amp_store.resize((size_t)npix * bands); amp = amp_store.data(); }
    std::vector<uint8_t> valid(npix);
    amplitude_sort(slc, mask, alpha, cols, lines, bands, amp, valid.data());
    std::memset(count, 0, sizeof(int32_t) * npix);
    std::memset(wts, 0, sizeof(uint32_t) * npix * nulong);
    const typename I::Mask bitmask = {Ny, Nx};

#pragma omp parallel
    {
        typename I::KS ks(bands);
        typename I::AD* ad = (method == NMAP_AD2) ? new typename I::AD(bands) : nullptr;
#pragma omp for schedule(dynamic, 256)
        for (long p = 0; p < npix; ++p) {
            if (!valid[p]) continue;
            const int r = (int)(p / cols), c = (int)(p % cols);
            for (int dy = -Ny; dy <= Ny; ++dy) {
                const int rr = r + dy;
                if (rr < 0 || rr >= lines) continue;
                for (int dx = -Nx; dx <= Nx; ++dx) {
                    const int cc = c + dx;
                    if (cc < 0 || cc >= cols) continue;
                    const long q = (long)rr * cols + cc;
                    if (!valid[q]) continue;
                    bool similar = true;
                    if (q != p) {
                        const float* first = amp + std::min(p, q) * bands;
                        const float* second = amp + std::max(p, q) * bands;
                        const double prob = (method == NMAP_KS2) ? ks.test(first, second)
                                                                 : ad->test(first, second);
                        similar = (prob >= thresh);
                    }
                    if (similar) {
                        count[p] += 1;
                        bitmask.setbit(wts + p * nulong, dy, dx, true);
                    }
                }
            }
        }
        delete ad;
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// Covariance + eigen solve + phase referencing + compression + temporal coherence for
// the lines [first_line, first_line+n_lines) of one block.
//   variant VARIANT_EVD        : src/evd/evd.cpp:512-788
//   variant VARIANT_PHASE_LINK : src/phase_link/phase_link.cpp:479-666
//   slc  [bands][lines*cols] complex64 ; wts [lines*cols][nulong]
//   out  [bands][lines*cols] complex64 ; tcorr [lines*cols] ; comp [lines*cols]
// All three outputs are zeroed first; pixels that are skipped stay zero and failed
// pixels carry the reference's negative sentinel in tcorr.
// ---------------------------------------------------------------------------------
template <class I>
int evd_block(const cfloat* slc, const uint32_t* wts, int cols, int lines, int bands, int Nx,
              int Ny, int first_line, int n_lines, int method, int bandwidth,
              int mini_stack_count, int variant, int min_neighbors, cfloat* out, float* tcorr,
              cfloat* comp, int32_t* npix_out) {
    const long npix_block = (long)cols * lines;
    const int nulong = (int)std::ceil(((2 * Ny + 1) * (2 * Nx + 1)) / 32.0);
    const int N = bands;
    const typename I::Mask bitmask = {Ny, Nx};
    const bool isstbas = (method == EVD_STBAS);
    const bool ismle = (method == EVD_MLE);
    const int BW = bandwidth;
    const int k0 = mini_stack_count - 1;

    std::memset((void*)out, 0, sizeof(cfloat) * npix_block * N);
    std::memset((void*)tcorr, 0, sizeof(float) * npix_block);
    std::memset((void*)comp, 0, sizeof(cfloat) * npix_block);
    if (npix_out) std::memset(npix_out, 0, sizeof(int32_t) * npix_block);

#pragma omp parallel
    {
        typename I::Eig worker;
        worker.prepare(N);
        // column-major N x N, element (r,c) at r + N*c, like an Armadillo cube slice
        std::vector<cdouble> Numer((size_t)N * N), Covar((size_t)N * N), Wk1((size_t)N * N),
            Wk2((size_t)N * N);
        std::vector<double> Pow(N);
        std::vector<cfloat> z(N);
#define AT(M, r, c) M[(size_t)(r) + (size_t)N * (c)]

#pragma omp for schedule(dynamic, 64)
        for (long pp = (long)first_line * cols; pp < (long)(first_line + n_lines) * cols; ++pp) {
            uint32_t* cen = const_cast<uint32_t*>(wts + pp * nulong);
            if (!bitmask.getbit(cen, 0, 0)) continue;

            std::fill(Numer.begin(), Numer.end(), cdouble(0, 0));
            std::fill(Pow.begin(), Pow.end(), 0.0);
            const int ci = (int)(pp / cols), cj = (int)(pp % cols);
            const int xmin = std::max(cj - Nx, 0), xmax = std::min(cols - 1, cj + Nx);
            const int ymin = std::max(ci - Ny, 0), ymax = std::min(lines - 1, ci + Ny);

            int npix = 0;
            for (int ii = ymin; ii <= ymax; ++ii)
                for (int jj = xmin; jj <= xmax; ++jj) {
                    if (!bitmask.getbit(cen, ii - ci, jj - cj)) continue;
                    ++npix;
                    const long q = (long)ii * cols + jj;
                    for (int t = 0; t < N; ++t) z[t] = slc[(long)t * npix_block + q];
                    // float product, double accumulate (evd.cpp:557); the two power sums of
                    // the reference (Amp1/Amp2, :558-559) depend on one index only.
                    for (int t = 0; t < N; ++t) {
                        const double a = std::abs(z[t]);
                        Pow[t] += a * a;
                    }
                    for (int ti = 0; ti < N; ++ti)
                        for (int tj = ti + 1; tj < N; ++tj) AT(Numer, ti, tj) += z[ti] * std::conj(z[tj]);
                }
            if (npix_out) npix_out[pp] = npix;
            if (variant == VARIANT_EVD) { if (npix < 2) continue; }
            else { if (npix < min_neighbors) continue; }

            for (int ti = 0; ti < N; ++ti) {
                for (int tj = ti + 1; tj < N; ++tj) {
                    const cdouble res = AT(Numer, ti, tj) / std::sqrt(Pow[ti] * Pow[tj]);
                    AT(Covar, ti, tj) = res;
                    AT(Covar, tj, ti) = std::conj(res);
                }
                AT(Covar, ti, ti) = cdouble(1.0, 0.0);
            }
            auto fill_abs = [&](std::vector<cdouble>& M) {
                for (int ti = 0; ti < N; ++ti) {
                    for (int tj = ti + 1; tj < N; ++tj) {
                        const double coh = std::abs(AT(Covar, ti, tj));
                        AT(M, ti, tj) = cdouble(coh, 0.0);
                        AT(M, tj, ti) = cdouble(coh, 0.0);
                    }
                    AT(M, ti, ti) = cdouble(1.0, 0.0);
                }
            };
            auto hadamard = [&](std::vector<cdouble>& M) {
                for (size_t e = 0; e < (size_t)N * N; ++e) M[e] *= Covar[e];
            };

            if (variant == VARIANT_EVD) {
                if (ismle) {                       // evd.cpp:584-688
                    fill_abs(Wk1);
                    Wk2 = Covar;
                    if (worker.smallestEigen(Wk2.data(), false) != 0) { tcorr[pp] = -1; continue; }
                    if (worker.eigval[0] < 1.0e-6) { tcorr[pp] = -2; continue; }
                    Wk2 = Wk1;
                    if (worker.smallestEigen(Wk2.data(), false) != 0) { tcorr[pp] = -3; continue; }
                    if (worker.eigval[0] < 1.0e-6) { tcorr[pp] = -4; continue; }
                    Wk2 = Wk1;
                    if (worker.positiveDefiniteInverse(Wk2.data()) != 0) { tcorr[pp] = -5; continue; }
                    hadamard(Wk2);
                    if (worker.smallestEigen(Wk2.data(), true) != 0) { tcorr[pp] = -6; continue; }
                    if (worker.eigval[0] < 1.0e-6) { tcorr[pp] = -7; continue; }
                } else {                           // evd.cpp:689-732
                    if (isstbas)
                        for (int ti = 0; ti < N; ++ti)
                            for (int tj = ti + BW + 1; tj < N; ++tj) {
                                if (tj < 0) continue;
                                AT(Covar, tj, ti) = cdouble(0., 0.);
                                AT(Covar, ti, tj) = cdouble(0., 0.);
                            }
                    Wk1 = Covar;
                    if (worker.largestEigen(Wk1.data(), true) != 0) { tcorr[pp] = -6; continue; }
                    if (worker.eigval[0] < 1.0e-6) { tcorr[pp] = -7; continue; }
                }
            } else {                               // phase_link.cpp:540-618
                Wk1 = Covar;
                if (worker.smallestEigen(Wk1.data(), false) != 0) { tcorr[pp] = -1; continue; }
                fill_abs(Wk1);
                bool run_evd = false;
                Wk2 = Wk1;
                if (worker.positiveDefiniteInverse(Wk2.data()) != 0) run_evd = true;
                if (!run_evd) {
                    hadamard(Wk2);
                    if (worker.smallestEigen(Wk2.data(), true) != 0) run_evd = true;
                }
                if (run_evd) {
                    Wk1 = Covar;
                    if (worker.largestEigen(Wk1.data(), true) != 0) { tcorr[pp] = -6; continue; }
                }
            }

            // phase reference to band k0 (evd.cpp:738-749)
            {
                const cdouble cJ(0.0, 1.0);
                const double ph0 = std::arg(worker.eigvec[k0]);
                for (int t = 0; t < N; ++t) {
                    const double res = std::arg(worker.eigvec[t]);
                    out[(long)t * npix_block + pp] = std::exp(cJ * (res - ph0));
                }
                out[(long)k0 * npix_block + pp] = cfloat(1.0f, 0.0f);
            }
            // compressed SLC over the non-compressed bands (evd.cpp:755-762)
            {
                cdouble acc = 0.0;
                for (int t = k0; t < N; ++t)
                    acc += slc[(long)t * npix_block + pp] * std::conj(out[(long)t * npix_block + pp]);
                comp[pp] = acc / (1.0 * (N - mini_stack_count + 1));
            }
            // temporal coherence (evd.cpp:770-786)
            {
                cdouble acc(0.0, 0.0);
                const cdouble cJ(0.0, 1.0);
                int counter = 0;
                for (int ti = 0; ti < N; ++ti) {
                    const int ulim = isstbas ? std::min(ti + BW + 1, N) : N;
                    for (int tj = ti + 1; tj < ulim; ++tj) {
                        acc += std::exp(cJ * (std::arg(AT(Covar, ti, tj)) -
                                              std::arg(out[(long)ti * npix_block + pp]) +
                                              std::arg(out[(long)tj * npix_block + pp])));
                        ++counter;
                    }
                }
                tcorr[pp] = std::abs(acc) / (counter * 1.0);
            }
        }
#undef AT
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// despeck: SHP-weighted average of one band (amplitude) or of an interferogram, optionally
// normalised to a coherence.  src/despeck/despeck.cpp:300-362 (per-block preparation) and :387-432
// (the pixel loop); plain arrays instead of Armadillo columns, no GDAL.  z1 / z2: the two bands of
// the block ([lines*cols] complex64), z2 == nullptr = single-band amplitude mode (ibands[1] = -1).
// Sums are complex<float> accumulated in window raster order, exactly as the reference does.
// ---------------------------------------------------------------------------------------
template <class I>
int despeck_block(const cfloat* z1, const cfloat* z2, const uint32_t* wts, int cols, int lines, int Nx, int Ny,
                  int first_line, int n_lines, int compute_coherence, cfloat* out) {
    const long npix_block = (long)cols * lines;
    const int nulong = (int)std::ceil(((2 * Ny + 1) * (2 * Nx + 1)) / 32.0);
    const typename I::Mask bitmask = {Ny, Nx};
    std::vector<cfloat> d1(npix_block), d2(npix_block);
    for (long jj = 0; jj < npix_block; ++jj) {                     // despeck.cpp:321-361
        if (z2) {
            cfloat a = z1[jj];
            const cfloat b = z2[jj];
            if (compute_coherence) {
                const float amp1 = std::abs(a), amp2 = std::abs(b);
                a *= std::conj(b);
                d1[jj] = a;
                d2[jj] = cfloat(amp1 * amp1, amp2 * amp2);
            } else {
                a *= std::conj(b);
                d1[jj] = a;
                d2[jj] = 1.0f;
            }
        } else {
            d1[jj] = std::abs(z1[jj]);
            d2[jj] = cfloat(1.0f, 0.0f);                           // arma ones(): 1 + 0i (despeck.cpp:360)
        }
    }
    std::memset((void*)out, 0, sizeof(cfloat) * npix_block);
#pragma omp parallel for schedule(dynamic, 256)
    for (long pp = (long)first_line * cols; pp < (long)(first_line + n_lines) * cols; ++pp) {   // despeck.cpp:387-432
        uint32_t* cen = const_cast<uint32_t*>(wts + pp * nulong);
        if (bitmask.getbit(cen, 0, 0) == 0) continue;
        const int ci = (int)(pp / cols), cj = (int)(pp % cols);
        const int xmin = std::max(cj - Nx, 0), xmax = std::min(cols - 1, cj + Nx);
        const int ymin = std::max(ci - Ny, 0), ymax = std::min(lines - 1, ci + Ny);
        cfloat val = 0.0f, sumw = 0.0f;
        for (int ii = ymin; ii <= ymax; ++ii)
            for (int jj = xmin; jj <= xmax; ++jj)
                if (bitmask.getbit(cen, ii - ci, jj - cj) != 0) {
                    val += d1[(long)ii * cols + jj];
                    sumw += d2[(long)ii * cols + jj];
                }
        if (sumw.real() > 0) {
            if (compute_coherence) {
                if (sumw.imag() > 0) out[pp] = val / (std::sqrt(sumw.real()) * std::sqrt(sumw.imag()));
            } else {
                out[pp] = val / sumw.real();
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// ampdispersion: per-pixel mean calibrated amplitude and amplitude dispersion over the stack.
// src/ampdispersion/ampdispersion.cpp:207-247 (sums band by band in double, then the statistics);
// alpha = calibration constants already normalised by the reference band (:119-127).  Outputs are
// the doubles the reference hands to GDAL for its Float32 rasters, converted here.
// ---------------------------------------------------------------------------------------
inline int ampdispersion_block(const cfloat* slc, const double* alpha, int cols, int lines, int bands, float* da,
                               float* meanamp) {
    const long n = (long)cols * lines;
    std::vector<double> mean(n, 0.0), meansq(n, 0.0), norms(n, 0.0);
    for (int bb = 0; bb < bands; ++bb) {
        const double al = alpha ? alpha[bb] : 1.0;
        for (long ii = 0; ii < n; ++ii) {
            volatile double absval = std::abs(slc[(long)bb * n + ii]);
            const int valid = (absval != 0.0);
            absval *= (valid / al);
            volatile double sq = absval * absval;          // volatile: no contraction into the sums
            mean[ii] += absval;
            meansq[ii] += sq;
            norms[ii] += valid;
        }
    }
    for (long ii = 0; ii < n; ++ii) {
        if (norms[ii] > 1) {
            const double avg = mean[ii] / norms[ii];
            const double avg2 = meansq[ii] / norms[ii];
            volatile double a2 = avg * avg;
            const double sdev = ::sqrt(avg2 - a2);
            meanamp[ii] = (float)avg;
            da[ii] = (float)(((!std::isnan(sdev)) && (sdev > 0)) ? (sdev / avg) : -1);
        } else {
            meanamp[ii] = 0.0f;
            da[ii] = -1.0f;
        }
    }
    return 0;
}

}  // namespace oracle
