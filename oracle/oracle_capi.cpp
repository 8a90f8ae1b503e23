// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (ctypes-loadable).
//
// Compiled twice by oracle/Makefile:
//   liboracle.so             : workers from oracle/restated.hpp               ("port")
//   _ref/libfringe_ref.so    : -DORACLE_USE_REFERENCE_HEADERS, workers are the reference's
//                              own KS2sample.hpp / AD2unique.hpp / ulongmask.hpp /
//                              EigenLapack.hpp included from /root/reference  ("reference")
// Nothing in fringe_b200/ links or loads either library.
#include <cstdio>

#ifdef ORACLE_USE_REFERENCE_HEADERS
#include <complex>
#include "KS2sample.hpp"
#include "AD2unique.hpp"
#include "fringe/ulongmask.hpp"
#include "fringe/EigenLapack.hpp"
struct Impl {
    typedef KS2sample KS;
    typedef AD2unique AD;
    typedef Ulongmask Mask;
    typedef EVWorker Eig;
};
static const char* kKind = "reference";
static double kprob(double z) { KS2sample t(2); return t.KolmogorovProb(z); }
static double ad_sigma(int n) { AD2unique t(n); return t.sigmaNorm; }
static double ad_pvalue(double a2, int n) { AD2unique t(n); return t.PValueADKSamples(a2); }
#else
#include "restated.hpp"
struct Impl {
    typedef restated::KSTest KS;
    typedef restated::ADTest AD;
    typedef restated::WindowMask Mask;
    typedef restated::EigSolver Eig;
};
static const char* kKind = "port";
static double kprob(double z) { return restated::KSTest::kolmogorov_prob(z); }
static double ad_sigma(int n) { restated::ADTest t(n); return t.sigma; }
static double ad_pvalue(double a2, int n) { restated::ADTest t(n); return t.pvalue(a2); }
#endif

#include "loops.hpp"

extern "C" {

const char* oracle_kind() { return kKind; }
int oracle_max_threads() { return omp_get_max_threads(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }

double oracle_ks2_prob(const float* a, const float* b, int n) {
    Impl::KS t(n);
    return t.test(a, b);
}
double oracle_kolmogorov_prob(double z) { return kprob(z); }
double oracle_ad2_prob(const float* a, const float* b, int n) {
    Impl::AD t(n);
    return t.test(a, b);
}
double oracle_ad2_sigma(int n) { return ad_sigma(n); }
double oracle_ad2_pvalue_of_stat(double a2, int n) { return ad_pvalue(a2, n); }
void oracle_mask_setbit(uint32_t* words, int Ny, int Nx, int dy, int dx, int on) {
    const Impl::Mask m = {Ny, Nx};
    m.setbit(words, dy, dx, on != 0);
}
int oracle_mask_getbit(uint32_t* words, int Ny, int Nx, int dy, int dx) {
    const Impl::Mask m = {Ny, Nx};
    return m.getbit(words, dy, dx) ? 1 : 0;
}
// which: 0 = smallest, 1 = largest.  A is column-major n x n complex128, destroyed.
int oracle_eig_extreme(double* A, int n, int which, int want_vec, double* eigval, double* eigvec) {
    Impl::Eig w;
    w.prepare(n);
    std::complex<double>* Ac = reinterpret_cast<std::complex<double>*>(A);
    const int info = which ? w.largestEigen(Ac, want_vec != 0) : w.smallestEigen(Ac, want_vec != 0);
    eigval[0] = w.eigval[0];
    if (want_vec)
        for (int i = 0; i < n; ++i) { eigvec[2 * i] = w.eigvec[i].real(); eigvec[2 * i + 1] = w.eigvec[i].imag(); }
    return info;
}
int oracle_pd_inverse(double* A, int n) {
    Impl::Eig w;
    w.prepare(n);
    return w.positiveDefiniteInverse(reinterpret_cast<std::complex<double>*>(A));
}

int oracle_nmap_block(const float* slc, const uint8_t* mask, const double* alpha, int cols, int lines,
                      int bands, int Nx, int Ny, int method, double thresh, int32_t* count,
                      uint32_t* wts, float* amp_sorted) {
    return oracle::nmap_block<Impl>(reinterpret_cast<const oracle::cfloat*>(slc), mask, alpha, cols,
                                    lines, bands, Nx, Ny, method, thresh, count, wts, amp_sorted);
}

int oracle_evd_block(const float* slc, const uint32_t* wts, int cols, int lines, int bands, int Nx,
                     int Ny, int first_line, int n_lines, int method, int bandwidth,
                     int mini_stack_count, int variant, int min_neighbors, float* out, float* tcorr,
                     float* comp, int32_t* npix) {
    return oracle::evd_block<Impl>(reinterpret_cast<const oracle::cfloat*>(slc), wts, cols, lines,
                                   bands, Nx, Ny, first_line, n_lines, method, bandwidth,
                                   mini_stack_count, variant, min_neighbors,
                                   reinterpret_cast<oracle::cfloat*>(out), tcorr,
                                   reinterpret_cast<oracle::cfloat*>(comp), npix);
}

int oracle_ampdispersion_block(const float* slc, const double* alpha, int cols, int lines, int bands, float* da,
                               float* meanamp) {
    return oracle::ampdispersion_block((const oracle::cfloat*)slc, alpha, cols, lines, bands, da, meanamp);
}

int oracle_despeck_block(const float* z1, const float* z2, const uint32_t* wts, int cols, int lines, int Nx, int Ny,
                         int first_line, int n_lines, int compute_coherence, float* out) {
    return oracle::despeck_block<Impl>((const oracle::cfloat*)z1, (const oracle::cfloat*)z2, wts, cols, lines, Nx, Ny,
                                       first_line, n_lines, compute_coherence, (oracle::cfloat*)out);
}

// Datum adjustment product (python/adjustMiniStacks.py:180-199): the reference writes a VRT whose
// "mul" pixel function multiplies the two complex sources; GDAL (third party, version unpinned by the
// reference) evaluates that in double -- starting from 1+0j, times source 1, times source 2 -- and
// stores CFloat32.  Restated with the operations spelled out (no complex operator*, no contraction);
// parity at this boundary is unpinned: GDAL is not available in this image.
void oracle_cmul(const float* a, const float* b, long n, float* out) {
    for (long i = 0; i < n; ++i) {
        volatile double ar = (double)a[2 * i], ai = (double)a[2 * i + 1];       // 1+0j times source 1: exact
        volatile double br = (double)b[2 * i], bi = (double)b[2 * i + 1];
        volatile double p0 = ar * br, p1 = ai * bi, p2 = ar * bi, p3 = ai * br;
        volatile double re = p0 - p1, im = p2 + p3;
        out[2 * i] = (float)re;
        out[2 * i + 1] = (float)im;
    }
}


// calamp: src/calamp/calamp.cpp:207-226 -- per band, sum of the valid amplitudes and their number (float hypot, NaN -> 0,
// valid = amplitude != 0 and mask > 0).  Serial double sums in raster order: the reference's OpenMP reduction has no
// defined order, so its own result is reproducible only to rounding (parity gate: relative 1e-12 on the sums).
void oracle_calamp_block(const float* slc, const unsigned char* mask, long npix, int bands, double* sums, double* counts) {
    const oracle::cfloat* z = reinterpret_cast<const oracle::cfloat*>(slc);
    for (int bb = 0; bb < bands; ++bb) {
        double blocksum = 0.0, blocknorm = 0.0;
        for (long ii = 0; ii < npix; ++ii) {
            double absval = std::abs(z[(long)bb * npix + ii]);
            absval = std::isnan(absval) ? 0.0 : absval;
            const int valid = (absval != 0.0) * (mask ? (mask[ii] > 0) : 1);
            blocksum += valid * absval;
            blocknorm += valid;
        }
        sums[bb] += blocksum;
        counts[bb] += blocknorm;
    }
}

// PS / DS integration, python/integratePS.py:97-130: numpy expressions restated in float arithmetic
// (complex64 product, angle, exp(1j * angle)); the reference evaluates them with numpy, whose float32 atan2 / sincos are
// not bit-reproducible across builds: parity gate 1e-6 absolute.
void oracle_integrate_ps(const float* ds_i, const float* ds_j, const float* slc_i, const float* slc_j, const unsigned char* ps,
                         long n, float* out) {
    for (long k = 0; k < n; ++k) {
        const bool is_ps = ps[k] == 1;
        const float* a = (is_ps ? slc_j : ds_j) + 2 * k;
        const float* b = (is_ps ? slc_i : ds_i) + 2 * k;
        volatile float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[1] * b[0], p3 = a[0] * b[1];
        float re = p0 + p1, im = p2 - p3;                       // a * conj(b)
        if (is_ps) { const float ang = std::atan2(im, re); re = std::cos(ang); im = std::sin(ang); }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

}  // extern "C"
