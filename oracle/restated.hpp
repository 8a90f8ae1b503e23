// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FRInGE phase-linking hot path.
//
// This header is an independent restatement (plain C++, no Armadillo/GDAL) of the
// per-pair / per-matrix numerics of the reference.  It is never linked into the
// product library; only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
// Parity status: the reference ships no stored golden vectors for this path
// (SURVEY.md section 8c).  The restatement is pinned instead against the reference's
// own headers compiled in place (oracle/_ref, see oracle/Makefile) -- bit-for-bit on
// p-values, bitmask words and LAPACK results -- and against the few known answers the
// reference tests print (tests/eigen/test_eig.cpp 3x3 matrix, tests/bitmask layout,
// KS statistic == scipy.stats.ks_2samp).  tests/test_oracle_pins.py runs those checks.
//
// Each function cites the reference file:line it follows.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace restated {

// ---------------------------------------------------------------------------------
// Neighbourhood bitmask.  Follows include/fringe/ulongmask.hpp:57-95.
// Bit for window offset (dy,dx), dy in [-Ny,Ny], dx in [-Nx,Nx]:
//   flat = (dy+Ny)*(2Nx+1) + (dx+Nx);  word = flat/32;  bit = flat%32 (LSB first).
// ---------------------------------------------------------------------------------
struct WindowMask {
    int Ny, Nx;
    inline int flat(int dy, int dx) const { return (dy + Ny) * (2 * Nx + 1) + (dx + Nx); }
    inline void setbit(uint32_t* w, int dy, int dx, bool on) const {
        const int f = flat(dy, dx);
        const uint32_t m = uint32_t(1) << (f & 31);
        if (on) w[f >> 5] |= m; else w[f >> 5] &= ~m;
    }
    inline bool getbit(const uint32_t* w, int dy, int dx) const {
        const int f = flat(dy, dx);
        return ((w[f >> 5] >> (f & 31)) & 1u) != 0;
    }
};

// ---------------------------------------------------------------------------------
// Two-sample Kolmogorov-Smirnov on two ascending vectors of equal length.
// Follows src/nmap/KS2sample.hpp:42-77 (probability series, ROOT constants) and
// :91-144 (merge walk; ties consumed on both sides before the distance is sampled).
// The floating-point operation order is kept so p-values agree bit-for-bit.
// ---------------------------------------------------------------------------------
struct KSTest {
    int n;
    KSTest() : n(0) {}
    explicit KSTest(int len) : n(len) {}

    static double kolmogorov_prob(double z) {
        const double u = std::fabs(z);
        if (u < 0.2) return 1.0;
        if (u < 0.755) {
            const double k1 = -1.2337005501361697;   // -pi^2/8
            const double k2 = -11.103304951225528;   // 9*k1
            const double k3 = -30.842513753404244;   // 25*k1
            const double v = 1.0 / (u * u);
            return 1 - 2.50662827 * (::exp(k1 * v) + ::exp(k2 * v) + ::exp(k3 * v)) / u;
        }
        if (u < 6.8116) {
            static const double coef[4] = {-2, -8, -18, -32};
            double term[4] = {0, 0, 0, 0};
            const double v = u * u;
            const int nterm = std::max(1, (int)(::round(3.0 / u)));
            for (int j = 0; j < nterm; ++j) term[j] = ::exp(coef[j] * v);
            return 2 * (term[0] - term[1] + term[2] - term[3]);
        }
        return 0.0;
    }

    // Largest |F_a - F_b| in units of 1/n is also returned through *kmax (integer),
    // used to pin the integer formulation the CUDA kernel relies on.
    double test(const float* a, const float* b, int* kmax_out = nullptr) const {
        const double rn = n;
        const double step = 1.0 / rn;
        int ia = 0, ib = 0;
        double d = 0.0, dmax = 0.0;
        for (int it = 0; it < 2 * n; ++it) {
            if (a[ia] < b[ib]) { d -= step; ++ia; }
            else if (a[ia] > b[ib]) { d += step; ++ib; }
            else {
                const float tie = a[ia];
                while (ia < n && a[ia] == tie) { d -= step; ++ia; }
                while (ib < n && b[ib] == tie) { d += step; ++ib; }
            }
            if (ia >= n || ib >= n) break;
            dmax = std::max(dmax, std::fabs(d));
        }
        dmax = std::max(dmax, std::fabs(d));
        if (kmax_out) *kmax_out = (int)std::lround(dmax * rn);
        const double z = dmax * ::sqrt(rn * rn / (rn + rn));
        return kolmogorov_prob(z);
    }
};

// ---------------------------------------------------------------------------------
// Two-sample Anderson-Darling (ties not treated), equal lengths.
// Follows src/nmap/AD2unique.hpp:211-351 (merge + A_kN^2 sum), :162-208 (sigma_N),
// :125-160 (p-value: log-odds interpolation over column 0 of the ts[] table, :9-51).
// Only column 0 of the reference's 35x8 table is ever read (ts[i*8], :112), so only
// those 35 knots are restated here.
// ---------------------------------------------------------------------------------
struct ADTest {
    static constexpr int kBins = 35;
    int n;
    double sigma;
    double knot[kBins];
    double logodds[kBins];
    std::vector<unsigned char> from_a;   // merged indicator sequence (1 = element of a)

    ADTest() : n(0), sigma(0) {}
    explicit ADTest(int len) { init(len); }

    void init(int len) {
        static const double col0[kBins] = {
            -1.1954, -1.1786, -1.166, -1.1407, -1.1253, -1.0777, -1.0489, -0.9978, -0.9417,
            -0.8981, -0.8598, -0.7258, -0.5966, -0.4572, -0.2966, -0.1009, 0.1571, 0.5357,
            1.2255, 1.5262, 1.9633, 2.7314, 3.7825, 4.1241, 4.6044, 5.409, 6.4954, 6.8279,
            7.2755, 8.1885, 9.3061, 9.6132, 10.0989, 10.8825, 11.8537};
        static const double pbin[kBins] = {
            .00001, .00005, .0001, .0005, .001, .005, .01, .025, .05, .075, .1, .2, .3, .4, .5,
            .6, .7, .8, .9, .925, .95, .975, .99, .9925, .995, .9975, .999, .99925, .9995,
            .99975, .9999, .999925, .99995, .999975, .99999};
        n = len;
        for (int i = 0; i < kBins; ++i) {
            knot[i] = col0[i];
            logodds[i] = std::log((1.0 - pbin[i]) / pbin[i]);
        }
        from_a.assign(2 * n, 0);
        sigma = sigma_n(n, n);
    }

    // AD2unique.hpp:162-208 with k=2 samples.  Summation order preserved.
    static double sigma_n(int n1, int n2) {
        const int N = n1 + n2;
        const double k = 2.0;
        const double H = 1.0 / (1.0 * n1) + 1.0 / (1.0 * n2);
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
        const double k2 = std::pow(k, 2);
        const double a = (4 * g - 6) * (k - 1) + (10 - 6 * g) * H;
        const double b = (2 * g - 4) * k2 + 8 * h * k + (2 * g - 14 * h - 4) * H - 8 * h + 4 * g - 6;
        const double c = (6 * h + 2 * g - 2) * k2 + (4 * h - 4 * g + 6) * k + (2 * h - 6) * H + 4 * h;
        const double d = (2 * h + 6) * k2 - 4 * h * k;
        double s = 0.0;
        s += a * std::pow(double(N), 3) + b * std::pow(double(N), 2) + c * N + d;
        s /= (double(N - 1) * double(N - 2) * double(N - 3));
        return std::sqrt(s);
    }

    // AD2unique.hpp:125-160.
    double pvalue(double tx) const {
        int lo = -1;
        for (int i = 0; i < kBins; ++i) { lo = i - 1; if (tx <= knot[i]) break; }
        int hi = lo + 1;
        if (lo < 0) { lo = 0; hi = 1; }
        if (hi >= kBins) { lo = kBins - 2; hi = kBins - 1; }
        const double l1 = logodds[lo], l2 = logodds[hi];
        const double t1 = knot[lo], t2 = knot[hi];
        const double l0 = (l1 - l2) * (tx - t2) / (t1 - t2) + l2;
        return std::exp(l0) / (1. + std::exp(l0));
    }

    // The inner sum over the merged sequence for one sample (AD2unique.hpp:287-303);
    // `which`=1 sums for a, 0 for b.  Returned so tests can pin the table-sum route.
    double inner_sum(int which) const {
        const int L = 2 * n;
        const double nsum = (double)L;
        double m = 0.0, acc = 0.0;
        for (int j = 0; j < L; ++j) {
            m += (from_a[j] == which) ? 1 : 0;
            const double bj = j + 1;
            if (j < L - 1) {
                const double t = nsum * m - (double)n * bj;
                acc = acc + t * t / (bj * (nsum - bj));
            }
        }
        return acc;
    }

    double statistic_from_sum(double sa, double sb) const {
        double akn2 = 0.0;
        akn2 = akn2 + sa / (n * 1.0);
        akn2 = akn2 + sb / (n * 1.0);
        akn2 = akn2 / (double)(2 * n);
        double a2 = akn2 - 1;
        a2 /= sigma;
        return a2;
    }

    double test(const float* a, const float* b, double* sum_a = nullptr) {
        // merge; on equality the element of b goes first (AD2unique.hpp:228-241)
        int ia = 0, ib = 0, k = 0;
        while (ia < n && ib < n) {
            if (a[ia] < b[ib]) { from_a[k++] = 1; ++ia; }
            else { from_a[k++] = 0; ++ib; }
        }
        while (ib < n) { from_a[k++] = 0; ++ib; }
        while (ia < n) { from_a[k++] = 1; ++ia; }
        const double sa = inner_sum(1);
        const double sb = inner_sum(0);
        if (sum_a) *sum_a = sa;
        return pvalue(statistic_from_sum(sa, sb));
    }
};

// ---------------------------------------------------------------------------------
// LAPACK single-eigenpair / Cholesky-inverse wrapper.
// Follows include/fringe/EigenLapack.hpp:131-202 (zheevr, JOBZ V|N, RANGE 'I',
// UPLO 'U', ABSTOL 1e-6, workspace sizes :104-110) and :217-251 (zpotrf + zpotri +
// mirror upper->lower).  LAPACK itself is third-party (unpinned in the reference's
// CMake: FIND_PACKAGE(LAPACK)); here it is the OpenBLAS bundled in scipy's wheel.
// ---------------------------------------------------------------------------------
extern "C" {
void zheevr_(char*, char*, char*, int*, std::complex<double>*, int*, double*, double*, int*,
             int*, double*, int*, double*, std::complex<double>*, int*, int*,
             std::complex<double>*, int*, double*, int*, int*, int*, int*);
void zpotrf_(char*, int*, std::complex<double>*, int*, int*);
void zpotri_(char*, int*, std::complex<double>*, int*, int*);
}

struct EigSolver {
    int n = 0;
    std::vector<double> w;
    std::vector<std::complex<double>> z, cwork;
    std::vector<double> rwork;
    std::vector<int> iwork, isuppz;
    double* eigval = nullptr;
    std::complex<double>* eigvec = nullptr;

    void prepare(int order) {
        n = order;
        w.assign(n, 0.0);
        z.assign((size_t)n * n, 0.0);
        isuppz.assign(2 * n, 0);
        cwork.assign(2 * n, 0.0);
        rwork.assign(24 * n, 0.0);
        iwork.assign(10 * n, 0);
        eigval = w.data();
        eigvec = z.data();
    }
    int solve_index(std::complex<double>* A, int index, bool want_vec) {
        char jobz = want_vec ? 'V' : 'N', range = 'I', uplo = 'U';
        int N = n, lda = n, il = index, iu = index, m = 0, ldz = n, info = 0;
        double vl = 0.0, vu = 0.0, abstol = 1.0e-6;
        int lc = (int)cwork.size(), lr = (int)rwork.size(), li = (int)iwork.size();
        zheevr_(&jobz, &range, &uplo, &N, A, &lda, &vl, &vu, &il, &iu, &abstol, &m, w.data(),
                z.data(), &ldz, isuppz.data(), cwork.data(), &lc, rwork.data(), &lr,
                iwork.data(), &li, &info);
        return info;
    }
    int largestEigen(std::complex<double>* A, bool v) { return solve_index(A, n, v); }
    int smallestEigen(std::complex<double>* A, bool v) { return solve_index(A, 1, v); }
    int positiveDefiniteInverse(std::complex<double>* A) {
        char uplo = 'U';
        int N = n, lda = n, info = 0;
        zpotrf_(&uplo, &N, A, &lda, &info);
        if (info != 0) return info;
        zpotri_(&uplo, &N, A, &lda, &info);
        for (int r = 0; r < n; ++r)          // column-major: A[c*n + r]
            for (int c = r + 1; c < n; ++c) A[(size_t)r * n + c] = std::conj(A[(size_t)c * n + r]);
        return info;
    }
};

}  // namespace restated
