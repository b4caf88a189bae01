// log-density of the reference's probit Bernoulli likelihood as a function of the latent,
//     g(f) = log( eps + (1 - 2 eps) Phi(f) ),   eps = 1e-3          (likelihoods.py:828-829, 836-852)
// so that  log p(y=1 | f) = g(f)  and  log p(y!=1 | f) = log(1 - p) = g(-f).
//
// The cubature sites evaluate it 20 times per time step; through erf() + log() that is ~100 fp64
// instructions per point and makes the site kernels the most expensive part of an iteration.  Here
// g is tabulated once per process as piecewise degree-6 polynomials on 577 intervals of width 1/32
// centred on the grid -9 + i/32 (Chebyshev interpolation in long double, max abs error 2.6e-15 =
// 3 ulp of |g| <= 6.9, measured by tests/test_probit_table.py); outside [-9, 9] g is constant to
// fp64.  One evaluation = 4 fp64 ops of index arithmetic + 6 DFMA + 7 shared-memory loads.
// The nearest singularities of g (zeros of eps + (1-2eps) Phi) sit ~0.95 from the real axis near
// f = -3.2, which is what forces the narrow intervals; the error bound is measured, not assumed.
#pragma once
#include <cmath>
#include <vector>
#include "smallmat.cuh"

namespace bn {

constexpr int kPtDeg = 6;
constexpr int kPtN = 577;            // interval centres -9 + i/32, i = 0..576
constexpr double kPtFmax = 9.0;
constexpr double kPtInvH = 32.0;
constexpr int kPtDoubles = (kPtDeg + 1) * kPtN;   // coefficient-major: tab[k * kPtN + i]

// g at table coordinate s = 32 f + 288, which must lie in [0, 576]
BN_DEV double probit_log_phi_s(const double* tab, double s) {
#ifdef __CUDA_ARCH__
    const double r = s + 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
    const int i = __double2loint(r);
    const double u = s - (r - 6755399441055744.0);         // in [-0.5, 0.5]
#else
    const double sr = nearbyint(s);
    const int i = (int)sr;
    const double u = s - sr;
#endif
    const double* c = tab + i;
    double p = c[kPtDeg * kPtN];
#pragma unroll
    for (int k = kPtDeg - 1; k >= 0; --k) p = fma(p, u, c[k * kPtN]);
    return p;
}

// tab -> g(f) for any f.  NaN inputs come back as a finite number (fmin/fmax drop NaN): callers poison.
BN_DEV double probit_log_phi(const double* tab, double f) {
    f = fmin(fmax(f, -kPtFmax), kPtFmax);
    return probit_log_phi_s(tab, fma(f, kPtInvH, kPtFmax * kPtInvH));
}

// host-side construction (long double), done once
inline long double probit_log_phi_ld(long double f) {
    const long double P = 0.5L * erfcl(-f / sqrtl(2.0L));
    return logl(1e-3L + (1.0L - 2e-3L) * P);
}

inline const std::vector<double>& probit_table_host() {
    static const std::vector<double> tab = [] {
        std::vector<double> t(kPtDoubles);
        constexpr int M = kPtDeg + 1;
        const long double pi = 3.14159265358979323846264338327950288L;
        for (int i = 0; i < kPtN; ++i) {
            const long double c = -(long double)kPtFmax + (long double)i / (long double)kPtInvH;
            // interpolate in v = 2u on [-1, 1] at the Chebyshev nodes, f = c + v / (2 * 32)
            long double fv[M], ck[M];
            for (int j = 0; j < M; ++j) {
                const long double v = cosl(pi * (j + 0.5L) / M);
                fv[j] = probit_log_phi_ld(c + v / (2.0L * (long double)kPtInvH));
            }
            for (int k = 0; k < M; ++k) {
                long double s = 0;
                for (int j = 0; j < M; ++j) s += fv[j] * cosl(pi * k * (j + 0.5L) / M);
                ck[k] = 2 * s / M;
            }
            ck[0] /= 2;
            // Chebyshev -> monomial in v
            long double mono[M] = {0}, T0[M] = {0}, T1[M] = {0}, T2[M];
            T0[0] = 1;
            T1[1] = 1;
            for (int k = 0; k < M; ++k) {
                const long double* T = (k == 0) ? T0 : T1;
                if (k >= 2) {
                    for (int j = 0; j < M; ++j) T2[j] = (j > 0 ? 2 * T1[j - 1] : 0) - T0[j];
                    for (int j = 0; j < M; ++j) { T0[j] = T1[j]; T1[j] = T2[j]; }
                    T = T1;
                }
                for (int j = 0; j < M; ++j) mono[j] += ck[k] * T[j];
            }
            // v = 2u: coefficient of u^k is mono[k] * 2^k
            long double sc = 1;
            for (int k = 0; k < M; ++k) {
                t[(size_t)k * kPtN + i] = (double)(mono[k] * sc);
                sc *= 2;
            }
        }
        return t;
    }();
    return tab;
}

}  // namespace bn
