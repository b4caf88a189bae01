// log-density of the reference's probit Bernoulli likelihood as a function of the latent,
//     g(f) = log( eps + (1 - 2 eps) Phi(f) ),   eps = 1e-3          (likelihoods.py:828-829, 836-852)
// so that  log p(y=1 | f) = g(f)  and  log p(y!=1 | f) = log(1 - p) = g(-f).
//
// The cubature sites evaluate it 20 times per time step; through erf() + log() that is ~100 fp64
// instructions per point and makes the site kernels the most expensive part of an iteration.  Here
// g is tabulated once per process as piecewise cubics on 9217 intervals of width
// 1/512 centred on the grid -9 + i/512 (Chebyshev interpolation in long double); outside [-9, 9]
// g is constant to fp64.  The nearest singularities of g (zeros of eps + (1-2eps) Phi) sit ~0.95
// from the real axis near f = -3.2, which is what forces the narrow intervals.
//
// The table lives in shared memory and is gathered with a different index per lane, so the cost of
// an evaluation is the number of shared-memory wavefronts, i.e. the number of 8-byte words per
// interval: c0, c1 are doubles, the two highest coefficients are floats packed in one word
// (|c2 u^2| <= 5e-7: float rounding there is below 3e-14) and are combined on the fp32 pipe.
// One evaluation = 3 words (the degree-4 / width-1/256 table it replaces needed 4: the site kernels are bound by the
// shared-memory gathers), 4 fp64 ops of index arithmetic + 2 DFMA + 1 FFMA + 2 conversions; the table fills 216 KB of
// the 227 KB of shared memory a CTA can own.  Max abs error 1e-13 (tests/test_probit_table.py against long double).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "smallmat.cuh"

namespace bn {

constexpr int kPtDeg = 3;
constexpr int kPtN = 9217;           // interval centres -9 + i/512, i = 0..9216
constexpr double kPtFmax = 9.0;
constexpr double kPtInvH = 512.0;
constexpr int kPtDoubles = 3 * kPtN;   // word-major: c0[kPtN] | c1[kPtN] | {float c2, float c3}[kPtN]

// g at table coordinate s = 512 f + 4608, which must lie in [0, 9216]
BN_DEV double probit_log_phi_s(const double* tab, double s) {
#ifdef __CUDA_ARCH__
    const double r = s + 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
    const int i = __double2loint(r);
    const double u = s - (r - 6755399441055744.0);         // in [-0.5, 0.5]
    const float2 c23 = reinterpret_cast<const float2*>(tab + 2 * kPtN)[i];
    const float t = fmaf(c23.y, (float)u, c23.x);
#else
    const double sr = nearbyint(s);
    const int i = (int)sr;
    const double u = s - sr;
    float c23[2];
    memcpy(c23, tab + 2 * kPtN + i, sizeof(c23));
    const float t = fmaf(c23[1], (float)u, c23[0]);
#endif
    return fma(fma((double)t, u, tab[kPtN + i]), u, tab[i]);
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
            // interpolate in v = 2u on [-1, 1] at the Chebyshev nodes, f = c + v / (2 * 512)
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
            double cf[M];
            long double sc = 1;
            for (int k = 0; k < M; ++k) {
                cf[k] = (double)(mono[k] * sc);
                sc *= 2;
            }
            t[i] = cf[0];
            t[(size_t)kPtN + i] = cf[1];
            const float hi[2] = {(float)cf[2], (float)cf[3]};
            memcpy(&t[(size_t)2 * kPtN + i], hi, sizeof(hi));
        }
        return t;
    }();
    return tab;
}

}  // namespace bn
