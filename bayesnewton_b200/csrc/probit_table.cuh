// log-density of the reference's probit Bernoulli likelihood as a function of the latent,
//     g(f) = log( eps + (1 - 2 eps) Phi(f) ),   eps = 1e-3          (likelihoods.py:828-829, 836-852)
// so that  log p(y=1 | f) = g(f)  and  log p(y!=1 | f) = log(1 - p) = g(-f).
//
// The cubature sites evaluate it 20 times per time step; through erf() + log() that is ~100 fp64
// instructions per point.  Here g is tabulated once per process as piecewise cubics on 8192 intervals of width
// 1/512 covering [-8.5, 7.5) (outside it g is constant to 3e-14); the nearest singularities of g (zeros of
// eps + (1-2eps) Phi) sit ~0.95 from the real axis near f = -3.2, which is what forces the narrow intervals.
//
// The table lives in shared memory and is gathered with a different index per lane, so the cost of an evaluation
// is the number of shared-memory wavefronts, i.e. the BYTES per interval: a cubic to 1e-13 needs ~125 bits of
// coefficients (c0 to 2^-46, c1 to 2^-45, c2 to 2^-40, c3 to 2^-40 in the interval coordinate u in [-1/2, 1/2)), and
// they are packed in ONE 16-byte entry fetched by a single LDS.128 (the previous layout, doubles c0, c1 and a float
// pair, took three 8-byte gathers and bound the site kernels on the shared-memory pipe):
//     w.x                      low word of  D0 = 8 + (c0 + 7)           in [8, 16)      c0 = D0 - 15
//     w.y  [19:0]              high mantissa bits of D0
//     w.z  [19:0], w.w [31:14] mantissa of  D1 = 2^-7 + c1              in [2^-7, 2^-6) c1 = D1 - 2^-7   (38 bits)
//     w.y [31:20], w.z [31:21] 23-bit c2 in units of 2^-40, offset 2^-19
//     w.w  [13:0]              14-bit c3 in units of 2^-40, offset 2^-27
// Each field is the mantissa of a float / double with a FIXED exponent, so decoding is a mask-and-or plus one exact
// subtraction (no integer -> floating conversions); the bits of the neighbouring field that share a word with D1 are
// known when the table is built and are compensated there.  The quantisation of c3 and c2 is folded back into the
// lower coefficients (u^3 -> 3u/16, u^2 -> 1/8: what remains is a multiple of a Chebyshev polynomial), and the index
// and u come from the bits of ONE fma:  r = 8192 + 512 (f + 8.5)  has the interval number in its 13 leading
// mantissa bits and u + 1/2 in the 39 below.  One evaluation = 1 LDS.128, 6 fp64 operations, ~12 integer / fp32 ones.
// Max abs error 2.5e-13, typically 5e-14 (tests/test_probit_table.py against 40-digit arithmetic).  The table is 128 KB.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "real.cuh"
#include "smallmat.cuh"

namespace BN_NS {

constexpr int kPtN = 8192;            // intervals [i, i+1) of the table coordinate s = 512 (f - kPtLo)
constexpr double kPtLo = -8.5;
constexpr double kPtHi = 7.5;         // kPtLo + kPtN / 512
constexpr double kPtInvH = 512.0;
constexpr double kPtBias = 8192.0;    // r = s + kPtBias lies in [2^13, 2^14): fixed exponent
constexpr double kPtOff = -kPtLo * kPtInvH + kPtBias;   // r = 512 f + kPtOff
constexpr int kPtDoubles = 2 * kPtN;  // 16 bytes per interval

struct PtEntry { uint32_t x, y, z, w; };

#ifdef __CUDA_ARCH__
BN_DEV double pt_make_double(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
BN_DEV float pt_make_float(uint32_t b) { return __uint_as_float(b); }
#else
inline double pt_make_double(uint32_t hi, uint32_t lo) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
inline float pt_make_float(uint32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}
#endif

constexpr double kPtC0Bias = 15.0;  // the table stores D0 = c0 + 15 in [8, 16)

// (a & m) | e in one LOP3: the mask travels in a register so the exponent pattern can be the instruction's immediate
BN_DEV uint32_t pt_mask_or(uint32_t a, uint32_t m, uint32_t e) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(m), "r"(e));
    return d;
#else
    return (a & m) | e;
#endif
}

// the 20-bit mantissa mask as an opaque register value (a literal would be folded back into two-instruction and / or pairs)
BN_DEV uint32_t pt_mask20() {
#ifdef __CUDA_ARCH__
    uint32_t m;
    asm volatile("mov.u32 %0, 0xFFFFF;" : "=r"(m));
    return m;
#else
    return 0xFFFFFu;
#endif
}

// polynomial of one entry at u in [-1/2, 1/2), PLUS kPtC0Bias (callers that sum many evaluations remove the bias once)
BN_DEV double pt_eval_biased(const PtEntry& w, double u, uint32_t m20) {
    const double d0 = pt_make_double(pt_mask_or(w.y, m20, 0x40200000u), w.x);
    const double c1 = pt_make_double(pt_mask_or(w.z, m20, 0x3F800000u), w.w) - 0.0078125;
    const float c2 = pt_make_float(0x37000000u | ((w.y >> 9) & 0x7FF800u) | (w.z >> 21)) - 9.5367431640625e-6f;      // 2^-17 + 2^-19
    const float c3 = pt_make_float(0x37000000u | (w.w & 0x3FFFu)) - 7.636845111846924e-6f;                            // 2^-17 + 2^-27
    const float t = fmaf(c3, (float)u, c2);
    return fma(fma((double)t, u, c1), u, d0);
}

// g + kPtC0Bias at r = 512 f + kPtOff, which must lie in [2^13, 2^14)
BN_DEV double probit_log_phi_sb(const double* tab, double r, uint32_t m20) {
#ifdef __CUDA_ARCH__
    const uint32_t hi = (uint32_t)__double2hiint(r), lo = (uint32_t)__double2loint(r);
    const uint4 q = reinterpret_cast<const uint4*>(tab)[(hi >> 7) & 0x1FFFu];
    const PtEntry w{q.x, q.y, q.z, q.w};
    const uint32_t uh = pt_mask_or(__funnelshift_l(lo, hi, 13), m20, 0x3FF00000u);
#else
    uint64_t b;
    memcpy(&b, &r, 8);
    const uint32_t hi = (uint32_t)(b >> 32), lo = (uint32_t)b;
    PtEntry w;
    memcpy(&w, reinterpret_cast<const char*>(tab) + 16 * (size_t)((hi >> 7) & 0x1FFFu), 16);
    const uint32_t uh = 0x3FF00000u | (((hi << 13) | (lo >> 19)) & 0xFFFFFu);
#endif
    const double u = pt_make_double(uh, lo << 13) - 1.5;  // the 39 fraction bits of r as 1.f, minus 1.5
    return pt_eval_biased(w, u, m20);
}

// g at r = 512 f + kPtOff
BN_DEV double probit_log_phi_s(const double* tab, double r) { return probit_log_phi_sb(tab, r, pt_mask20()) - kPtC0Bias; }

// tab -> g(f) for any f.  NaN inputs come back as a finite number (fmin/fmax drop NaN): callers poison.
BN_DEV double probit_log_phi(const double* tab, double f) {
    f = fmin(fmax(f, kPtLo), kPtHi - 1e-6);
    return probit_log_phi_s(tab, fma(f, kPtInvH, kPtOff));
}

// host-side construction (long double), done once
inline long double probit_log_phi_ld(long double f) {
    const long double P = 0.5L * erfcl(-f / sqrtl(2.0L));
    return logl(1e-3L + (1.0L - 2e-3L) * P);
}

inline const std::vector<double>& probit_table_host() {
    static const std::vector<double> tab = [] {
        std::vector<double> t(kPtDoubles);
        constexpr int M = 4;
        const long double pi = 3.14159265358979323846264338327950288L;
        for (int i = 0; i < kPtN; ++i) {
            const long double c = (long double)kPtLo + ((long double)i + 0.5L) / (long double)kPtInvH;
            // cubic through the Chebyshev nodes of the interval, in v = 2u on [-1, 1]: f = c + v / (2 * 512)
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
            // Chebyshev -> monomial in u:  T0 = 1, T1 = 2u, T2 = 8u^2 - 1, T3 = 32u^3 - 6u
            long double m0 = ck[0] - ck[2], m1 = 2 * ck[1] - 6 * ck[3], m2 = 8 * ck[2], m3 = 32 * ck[3];
            // c3: 14 bits in units of 2^-40, offset 2^-27; its rounding error d u^3 = d (T3 + 3 T1) / 32 -> 3 d u / 16 moves to c1
            const long double lsb = ldexpl(1.0L, -40);
            long long q3 = llroundl((m3 + ldexpl(1.0L, -27)) / lsb);
            if (q3 < 0) q3 = 0;
            if (q3 > 0x3FFF) q3 = 0x3FFF;
            const long double c3 = (long double)q3 * lsb - ldexpl(1.0L, -27);
            m1 += (m3 - c3) * 3.0L / 16.0L;
            // c2: 23 bits in units of 2^-40, offset 2^-19; d u^2 = d (T2 + 1) / 8 -> d / 8 moves to c0
            long long q2 = llroundl((m2 + ldexpl(1.0L, -19)) / lsb);
            if (q2 < 0) q2 = 0;
            if (q2 > 0x7FFFFF) q2 = 0x7FFFFF;
            const long double c2 = (long double)q2 * lsb - ldexpl(1.0L, -19);
            m0 += (m2 - c2) / 8.0L;
            // c1: D1 = 2^-7 (1 + M1 / 2^52); the low 14 bits of M1 are c3's field: choose the upper 38 with them in place
            const long double x1 = (m1 + ldexpl(1.0L, -7)) / ldexpl(1.0L, -7) - 1.0L;   // in [0, 1)
            long long M1 = llroundl((x1 * ldexpl(1.0L, 52) - (long double)q3) / 16384.0L);
            if (M1 < 0) M1 = 0;
            const uint64_t m1bits = ((uint64_t)M1 << 14) | (uint64_t)q3;
            // c0: D0 = 8 (1 + M0 / 2^52) = c0 + 15, all 52 bits its own
            const long double x0 = (m0 + 15.0L) / 8.0L - 1.0L;
            long long M0 = llroundl(x0 * ldexpl(1.0L, 52));
            if (M0 < 0) M0 = 0;
            PtEntry e;
            e.x = (uint32_t)((uint64_t)M0 & 0xFFFFFFFFu);
            e.y = (uint32_t)(((uint64_t)M0 >> 32) & 0xFFFFFu) | ((uint32_t)(q2 >> 11) << 20);
            e.z = (uint32_t)((m1bits >> 32) & 0xFFFFFu) | ((uint32_t)(q2 & 0x7FF) << 21);
            e.w = (uint32_t)(m1bits & 0xFFFFFFFFu);
            memcpy(reinterpret_cast<char*>(t.data()) + 16 * (size_t)i, &e, 16);
        }
        return t;
    }();
    return tab;
}

}  // namespace BN_NS
