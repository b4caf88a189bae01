// fp32 companion of probit_table.cuh for the fp32 build of the fused iteration: g(f) = log(eps + (1 - 2 eps) Phi(f))
// (likelihoods.py:828-829) as cubic Taylor polynomials about the midpoints of 512 intervals of width 1/32 on [-8, 8),
// four floats per interval = ONE LDS.128 per evaluation, 8 KB in all (several CTAs per SM keep their own copy; lanes of a
// warp mostly hit a handful of entries, so the gather is largely a broadcast).  Truncation error h^4 g''''/24/16 < 1e-8,
// below the float rounding of g itself: the fp32 mode's bar is 1e-4 on the posterior.
#pragma once
#include <cmath>
#include <vector>
#include "real.cuh"

namespace BN_NS {

constexpr int kPt32N = 512;
constexpr float kPt32Lo = -8.0f, kPt32Hi = 8.0f, kPt32InvH = 32.0f;
constexpr int kPt32Bytes = kPt32N * 16;

inline const std::vector<float>& probit_table32_host() {
    static const std::vector<float> tab = [] {
        std::vector<float> t((size_t)kPt32N * 4);
        const long double eps = 1e-3L, h = 1.0L / 32.0L, is2pi = 0.39894228040143267793994605993438L;
        for (int i = 0; i < kPt32N; ++i) {
            const long double f = -8.0L + (i + 0.5L) * h;
            const long double Phi = 0.5L * erfcl(-f / sqrtl(2.0L));
            const long double p = eps + (1.0L - 2.0L * eps) * Phi;
            const long double p1 = (1.0L - 2.0L * eps) * is2pi * expl(-0.5L * f * f), p2 = -f * p1, p3 = (f * f - 1.0L) * p1;
            const long double r = p1 / p;
            const long double g0 = logl(p), g1 = r, g2 = p2 / p - r * r, g3 = p3 / p - 3.0L * r * (p2 / p) + 2.0L * r * r * r;
            t[4 * i + 0] = (float)g0;
            t[4 * i + 1] = (float)(g1 * h);
            t[4 * i + 2] = (float)(g2 * h * h / 2.0L);
            t[4 * i + 3] = (float)(g3 * h * h * h / 6.0L);
        }
        return t;
    }();
    return tab;
}

#ifdef __CUDACC__
// t = (f - lo) / h in [0, kPt32N): interval index and centred in-interval coordinate from one truncation
__device__ __forceinline__ float probit32_eval(const float4* tab, float t) {
    const int i = (int)t;
    const float u = t - (float)i - 0.5f;
    const float4 c = tab[i];
    return fmaf(fmaf(fmaf(c.w, u, c.z), u, c.y), u, c.x);
}
#endif

}  // namespace BN_NS
