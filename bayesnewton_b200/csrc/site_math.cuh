// Scalar-latent site mathematics shared by the stand-alone site kernels (sites_impl.cuh) and the fused iteration
// (iter_impl.cuh), written on `real` (real.cuh): cubature (or closed-form) likelihood statistics of the four schemes,
// the EP cavity and scale factor, ensure_psd, the natural-parameter Newton step and damping.  Reference:
//   VI      inference.py:170-195  + likelihoods.py:363-383 + cubature.py:198-246
//   EP      inference.py:238-284  + utils.py:534-541 + cubature.py:310-371
//   Newton  inference.py:105-128  + likelihoods.py:322-355
//   PL      inference.py:339-371  + cubature.py:374-435
//   newton_update inference.py:21-39; ensure_psd utils.py:89-96; damping inference.py:83-86
#pragma once
#include "common.cuh"
#include "core.cuh"
#include "probit_table.cuh"
#ifdef BN_REAL32
#include "probit_table32.cuh"
#endif

namespace BN_NS {
constexpr real kSqrt2 = 1.4142135623730951;
constexpr real kInvSqrt2Pi = 0.3989422804014327;

// one-dimensional cubature rule held by value (kernel parameter -> constant bank; the sums over
// the points then take their weights as instruction operands).  wx = w x, wxx = w x^2.
constexpr int kMaxQ1 = 64;
struct Cub1 {
    int Q;
    int pad_;
    real x[kMaxQ1], w[kMaxQ1], wx[kMaxQ1], wxx[kMaxQ1];
    real wd[kMaxQ1];       // w (x^2 - 1)
    real xmax;             // max |x|
    real bw, bwx, bwd;     // kPtC0Bias times the sums of w, wx, wd: what a biased table evaluation adds to the three sums
};

#ifndef BN_SITE_TAB_UNROLL
#define BN_SITE_TAB_UNROLL 4
#endif


// digamma / trigamma: upward recurrence to x >= 10, then the asymptotic series (error below 1e-15 there); what the
// derivatives of gammaln need for the Beta likelihood's Newton statistics
BN_DEV real digamma(real x) {
    real r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    const real i = 1.0 / x, i2 = i * i;
    const real ser = i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 - i2 * (1.0 / 132 - i2 * (691.0 / 32760 - i2 / 12))))));
    return r + log(x) - 0.5 * i - ser;
}
BN_DEV real trigamma(real x) {
    real r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    const real i = 1.0 / x, i2 = i * i;
    const real ser = i * i2 * (1.0 / 6 - i2 * (1.0 / 30 - i2 * (1.0 / 42 - i2 * (1.0 / 30 - i2 * (5.0 / 66 - i2 * (691.0 / 2730 - i2 * (7.0 / 6)))))));
    return r + i + 0.5 * i2 + ser;
}

// ------------------------------------------------------------------------------ single-latent likelihoods
template <int LIK, bool TAB = false>
struct Lik1 {
    real param;       // Gaussian variance / Poisson bin size / Student-t scale / Gamma shape / NegBin alpha / Beta scale
    const double* tab;  // TAB: probit log-density table (probit_table.cuh), fp64 builds only
    real param2 = 0.0;  // Student-t degrees of freedom / NegBin scale

    BN_DEV real prob(real f) const {
        if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) return 1.0 / (1.0 + exp(-f));
        else return 0.5 * (1.0 + erf(f / kSqrt2)) * (1.0 - 2e-3) + 1e-3;  // likelihoods.py:828-829
    }
    BN_DEV real log_lik(real y, real f) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            real r = y - f;
            return -0.5 * log(2.0 * 3.141592653589793 * param) - 0.5 * r * r / param;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            const real mu = exp(f) * param;  // likelihoods.py:939-940
            return y * log(mu) - mu - lgamma(y + 1.0);
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {  // likelihoods.py:1031-1041; param = scale, param2 = df
            const real df = param2, z = (y - f) / param;
            const real c = lgamma((df + 1.0) * 0.5) - lgamma(df * 0.5) - 0.5 * (log(param * param) + log(df) + log(3.141592653589793));
            return c - 0.5 * (df + 1.0) * log(1.0 + (1.0 / df) * (z * z));
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {  // likelihoods.py:1127-1134; param = shape, scale = exp(f)
            const real sc = exp(f);
            return -param * log(sc) - lgamma(param) + (param - 1.0) * log(y) - y / sc;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {  // likelihoods.py:1141-1149, 1179-1182; param = alpha, param2 = scale
            const real m = exp(f) * param2, k = 1.0 / param;
            return lgamma(k + y) - lgamma(y + 1.0) - lgamma(k) + y * log(m / (m + k)) - k * log(1.0 + m * param);
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {  // likelihoods.py:1081-1093; param = scale
            const real mean = prob(f), al = mean * param, be = param - al;
            const real yc = fmin(fmax(y, 1e-6), 1.0 - 1e-6);
            return (al - 1.0) * log(yc) + (be - 1.0) * log(1.0 - yc) + lgamma(al + be) - lgamma(al) - lgamma(be);
        } else {
            if constexpr (LIK == BN_LIK_BERNOULLI_PROBIT) {
#ifndef BN_REAL32
                if constexpr (TAB) return probit_log_phi(tab, y == 1.0 ? f : -f);  // log(1 - p(f)) = log p(-f)
#endif
            }
            real p = prob(f);
            return log(y == 1.0 ? p : 1.0 - p);
        }
    }
    // value and first two derivatives w.r.t. f (what jacrev gives, likelihoods.py:322-330)
    BN_DEV void derivs(real y, real f, real& ll, real& d1, real& d2) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            ll = log_lik(y, f);
            d1 = (y - f) / param;
            d2 = -1.0 / param;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            const real mu = exp(f) * param;
            ll = y * log(mu) - mu - lgamma(y + 1.0);
            d1 = y - mu;
            d2 = -mu;
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {
            ll = log_lik(y, f);
            const real r = y - f, den = param2 * param * param + r * r;
            d1 = (param2 + 1.0) * r / den;
            d2 = (param2 + 1.0) * (r * r - param2 * param * param) / (den * den);
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {
            ll = log_lik(y, f);
            const real t = y * exp(-f);
            d1 = -param + t;
            d2 = -t;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {
            ll = log_lik(y, f);
            const real m = exp(f) * param2, k = 1.0 / param, s = m + k;
            d1 = k * (y - m) / s;
            d2 = -k * m * (k + y) / (s * s);
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {
            ll = log_lik(y, f);
            const real mean = prob(f), al = mean * param, be = param - al;
            const real yc = fmin(fmax(y, 1e-6), 1.0 - 1e-6);
            const real dmu = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi, ddmu = -f * dmu;
            const real g = log(yc) - log(1.0 - yc) - digamma(al) + digamma(be);   // d ll / d alpha at beta = scale - alpha
            const real gp = -trigamma(al) - trigamma(be);
            const real da = param * dmu;
            d1 = g * da;
            d2 = gp * da * da + g * param * ddmu;
        } else {
            real p = prob(f), dp, ddp;
            if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) {
                real e = exp(f);
                dp = e / ((1.0 + e) * (1.0 + e));
                ddp = p * (1.0 - p) * (1.0 - 2.0 * p);
            } else {
                dp = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
                ddp = -f * dp;
            }
            bool one = (y == 1.0);
            real q = one ? p : 1.0 - p;
            real s = one ? 1.0 : -1.0;
            ll = log(q);
            real r = dp / q;
            d1 = s * r;
            d2 = s * ddp / q - r * r;
        }
    }
    // E[y|f], Var[y|f], dE[y|f]/df
    BN_DEV void moments(real f, real& E, real& V, real& dE) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            E = f; V = param; dE = 1.0;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            E = V = dE = exp(f) * param;  // likelihoods.py:952-959
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {  // likelihoods.py:1043-1044
            E = f; V = (param * param) * (param2 / (param2 - 2.0)); dE = 1.0;
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {  // likelihoods.py:1136-1138
            const real sc = exp(f);
            E = param * sc; V = param * (sc * sc); dE = E;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {  // likelihoods.py:1184-1189
            E = exp(f) * param2; V = E + E * E * param; dE = E;
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {  // likelihoods.py:1095-1097
            const real p = prob(f);
            E = p; V = (p - p * p) / (param + 1.0);
            dE = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
        } else {
            real p = prob(f);
            E = p; V = p - p * p;
            if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) {
                real e = exp(f);
                dE = e / ((1.0 + e) * (1.0 + e));
            } else {
                dE = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
            }
        }
    }
};

// 1x1 "Cholesky inverse": inv(x) as cho_solve would produce it
BN_DEV real inv1(real x) {
    real L = sqrt(x);
    return (1.0 / L) / L;
}

struct SiteStats1 { real mean, jac, hess, val; };

// (mean, jacobian, hessian, value) for a scalar latent.  val is the energy-term value of the method.
// RAW = the likelihood-level statistic evaluated AT (m, v) with nothing around it -- the reference's
// Likelihood.variational_expectation / moment_match / log_likelihood_gradients /
// statistical_linear_regression -- i.e. no cavity, no EP scale factor; PL returns (mu, dmu, omega)
// in (val, jac, hess).
template <int LIK, int METHOD, bool RAW = false, bool TAB = false, int UNR = BN_SITE_TAB_UNROLL>
BN_DEV SiteStats1 site_stats_1(const Lik1<LIK, TAB>& lik, real y, real m, real v, real n1, real n2, real power,
                               const Cub1& cub) {
    const int Q = cub.Q;
    const real* cx = cub.x;
    const real* cw = cub.w;
    SiteStats1 o;
    const bool missing = isnan(y);
    real mean = m, cov = v;
    if constexpr (METHOD == BN_METHOD_EP && !RAW) {  // compute_cavity, utils.py:534-541
        real pn2 = inv1(v + 1e-8);
        cov = inv1(pn2 - power * n2);
        mean = cov * (pn2 * m - power * n1);
    }
    if (missing) y = mean;
    real j, h, val;
    if constexpr (METHOD == BN_METHOD_NEWTON) {
        lik.derivs(y, mean, val, j, h);
    } else if constexpr (METHOD == BN_METHOD_VI && LIK == BN_LIK_GAUSSIAN) {
        real r = y - mean;
        val = -0.5 * log(2.0 * 3.141592653589793) - 0.5 * log(lik.param) - 0.5 * (r * r + cov) / lik.param;
        j = r / lik.param;
        h = -1.0 / lik.param;
    } else if constexpr (METHOD == BN_METHOD_VI && LIK == BN_LIK_POISSON_EXP) {
        // closed form, likelihoods.py:979-1008: E = y log b + y m - b exp(m + v/2) - log y!
        const real emc = lik.param * exp(mean + 0.5 * cov);
        val = y * log(lik.param) + y * mean - emc - lgamma(y + 1.0);
        j = y - emc;
        h = -emc;
    } else if constexpr (METHOD == BN_METHOD_EP && LIK == BN_LIK_GAUSSIAN) {
        real var = lik.param / power + cov;  // mvn_logpdf_and_derivs, utils.py:448-466
        real L = sqrt(var);
        real prec = (1.0 / L) / L;
        real r = y - mean;
        val = -0.5 * (r * (prec * r) + kLog2Pi + 2.0 * log(fabs(L)));
        j = prec * r;
        h = -prec;
        real Lc = sqrt(lik.param);  // pep_constant, utils.py:431-445
        val += 0.5 * ((1.0 - power) * kLog2Pi - log(power)) + 0.5 * (1.0 - power) * 2.0 * log(fabs(Lc));
    } else if constexpr (METHOD == BN_METHOD_VI) {
        // cubature.py:214-246 with f_i - m = sd x_i taken exactly:
        //   E = sum w l,  dE/dm = (sum w x l) / sd,  d2E/dm2 = (sum w x^2 l - sum w l) / v
        const real sd = sqrt(cov);
        real E = 0.0, S1 = 0.0, Dh = 0.0;  // sum w l, sum w x l, sum w (x^2 - 1) l
        bool done = false;
#ifndef BN_REAL32
        if constexpr (TAB && LIK == BN_LIK_BERNOULLI_PROBIT) {
            // every point inside the table's range (the usual case): evaluate in table coordinates,
            // log p(y | f) = g(+-f) with the sign folded into the affine map, no clamping per point.  The table returns
            // g + 15 (its c0 field is stored biased); the three sums are corrected once, by 15 x the sums of the weights.
            const real sm = (y == 1.0) ? mean : -mean, reach = cub.xmax * sd;
            if (sm - reach >= kPtLo && sm + reach < kPtHi) {
                const real sg = (y == 1.0) ? kPtInvH : -kPtInvH;
                const real a1 = sg * sd, a0 = fma(sg, mean, kPtOff);
                const uint32_t m20 = pt_mask20();
#pragma unroll UNR
                for (int q = 0; q < Q; ++q) {
                    const real l = probit_log_phi_sb(lik.tab, fma(a1, cx[q], a0), m20);
                    E = fma(cw[q], l, E);
                    S1 = fma(cub.wx[q], l, S1);
                    Dh = fma(cub.wd[q], l, Dh);
                }
                E -= cub.bw;
                S1 -= cub.bwx;
                Dh -= cub.bwd;
                done = true;
            }
        }
#elif defined(__CUDA_ARCH__)
        if constexpr (TAB && LIK == BN_LIK_BERNOULLI_PROBIT) {
            // fp32 build: the 8 KB cubic table of probit_table32.cuh (lik.tab addresses its float4 entries)
            const float sm = (y == 1.0) ? mean.v : -mean.v, reach = cub.xmax.v * sd.v;
            if (sm - reach >= kPt32Lo && sm + reach < kPt32Hi) {
                const float sg = (y == 1.0) ? kPt32InvH : -kPt32InvH;
                const float a1 = sg * sd.v, a0 = fmaf(sg, mean.v, -kPt32Lo * kPt32InvH);
                const float4* t4 = reinterpret_cast<const float4*>(lik.tab);
                float e = 0.0f, s1 = 0.0f, dh = 0.0f;
#pragma unroll UNR
                for (int q = 0; q < Q; ++q) {
                    const float l = probit32_eval(t4, fmaf(a1, cx[q].v, a0));
                    e = fmaf(cw[q].v, l, e);
                    s1 = fmaf(cub.wx[q].v, l, s1);
                    dh = fmaf(cub.wd[q].v, l, dh);
                }
                E = e; S1 = s1; Dh = dh;
                done = true;
            }
        }
#endif
        if (!done) {
            real S2 = 0.0;
#pragma unroll 2
            for (int q = 0; q < Q; ++q) {
                const real l = lik.log_lik(y, fma(sd, cx[q], mean));
                E = fma(cw[q], l, E);
                S1 = fma(cub.wx[q], l, S1);
                S2 = fma(cub.wxx[q], l, S2);
            }
            Dh = S2 - E;
        }
        const real poison = (mean - mean) + (sd - sd);  // NaN / inf inputs must come out as NaN
        val = E + poison;
        j = S1 / sd + poison;
        h = Dh / cov + poison;
    } else if constexpr (METHOD == BN_METHOD_EP) {
        // cubature.py:328-371 with f_i - m = sd x_i: Z = sum w p, dZ = C^-1 sd sum w x p,
        // d2Z = C^-1 sd^2 C^-1 sum w x^2 p - C^-1 Z,  p_i = exp(power * log-lik_i)
        const real sd = sqrt(cov), ic = inv1(cov);
        real Z = 0.0, Z1 = 0.0, Z2 = 0.0;
#pragma unroll 2
        for (int q = 0; q < Q; ++q) {
            const real p = exp(power * lik.log_lik(y, fma(sd, cx[q], mean)));
            Z = fma(cw[q], p, Z);
            Z1 = fma(cub.wx[q], p, Z1);
            Z2 = fma(cub.wxx[q], p, Z2);
        }
        const real poison = (mean - mean) + (sd - sd);
        Z += poison;
        const real dZ = ic * (sd * Z1);
        const real d2Z = ic * (sd * sd) * ic * Z2 - ic * Z;
        real Zc = fmax(Z, 1e-8);
        if (isnan(Z)) Zc = Z;  // fmax drops NaN; jnp.maximum propagates it
        val = log(Zc);
        real Zinv = 1.0 / Zc;
        j = Zinv * dZ;
        h = -j * j + Zinv * d2Z;
    } else {  // PL: statistical linear regression, cubature.py:374-435
        real sd = sqrt(cov);
        real mu = 0.0, dmu = 0.0;
        for (int q = 0; q < Q; ++q) {
            real E, V, dE;
            lik.moments(sd * cx[q] + mean, E, V, dE);
            mu += cw[q] * E;
            dmu += cw[q] * dE;
        }
        real S = 0.0, Cc = 0.0;
        for (int q = 0; q < Q; ++q) {
            real f = sd * cx[q] + mean;
            real E, V, dE;
            lik.moments(f, E, V, dE);
            S += cw[q] * ((E - mu) * (E - mu) + V);
            Cc += cw[q] * (f - mean) * (E - mu);
        }
        real omega = S - Cc * (Cc * inv1(cov));
        if constexpr (RAW) {
            val = mu; j = dmu; h = omega;
        } else {
            real res = y - mu;
            if (missing) { res = 0.0; omega = 1e6; }
            real dmo = dmu * inv1(omega);
            j = dmo * res;
            h = -dmo * dmu;
            val = 0.0;
        }
    }
    if constexpr (METHOD == BN_METHOD_EP && !RAW) {  // inference.py:263-267
        real cp = inv1(cov);
        real sf = cp * inv1(h + cp) / power;
        j = sf * j;
        h = sf * h;
    }
    if (missing && METHOD != BN_METHOD_PL && !(RAW && METHOD == BN_METHOD_EP)) {
        j = nan("");
        h = nan("");
        val = 0.0;
    }
    o.mean = mean; o.jac = j; o.hess = h; o.val = val;
    return o;
}

// utils.py:89-96 applied as -f(-H) on a scalar
BN_DEV real ensure_psd1(real h) {
    real k = -h;
    k = (k < 0.0) ? 1e-2 : k;
    return -k;
}

// ------------------------------------------------------------------------------ scalar-latent site update on values
// (y, posterior marginal (pm, pc), old natural parameters (o1, o2)) -> damped new natural parameters (r1, r2):
// the likelihood statistics of the scheme, ensure_psd (utils.py:89-96), newton_update (inference.py:21-39) and the
// damping of inference.py:83-86.  s / h receive the (mean, jacobian) and the hessian the reference returns as state;
// d1 / d2 the absolute change of the natural parameters before damping (the `diff` terms of inference.py:78-79).
template <int LIK, int METHOD, bool TAB, int UNR = BN_SITE_TAB_UNROLL>
BN_DEV void site_update_scalar(const Lik1<LIK, TAB>& lik, const Cub1& cub, real y, real pm, real pc, real o1,
                               real o2, real lr, real power, int ensure_psd, SiteStats1& s, real& h, real& r1,
                               real& r2, real& d1, real& d2) {
    s = site_stats_1<LIK, METHOD, false, TAB, UNR>(lik, y, pm, pc, o1, o2, power, cub);
    h = s.hess;
    if (ensure_psd && METHOD != BN_METHOD_PL) h = ensure_psd1(h);
    const real hh = isnan(h) ? -1e-6 : h;
    const real j = isnan(s.jac) ? hh * s.mean : s.jac;
    const real nn1 = j - hh * s.mean, nn2 = -hh;
    d1 = fabs(nn1 - o1);
    d2 = fabs(nn2 - o2);
    r1 = (1.0 - lr) * o1 + lr * nn1;
    r2 = (1.0 - lr) * o2 + lr * nn2;
}


// gaussian_expected_log_lik (utils.py:510-531) at step n, D in {1, 2}
template <int D>
BN_DEV real gaussian_ell_step(const real* py, const real* pm, const real* pV, const real* pR,
                                const unsigned char* mask, long long n) {
    real y[D], m[D], V[symn(D)], R[symn(D)];
    bool mk[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        mk[i] = mask && mask[n * D + i];
        y[i] = py[n * D + i];
        m[i] = mk[i] ? y[i] : pm[n * D + i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real r = pR[n * D * D + i * D + j], v = pV[n * D * D + i * D + j];
            if (mk[i] || mk[j]) { r = 0.0; v = 0.0; }
            if (i == j && mk[i]) { r = kInv2Pi; v = 1e-20; }
            R[sidx(i, j)] = r;
            V[sidx(i, j)] = v;
        }
    real e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) e[i] = y[i] - m[i];
    real ml = mvn_logpdf_masked<D>(R, e, nullptr);
    chol<D>(R);
    real B[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) B[i * D + j] = V[sidx(i, j)];
    chol_solve<D, D>(R, B);
    real tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) tr += B[i * D + i];
    return ml - 0.5 * tr;
}

// log N(pseudo_y | cav_mean, pseudo_var/power + cav_cov) [+ pep_constant]  (basemodels.py:247-262)
template <int D>
BN_DEV real ep_pseudo_step(real power, int with_const, const real* py, const real* pR, const real* pm,
                             const real* pV, const real* n1, const real* n2, const unsigned char* mask,
                             long long n) {
    real cm[D], cC[symn(D)];
    {
        real Vj[symn(D)], pn2[symn(D)], t[symn(D)];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Vj[sidx(i, j)] = pV[n * D * D + i * D + j] + (i == j ? 1e-8 : 0.0);
        sym_inverse<D>(Vj, pn2);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) t[sidx(i, j)] = pn2[sidx(i, j)] - power * n2[n * D * D + i * D + j];
        sym_inverse<D>(t, cC);
        real r[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            real s = -power * n1[n * D + i];
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(pn2[sidx(i, j)], pm[n * D + j], s);
            r[i] = s;
        }
        symvec<D>(cC, r, cm);
    }
    real S[symn(D)], e[D];
    unsigned char mk[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        mk[i] = mask ? mask[n * D + i] : 0;
        e[i] = py[n * D + i] - cm[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) S[sidx(i, j)] = pR[n * D * D + i * D + j] / power + cC[sidx(i, j)];
    }
    real val = mvn_logpdf_masked<D>(S, e, mask ? mk : nullptr);
    if (with_const) {  // pep_constant (utils.py:431-445)
        real Rr[symn(D)];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Rr[sidx(i, j)] = pR[n * D * D + i * D + j];
        chol<D>(Rr);
        real dim = D, ld = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            real l = log(fabs(Rr[sidx(i, i)]));
            if (mk[i]) { l = 0.0; dim -= 1.0; }
            ld += l;
        }
        val += 0.5 * dim * ((1.0 - power) * kLog2Pi - log(power)) + 0.5 * (1.0 - power) * 2.0 * ld;
    }
    return val;
}

// 1-D rule from host arrays (cubature.py:76-84 builds them with numpy on the host as well)
inline void make_cub1(int Q, const double* x, const double* w, Cub1& c) {
    c.Q = Q;
    c.pad_ = 0;
    c.xmax = 0.0;
    for (int q = 0; q < kMaxQ1; ++q) {
        const bool in = q < Q && x && w;
        c.x[q] = in ? x[q] : 0.0;
        c.w[q] = in ? w[q] : 0.0;
        c.wx[q] = c.w[q] * c.x[q];
        c.wxx[q] = c.w[q] * c.x[q] * c.x[q];
        c.wd[q] = c.wxx[q] - c.w[q];
        if (fabs(c.x[q]) > c.xmax) c.xmax = fabs(c.x[q]);
    }
    // the bias sums in the order (and with the fused multiply-adds) the kernels accumulate in
    real bw = 0.0, bwx = 0.0, bwd = 0.0;
    for (int q = 0; q < Q && q < kMaxQ1; ++q) {
        bw = fma(c.w[q], kPtC0Bias, bw);
        bwx = fma(c.wx[q], kPtC0Bias, bwx);
        bwd = fma(c.wd[q], kPtC0Bias, bwd);
    }
    c.bw = bw; c.bwx = bwx; c.bwd = bwd;
}

}  // namespace BN_NS
