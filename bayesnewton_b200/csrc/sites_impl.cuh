// Fused per-time-step site kernels: cubature (or closed-form) likelihood statistics, the EP
// cavity and scale factor, ensure_psd, the natural-parameter Newton step, damping and the
// reparametrisation back to (pseudo_y, pseudo_var) -- one pass over HBM, nothing materialised
// in between.  Reference (all vmapped over N there):
//   VI      inference.py:170-195  + likelihoods.py:363-383 + cubature.py:198-246
//   EP      inference.py:238-284  + utils.py:534-541 + cubature.py:310-371
//   Newton  inference.py:105-128  + likelihoods.py:322-355
//   PL      inference.py:339-371  + cubature.py:374-435
//   multi-latent (HeteroscedasticNoise): likelihoods.py:561-664, 1244-1281 (autodiff forms written out)
//   newton_update inference.py:21-39; ensure_psd utils.py:89-96; damping inference.py:83-86;
//   reparametrise basemodels.py:85-100; Gaussian closed forms likelihoods.py:727-782
#pragma once
#include "common.cuh"
#include "core.cuh"
#include "probit_table.cuh"

namespace bn {

constexpr double kSqrt2 = 1.4142135623730951;
constexpr double kInvSqrt2Pi = 0.3989422804014327;

// one-dimensional cubature rule held by value (kernel parameter -> constant bank; the sums over
// the points then take their weights as instruction operands).  wx = w x, wxx = w x^2.
constexpr int kMaxQ1 = 64;
struct Cub1 {
    int Q;
    int pad_;
    double x[kMaxQ1], w[kMaxQ1], wx[kMaxQ1], wxx[kMaxQ1];
    double wd[kMaxQ1];       // w (x^2 - 1)
    double xmax;             // max |x|
    double bw, bwx, bwd;     // kPtC0Bias times the sums of w, wx, wd: what a biased table evaluation adds to the three sums
};

#ifndef BN_SITE_TAB_UNROLL
#define BN_SITE_TAB_UNROLL 4
#endif

// what a site kernel needs besides bn_site_args: the 1-D rule by value, the probit table (shared
// memory on the device, null = evaluate through erf/log), the multi-latent rule in device memory
struct SiteCtx {
    const Cub1* cub;
    const double* tab;
    const double* cx2;   // [2, Q]
    const double* cw2;   // [Q]
};

// digamma / trigamma: upward recurrence to x >= 10, then the asymptotic series (error below 1e-15 there); what the
// derivatives of gammaln need for the Beta likelihood's Newton statistics
BN_DEV double digamma(double x) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    const double i = 1.0 / x, i2 = i * i;
    const double ser = i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 - i2 * (1.0 / 132 - i2 * (691.0 / 32760 - i2 / 12))))));
    return r + log(x) - 0.5 * i - ser;
}
BN_DEV double trigamma(double x) {
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    const double i = 1.0 / x, i2 = i * i;
    const double ser = i * i2 * (1.0 / 6 - i2 * (1.0 / 30 - i2 * (1.0 / 42 - i2 * (1.0 / 30 - i2 * (5.0 / 66 - i2 * (691.0 / 2730 - i2 * (7.0 / 6)))))));
    return r + i + 0.5 * i2 + ser;
}

// ------------------------------------------------------------------------------ single-latent likelihoods
template <int LIK, bool TAB = false>
struct Lik1 {
    double param;       // Gaussian variance / Poisson bin size / Student-t scale / Gamma shape / NegBin alpha / Beta scale
    const double* tab;  // TAB: probit log-density table (probit_table.cuh)
    double param2 = 0.0;  // Student-t degrees of freedom / NegBin scale

    BN_DEV double prob(double f) const {
        if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) return 1.0 / (1.0 + exp(-f));
        else return 0.5 * (1.0 + erf(f / kSqrt2)) * (1.0 - 2e-3) + 1e-3;  // likelihoods.py:828-829
    }
    BN_DEV double log_lik(double y, double f) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            double r = y - f;
            return -0.5 * log(2.0 * 3.141592653589793 * param) - 0.5 * r * r / param;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            const double mu = exp(f) * param;  // likelihoods.py:939-940
            return y * log(mu) - mu - lgamma(y + 1.0);
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {  // likelihoods.py:1031-1041; param = scale, param2 = df
            const double df = param2, z = (y - f) / param;
            const double c = lgamma((df + 1.0) * 0.5) - lgamma(df * 0.5) - 0.5 * (log(param * param) + log(df) + log(3.141592653589793));
            return c - 0.5 * (df + 1.0) * log(1.0 + (1.0 / df) * (z * z));
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {  // likelihoods.py:1127-1134; param = shape, scale = exp(f)
            const double sc = exp(f);
            return -param * log(sc) - lgamma(param) + (param - 1.0) * log(y) - y / sc;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {  // likelihoods.py:1141-1149, 1179-1182; param = alpha, param2 = scale
            const double m = exp(f) * param2, k = 1.0 / param;
            return lgamma(k + y) - lgamma(y + 1.0) - lgamma(k) + y * log(m / (m + k)) - k * log(1.0 + m * param);
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {  // likelihoods.py:1081-1093; param = scale
            const double mean = prob(f), al = mean * param, be = param - al;
            const double yc = fmin(fmax(y, 1e-6), 1.0 - 1e-6);
            return (al - 1.0) * log(yc) + (be - 1.0) * log(1.0 - yc) + lgamma(al + be) - lgamma(al) - lgamma(be);
        } else {
            if constexpr (LIK == BN_LIK_BERNOULLI_PROBIT) {
                if constexpr (TAB) return probit_log_phi(tab, y == 1.0 ? f : -f);  // log(1 - p(f)) = log p(-f)
            }
            double p = prob(f);
            return log(y == 1.0 ? p : 1.0 - p);
        }
    }
    // value and first two derivatives w.r.t. f (what jacrev gives, likelihoods.py:322-330)
    BN_DEV void derivs(double y, double f, double& ll, double& d1, double& d2) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            ll = log_lik(y, f);
            d1 = (y - f) / param;
            d2 = -1.0 / param;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            const double mu = exp(f) * param;
            ll = y * log(mu) - mu - lgamma(y + 1.0);
            d1 = y - mu;
            d2 = -mu;
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {
            ll = log_lik(y, f);
            const double r = y - f, den = param2 * param * param + r * r;
            d1 = (param2 + 1.0) * r / den;
            d2 = (param2 + 1.0) * (r * r - param2 * param * param) / (den * den);
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {
            ll = log_lik(y, f);
            const double t = y * exp(-f);
            d1 = -param + t;
            d2 = -t;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {
            ll = log_lik(y, f);
            const double m = exp(f) * param2, k = 1.0 / param, s = m + k;
            d1 = k * (y - m) / s;
            d2 = -k * m * (k + y) / (s * s);
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {
            ll = log_lik(y, f);
            const double mean = prob(f), al = mean * param, be = param - al;
            const double yc = fmin(fmax(y, 1e-6), 1.0 - 1e-6);
            const double dmu = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi, ddmu = -f * dmu;
            const double g = log(yc) - log(1.0 - yc) - digamma(al) + digamma(be);   // d ll / d alpha at beta = scale - alpha
            const double gp = -trigamma(al) - trigamma(be);
            const double da = param * dmu;
            d1 = g * da;
            d2 = gp * da * da + g * param * ddmu;
        } else {
            double p = prob(f), dp, ddp;
            if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) {
                double e = exp(f);
                dp = e / ((1.0 + e) * (1.0 + e));
                ddp = p * (1.0 - p) * (1.0 - 2.0 * p);
            } else {
                dp = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
                ddp = -f * dp;
            }
            bool one = (y == 1.0);
            double q = one ? p : 1.0 - p;
            double s = one ? 1.0 : -1.0;
            ll = log(q);
            double r = dp / q;
            d1 = s * r;
            d2 = s * ddp / q - r * r;
        }
    }
    // E[y|f], Var[y|f], dE[y|f]/df
    BN_DEV void moments(double f, double& E, double& V, double& dE) const {
        if constexpr (LIK == BN_LIK_GAUSSIAN) {
            E = f; V = param; dE = 1.0;
        } else if constexpr (LIK == BN_LIK_POISSON_EXP) {
            E = V = dE = exp(f) * param;  // likelihoods.py:952-959
        } else if constexpr (LIK == BN_LIK_STUDENTS_T) {  // likelihoods.py:1043-1044
            E = f; V = (param * param) * (param2 / (param2 - 2.0)); dE = 1.0;
        } else if constexpr (LIK == BN_LIK_GAMMA_EXP) {  // likelihoods.py:1136-1138
            const double sc = exp(f);
            E = param * sc; V = param * (sc * sc); dE = E;
        } else if constexpr (LIK == BN_LIK_NEGBIN_EXP) {  // likelihoods.py:1184-1189
            E = exp(f) * param2; V = E + E * E * param; dE = E;
        } else if constexpr (LIK == BN_LIK_BETA_PROBIT) {  // likelihoods.py:1095-1097
            const double p = prob(f);
            E = p; V = (p - p * p) / (param + 1.0);
            dE = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
        } else {
            double p = prob(f);
            E = p; V = p - p * p;
            if constexpr (LIK == BN_LIK_BERNOULLI_LOGIT) {
                double e = exp(f);
                dE = e / ((1.0 + e) * (1.0 + e));
            } else {
                dE = (1.0 - 2e-3) * exp(-0.5 * f * f) * kInvSqrt2Pi;
            }
        }
    }
};

// ------------------------------------------------------------------------------ heteroscedastic (2 latents)
template <int LIK>
struct Lik2 {
    // log N(y | f1, g(f2)^2), gradient and Hessian w.r.t. (f1, f2); h = (h11, h12, h22)
    static BN_DEV void derivs(double y, double f1, double f2, double& ll, double* g, double* h) {
        double gg, g1, g2;
        if constexpr (LIK == BN_LIK_HETEROSCEDASTIC_EXP) {
            gg = g1 = g2 = exp(f2);
        } else {
            double e = exp(f2);
            gg = log(1.0 + e);          // softplus, naive form (utils.py:54-55)
            g1 = e / (e + 1.0);         // sigmoid (utils.py:59-60)
            g2 = g1 * (1.0 - g1);
        }
        double r = y - f1;
        double ig = 1.0 / gg, ig2 = ig * ig, ig3 = ig2 * ig;
        ll = -0.5 * log(2.0 * 3.141592653589793 * gg * gg) - 0.5 * r * r * ig2;
        g[0] = r * ig2;
        g[1] = -g1 * ig + r * r * g1 * ig3;
        h[0] = -ig2;
        h[1] = -2.0 * r * g1 * ig3;
        h[2] = -(g2 * gg - g1 * g1) * ig2 + r * r * (g2 * ig3 - 3.0 * g1 * g1 * ig2 * ig2);
    }
};

// 1x1 "Cholesky inverse": inv(x) as cho_solve would produce it
BN_DEV double inv1(double x) {
    double L = sqrt(x);
    return (1.0 / L) / L;
}

struct SiteStats1 { double mean, jac, hess, val; };

// (mean, jacobian, hessian, value) for a scalar latent.  val is the energy-term value of the method.
// RAW = the likelihood-level statistic evaluated AT (m, v) with nothing around it -- the reference's
// Likelihood.variational_expectation / moment_match / log_likelihood_gradients /
// statistical_linear_regression -- i.e. no cavity, no EP scale factor; PL returns (mu, dmu, omega)
// in (val, jac, hess).
template <int LIK, int METHOD, bool RAW = false, bool TAB = false, int UNR = BN_SITE_TAB_UNROLL>
BN_DEV SiteStats1 site_stats_1(const Lik1<LIK, TAB>& lik, double y, double m, double v, double n1, double n2, double power,
                               const Cub1& cub) {
    const int Q = cub.Q;
    const double* cx = cub.x;
    const double* cw = cub.w;
    SiteStats1 o;
    const bool missing = isnan(y);
    double mean = m, cov = v;
    if constexpr (METHOD == BN_METHOD_EP && !RAW) {  // compute_cavity, utils.py:534-541
        double pn2 = inv1(v + 1e-8);
        cov = inv1(pn2 - power * n2);
        mean = cov * (pn2 * m - power * n1);
    }
    if (missing) y = mean;
    double j, h, val;
    if constexpr (METHOD == BN_METHOD_NEWTON) {
        lik.derivs(y, mean, val, j, h);
    } else if constexpr (METHOD == BN_METHOD_VI && LIK == BN_LIK_GAUSSIAN) {
        double r = y - mean;
        val = -0.5 * log(2.0 * 3.141592653589793) - 0.5 * log(lik.param) - 0.5 * (r * r + cov) / lik.param;
        j = r / lik.param;
        h = -1.0 / lik.param;
    } else if constexpr (METHOD == BN_METHOD_VI && LIK == BN_LIK_POISSON_EXP) {
        // closed form, likelihoods.py:979-1008: E = y log b + y m - b exp(m + v/2) - log y!
        const double emc = lik.param * exp(mean + 0.5 * cov);
        val = y * log(lik.param) + y * mean - emc - lgamma(y + 1.0);
        j = y - emc;
        h = -emc;
    } else if constexpr (METHOD == BN_METHOD_EP && LIK == BN_LIK_GAUSSIAN) {
        double var = lik.param / power + cov;  // mvn_logpdf_and_derivs, utils.py:448-466
        double L = sqrt(var);
        double prec = (1.0 / L) / L;
        double r = y - mean;
        val = -0.5 * (r * (prec * r) + kLog2Pi + 2.0 * log(fabs(L)));
        j = prec * r;
        h = -prec;
        double Lc = sqrt(lik.param);  // pep_constant, utils.py:431-445
        val += 0.5 * ((1.0 - power) * kLog2Pi - log(power)) + 0.5 * (1.0 - power) * 2.0 * log(fabs(Lc));
    } else if constexpr (METHOD == BN_METHOD_VI) {
        // cubature.py:214-246 with f_i - m = sd x_i taken exactly:
        //   E = sum w l,  dE/dm = (sum w x l) / sd,  d2E/dm2 = (sum w x^2 l - sum w l) / v
        const double sd = sqrt(cov);
        double E = 0.0, S1 = 0.0, Dh = 0.0;  // sum w l, sum w x l, sum w (x^2 - 1) l
        bool done = false;
        if constexpr (TAB && LIK == BN_LIK_BERNOULLI_PROBIT) {
            // every point inside the table's range (the usual case): evaluate in table coordinates,
            // log p(y | f) = g(+-f) with the sign folded into the affine map, no clamping per point.  The table returns
            // g + 15 (its c0 field is stored biased); the three sums are corrected once, by 15 x the sums of the weights.
            const double sm = (y == 1.0) ? mean : -mean, reach = cub.xmax * sd;
            if (sm - reach >= kPtLo && sm + reach < kPtHi) {
                const double sg = (y == 1.0) ? kPtInvH : -kPtInvH;
                const double a1 = sg * sd, a0 = fma(sg, mean, kPtOff);
                const uint32_t m20 = pt_mask20();
#pragma unroll UNR
                for (int q = 0; q < Q; ++q) {
                    const double l = probit_log_phi_sb(lik.tab, fma(a1, cx[q], a0), m20);
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
        if (!done) {
            double S2 = 0.0;
#pragma unroll 2
            for (int q = 0; q < Q; ++q) {
                const double l = lik.log_lik(y, fma(sd, cx[q], mean));
                E = fma(cw[q], l, E);
                S1 = fma(cub.wx[q], l, S1);
                S2 = fma(cub.wxx[q], l, S2);
            }
            Dh = S2 - E;
        }
        const double poison = (mean - mean) + (sd - sd);  // NaN / inf inputs must come out as NaN
        val = E + poison;
        j = S1 / sd + poison;
        h = Dh / cov + poison;
    } else if constexpr (METHOD == BN_METHOD_EP) {
        // cubature.py:328-371 with f_i - m = sd x_i: Z = sum w p, dZ = C^-1 sd sum w x p,
        // d2Z = C^-1 sd^2 C^-1 sum w x^2 p - C^-1 Z,  p_i = exp(power * log-lik_i)
        const double sd = sqrt(cov), ic = inv1(cov);
        double Z = 0.0, Z1 = 0.0, Z2 = 0.0;
#pragma unroll 2
        for (int q = 0; q < Q; ++q) {
            const double p = exp(power * lik.log_lik(y, fma(sd, cx[q], mean)));
            Z = fma(cw[q], p, Z);
            Z1 = fma(cub.wx[q], p, Z1);
            Z2 = fma(cub.wxx[q], p, Z2);
        }
        const double poison = (mean - mean) + (sd - sd);
        Z += poison;
        const double dZ = ic * (sd * Z1);
        const double d2Z = ic * (sd * sd) * ic * Z2 - ic * Z;
        double Zc = fmax(Z, 1e-8);
        if (isnan(Z)) Zc = Z;  // fmax drops NaN; jnp.maximum propagates it
        val = log(Zc);
        double Zinv = 1.0 / Zc;
        j = Zinv * dZ;
        h = -j * j + Zinv * d2Z;
    } else {  // PL: statistical linear regression, cubature.py:374-435
        double sd = sqrt(cov);
        double mu = 0.0, dmu = 0.0;
        for (int q = 0; q < Q; ++q) {
            double E, V, dE;
            lik.moments(sd * cx[q] + mean, E, V, dE);
            mu += cw[q] * E;
            dmu += cw[q] * dE;
        }
        double S = 0.0, Cc = 0.0;
        for (int q = 0; q < Q; ++q) {
            double f = sd * cx[q] + mean;
            double E, V, dE;
            lik.moments(f, E, V, dE);
            S += cw[q] * ((E - mu) * (E - mu) + V);
            Cc += cw[q] * (f - mean) * (E - mu);
        }
        double omega = S - Cc * (Cc * inv1(cov));
        if constexpr (RAW) {
            val = mu; j = dmu; h = omega;
        } else {
            double res = y - mu;
            if (missing) { res = 0.0; omega = 1e6; }
            double dmo = dmu * inv1(omega);
            j = dmo * res;
            h = -dmo * dmu;
            val = 0.0;
        }
    }
    if constexpr (METHOD == BN_METHOD_EP && !RAW) {  // inference.py:263-267
        double cp = inv1(cov);
        double sf = cp * inv1(h + cp) / power;
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
BN_DEV double ensure_psd1(double h) {
    double k = -h;
    k = (k < 0.0) ? 1e-2 : k;
    return -k;
}

// ------------------------------------------------------------------------------ multi-latent statistics
struct SiteStats2 { double mean[2], jac[2], hess[4], val; };

BN_DEV void sym2_inverse(const double* M /*full 2x2, lower triangle read*/, double* out /*full*/) {
    double S[3] = {M[0], M[2], M[3]};
    double I[3];
    sym_inverse<2>(S, I);
    out[0] = I[0]; out[1] = I[1]; out[2] = I[1]; out[3] = I[2];
}

template <int LIK, int METHOD, bool RAW = false>
BN_DEV SiteStats2 site_stats_2(double y, const double* m, const double* V, const double* n1, const double* n2,
                               double power, int Q, const double* cx, const double* cw) {
    SiteStats2 o;
    double mean[2] = {m[0], m[1]};
    double cov[4] = {V[0], V[1], V[2], V[3]};
    if constexpr (METHOD == BN_METHOD_EP && !RAW) {
        double Vj[4] = {V[0] + 1e-8, V[1], V[2], V[3] + 1e-8};
        double pn2[4], t[4];
        sym2_inverse(Vj, pn2);
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = pn2[i] - power * n2[i];
        sym2_inverse(t, cov);
        double r0 = pn2[0] * m[0] + pn2[1] * m[1] - power * n1[0];
        double r1 = pn2[2] * m[0] + pn2[3] * m[1] - power * n1[1];
        mean[0] = cov[0] * r0 + cov[1] * r1;
        mean[1] = cov[2] * r0 + cov[3] * r1;
    }
    double g[2], h[3], ll;
    if constexpr (METHOD == BN_METHOD_NEWTON) {
        Lik2<LIK>::derivs(y, mean[0], mean[1], ll, g, h);
        o.val = ll;
        o.jac[0] = g[0]; o.jac[1] = g[1];
        o.hess[0] = h[0]; o.hess[1] = h[1]; o.hess[2] = h[1]; o.hess[3] = h[2];
    } else {
        // sigma points f = chol((V+V^T)/2) x + mean
        double S[3] = {cov[0], 0.5 * (cov[1] + cov[2]), cov[3]};
        chol<2>(S);
        const double* x0 = cx;
        const double* x1 = cx + Q;
        if constexpr (METHOD == BN_METHOD_VI) {
            double E = 0.0, dE[2] = {0.0, 0.0}, HH[3] = {0.0, 0.0, 0.0};
            for (int q = 0; q < Q; ++q) {
                double f1 = S[0] * x0[q] + mean[0];
                double f2 = S[1] * x0[q] + S[2] * x1[q] + mean[1];
                Lik2<LIK>::derivs(y, f1, f2, ll, g, h);
                double w = cw[q];
                E += w * ll;
                dE[0] += w * g[0]; dE[1] += w * g[1];
                HH[0] += w * h[0]; HH[1] += w * h[1]; HH[2] += w * h[2];
            }
            o.val = E;
            o.jac[0] = dE[0]; o.jac[1] = dE[1];
            o.hess[0] = HH[0]; o.hess[1] = HH[1]; o.hess[2] = HH[1]; o.hess[3] = HH[2];
        } else {  // EP, likelihoods.py:561-611 (no clamp on Z)
            double Z = 0.0, dZ[2] = {0.0, 0.0}, HH[3] = {0.0, 0.0, 0.0};
            for (int q = 0; q < Q; ++q) {
                double f1 = S[0] * x0[q] + mean[0];
                double f2 = S[1] * x0[q] + S[2] * x1[q] + mean[1];
                Lik2<LIK>::derivs(y, f1, f2, ll, g, h);
                double wp = cw[q] * exp(power * ll);
                Z += wp;
                dZ[0] += wp * power * g[0]; dZ[1] += wp * power * g[1];
                HH[0] += wp * (power * h[0] + power * power * g[0] * g[0]);
                HH[1] += wp * (power * h[1] + power * power * g[0] * g[1]);
                HH[2] += wp * (power * h[2] + power * power * g[1] * g[1]);
            }
            double Zi = 1.0 / Z;
            o.val = log(Z);
            double d0 = dZ[0] * Zi, d1 = dZ[1] * Zi;
            double H2[4] = {HH[0] * Zi - d0 * d0, HH[1] * Zi - d0 * d1, HH[1] * Zi - d0 * d1, HH[2] * Zi - d1 * d1};
            if constexpr (RAW) {
                o.jac[0] = d0; o.jac[1] = d1;
#pragma unroll
                for (int i = 0; i < 4; ++i) o.hess[i] = H2[i];
                o.mean[0] = mean[0]; o.mean[1] = mean[1];
                return o;
            }
            // scale factor (inference.py:263-267): cav_prec @ inv(d2 + cav_prec) / power
            double cp[4], t[4], ti[4], sf[4];
            sym2_inverse(cov, cp);
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = H2[i] + cp[i];
            sym2_inverse(t, ti);
            sf[0] = (cp[0] * ti[0] + cp[1] * ti[2]) / power; sf[1] = (cp[0] * ti[1] + cp[1] * ti[3]) / power;
            sf[2] = (cp[2] * ti[0] + cp[3] * ti[2]) / power; sf[3] = (cp[2] * ti[1] + cp[3] * ti[3]) / power;
            o.jac[0] = sf[0] * d0 + sf[1] * d1;
            o.jac[1] = sf[2] * d0 + sf[3] * d1;
            o.hess[0] = sf[0] * H2[0] + sf[1] * H2[2]; o.hess[1] = sf[0] * H2[1] + sf[1] * H2[3];
            o.hess[2] = sf[2] * H2[0] + sf[3] * H2[2]; o.hess[3] = sf[2] * H2[1] + sf[3] * H2[3];
        }
    }
    o.mean[0] = mean[0]; o.mean[1] = mean[1];
    return o;
}

// ------------------------------------------------------------------------------ scalar-latent site update on values
// (y, posterior marginal (pm, pc), old natural parameters (o1, o2)) -> damped new natural parameters (r1, r2):
// the likelihood statistics of the scheme, ensure_psd (utils.py:89-96), newton_update (inference.py:21-39) and the
// damping of inference.py:83-86.  s / h receive the (mean, jacobian) and the hessian the reference returns as state;
// d1 / d2 the absolute change of the natural parameters before damping (the `diff` terms of inference.py:78-79).
template <int LIK, int METHOD, bool TAB, int UNR = BN_SITE_TAB_UNROLL>
BN_DEV void site_update_scalar(const Lik1<LIK, TAB>& lik, const Cub1& cub, double y, double pm, double pc, double o1,
                               double o2, double lr, double power, int ensure_psd, SiteStats1& s, double& h, double& r1,
                               double& r2, double& d1, double& d2) {
    s = site_stats_1<LIK, METHOD, false, TAB, UNR>(lik, y, pm, pc, o1, o2, power, cub);
    h = s.hess;
    if (ensure_psd && METHOD != BN_METHOD_PL) h = ensure_psd1(h);
    const double hh = isnan(h) ? -1e-6 : h;
    const double j = isnan(s.jac) ? hh * s.mean : s.jac;
    const double nn1 = j - hh * s.mean, nn2 = -hh;
    d1 = fabs(nn1 - o1);
    d2 = fabs(nn2 - o2);
    r1 = (1.0 - lr) * o1 + lr * nn1;
    r2 = (1.0 - lr) * o2 + lr * nn2;
}

// ------------------------------------------------------------------------------ the fused update, one step
// returns |delta nat1| and |delta nat2| sums of this step through d1/d2
template <int LIK, int METHOD, bool TAB = false>
BN_DEV void site_update_step(const bn_site_args& a, const SiteCtx& sc, long long n, double& d1, double& d2) {
    if constexpr (LIK == BN_LIK_HETEROSCEDASTIC_SOFTPLUS || LIK == BN_LIK_HETEROSCEDASTIC_EXP) {
        const double* m = a.post_mean + 2 * n;
        const double* V = a.post_cov + 4 * n;
        double o1[2] = {a.nat1[2 * n], a.nat1[2 * n + 1]};
        double o2[4] = {a.nat2[4 * n], a.nat2[4 * n + 1], a.nat2[4 * n + 2], a.nat2[4 * n + 3]};
        SiteStats2 s = site_stats_2<LIK, METHOD>(a.y[n], m, V, o1, o2, a.power, a.Q, sc.cx2, sc.cw2);
        double H[4] = {s.hess[0], s.hess[1], s.hess[2], s.hess[3]};
        if (a.ensure_psd) {  // diagonal, negatives of -H replaced by 1e-2
            double k0 = -H[0], k1 = -H[3];
            k0 = (k0 < 0.0) ? 1e-2 : k0;
            k1 = (k1 < 0.0) ? 1e-2 : k1;
            H[0] = -k0; H[1] = -0.0; H[2] = -0.0; H[3] = -k1;
        }
        if (a.out_mean) { a.out_mean[2 * n] = s.mean[0]; a.out_mean[2 * n + 1] = s.mean[1]; }
        if (a.out_jac) { a.out_jac[2 * n] = s.jac[0]; a.out_jac[2 * n + 1] = s.jac[1]; }
        if (a.out_hess) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a.out_hess[4 * n + i] = H[i];
        }
        // newton_update (inference.py:21-39)
#pragma unroll
        for (int i = 0; i < 4; ++i) H[i] = isnan(H[i]) ? -1e-6 : H[i];
        double Hm0 = H[0] * s.mean[0] + H[1] * s.mean[1];
        double Hm1 = H[2] * s.mean[0] + H[3] * s.mean[1];
        double j0 = isnan(s.jac[0]) ? Hm0 : s.jac[0];
        double j1 = isnan(s.jac[1]) ? Hm1 : s.jac[1];
        double nn1[2] = {j0 - Hm0, j1 - Hm1};
        double nn2[4] = {-H[0], -H[1], -H[2], -H[3]};
        d1 = fabs(nn1[0] - o1[0]) + fabs(nn1[1] - o1[1]);
        d2 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) d2 += fabs(nn2[i] - o2[i]);
        double lr = a.lr;
        double r1[2] = {(1.0 - lr) * o1[0] + lr * nn1[0], (1.0 - lr) * o1[1] + lr * nn1[1]};
        double r2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r2[i] = (1.0 - lr) * o2[i] + lr * nn2[i];
        a.nat1[2 * n] = r1[0]; a.nat1[2 * n + 1] = r1[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) a.nat2[4 * n + i] = r2[i];
        if (a.site_mean || a.site_cov) {  // reparametrise (basemodels.py:85-90)
            double S[3] = {r2[0], r2[2], r2[3]};
            chol<2>(S);
            double B[6] = {r1[0], 1.0, 0.0, r1[1], 0.0, 1.0};  // [nat1 | I], 2 x 3
            chol_solve<2, 3>(S, B);
            if (a.site_mean) { a.site_mean[2 * n] = B[0]; a.site_mean[2 * n + 1] = B[3]; }
            if (a.site_cov) {
                a.site_cov[4 * n] = B[1]; a.site_cov[4 * n + 1] = B[2];
                a.site_cov[4 * n + 2] = B[4]; a.site_cov[4 * n + 3] = B[5];
            }
        }
    } else {
        Lik1<LIK, TAB> lik{a.lik_param, sc.tab, a.lik_param2};
        double o1 = a.nat1[n], o2 = a.nat2[n];
        SiteStats1 s;
        double h, r1, r2;
        site_update_scalar<LIK, METHOD, TAB>(lik, *sc.cub, a.y[n], a.post_mean[n], a.post_cov[n], o1, o2, a.lr, a.power,
                                             a.ensure_psd, s, h, r1, r2, d1, d2);
        if (a.out_mean) a.out_mean[n] = s.mean;
        if (a.out_jac) a.out_jac[n] = s.jac;
        if (a.out_hess) a.out_hess[n] = h;
        a.nat1[n] = r1;
        a.nat2[n] = r2;
        double L = sqrt(r2);
        if (a.site_mean) a.site_mean[n] = (r1 / L) / L;
        if (a.site_cov) a.site_cov[n] = (1.0 / L) / L;
    }
}

// value of the likelihood term of energy() at step n (NaN-safe: missing -> 0)
template <int LIK, int METHOD, bool TAB = false>
BN_DEV double expected_density_step(const bn_site_args& a, const SiteCtx& sc, long long n) {
    if constexpr (LIK == BN_LIK_HETEROSCEDASTIC_SOFTPLUS || LIK == BN_LIK_HETEROSCEDASTIC_EXP) {
        double o1[2] = {0.0, 0.0}, o2[4] = {0.0, 0.0, 0.0, 0.0};
        if (METHOD == BN_METHOD_EP) {
            o1[0] = a.nat1[2 * n]; o1[1] = a.nat1[2 * n + 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) o2[i] = a.nat2[4 * n + i];
        }
        SiteStats2 s = site_stats_2<LIK, METHOD>(a.y[n], a.post_mean + 2 * n, a.post_cov + 4 * n, o1, o2, a.power,
                                                 a.Q, sc.cx2, sc.cw2);
        return s.val;
    } else {
        Lik1<LIK, TAB> lik{a.lik_param, sc.tab, a.lik_param2};
        constexpr int M = (METHOD == BN_METHOD_PL) ? BN_METHOD_EP : METHOD;  // PL energy = EP energy at power 1
        double o1 = 0.0, o2 = 0.0;
        if (M == BN_METHOD_EP) { o1 = a.nat1[n]; o2 = a.nat2[n]; }
        SiteStats1 s = site_stats_1<LIK, M, false, TAB>(lik, a.y[n], a.post_mean[n], a.post_cov[n], o1, o2,
                                            METHOD == BN_METHOD_PL ? 1.0 : a.power, *sc.cub);
        return s.val;
    }
}

// gaussian_expected_log_lik (utils.py:510-531) at step n, D in {1, 2}
template <int D>
BN_DEV double gaussian_ell_step(const double* py, const double* pm, const double* pV, const double* pR,
                                const unsigned char* mask, long long n) {
    double y[D], m[D], V[symn(D)], R[symn(D)];
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
            double r = pR[n * D * D + i * D + j], v = pV[n * D * D + i * D + j];
            if (mk[i] || mk[j]) { r = 0.0; v = 0.0; }
            if (i == j && mk[i]) { r = kInv2Pi; v = 1e-20; }
            R[sidx(i, j)] = r;
            V[sidx(i, j)] = v;
        }
    double e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) e[i] = y[i] - m[i];
    double ml = mvn_logpdf_masked<D>(R, e, nullptr);
    chol<D>(R);
    double B[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) B[i * D + j] = V[sidx(i, j)];
    chol_solve<D, D>(R, B);
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) tr += B[i * D + i];
    return ml - 0.5 * tr;
}

// log N(pseudo_y | cav_mean, pseudo_var/power + cav_cov) [+ pep_constant]  (basemodels.py:247-262)
template <int D>
BN_DEV double ep_pseudo_step(double power, int with_const, const double* py, const double* pR, const double* pm,
                             const double* pV, const double* n1, const double* n2, const unsigned char* mask,
                             long long n) {
    double cm[D], cC[symn(D)];
    {
        double Vj[symn(D)], pn2[symn(D)], t[symn(D)];
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
        double r[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = -power * n1[n * D + i];
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(pn2[sidx(i, j)], pm[n * D + j], s);
            r[i] = s;
        }
        symvec<D>(cC, r, cm);
    }
    double S[symn(D)], e[D];
    unsigned char mk[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        mk[i] = mask ? mask[n * D + i] : 0;
        e[i] = py[n * D + i] - cm[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) S[sidx(i, j)] = pR[n * D * D + i * D + j] / power + cC[sidx(i, j)];
    }
    double val = mvn_logpdf_masked<D>(S, e, mask ? mk : nullptr);
    if (with_const) {  // pep_constant (utils.py:431-445)
        double Rr[symn(D)];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Rr[sidx(i, j)] = pR[n * D * D + i * D + j];
        chol<D>(Rr);
        double dim = D, ld = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double l = log(fabs(Rr[sidx(i, i)]));
            if (mk[i]) { l = 0.0; dim -= 1.0; }
            ld += l;
        }
        val += 0.5 * dim * ((1.0 - power) * kLog2Pi - log(power)) + 0.5 * (1.0 - power) * 2.0 * ld;
    }
    return val;
}

// likelihood-level statistics at (m, v) exactly as given (no cavity / scale factor / ensure_psd):
// val[N], d1[N,D,1], d2[N,D,D]
template <int LIK, int METHOD, bool TAB = false>
BN_DEV void likelihood_stats_step(const bn_site_args& a, const SiteCtx& sc, long long n, double* val, double* d1,
                                  double* d2) {
    if constexpr (LIK == BN_LIK_HETEROSCEDASTIC_SOFTPLUS || LIK == BN_LIK_HETEROSCEDASTIC_EXP) {
        double z1[2] = {0.0, 0.0}, z2[4] = {0.0, 0.0, 0.0, 0.0};
        SiteStats2 s = site_stats_2<LIK, METHOD, true>(a.y[n], a.post_mean + 2 * n, a.post_cov + 4 * n, z1, z2,
                                                       a.power, a.Q, sc.cx2, sc.cw2);
        if (val) val[n] = s.val;
        if (d1) { d1[2 * n] = s.jac[0]; d1[2 * n + 1] = s.jac[1]; }
        if (d2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) d2[4 * n + i] = s.hess[i];
        }
    } else {
        Lik1<LIK, TAB> lik{a.lik_param, sc.tab, a.lik_param2};
        SiteStats1 s = site_stats_1<LIK, METHOD, true, TAB>(lik, a.y ? a.y[n] : 0.0, a.post_mean[n], a.post_cov[n], 0.0,
                                                       0.0, a.power, *sc.cub);
        if (val) val[n] = s.val;
        if (d1) d1[n] = s.jac;
        if (d2) d2[n] = s.hess;
    }
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
    double bw = 0.0, bwx = 0.0, bwd = 0.0;
    for (int q = 0; q < Q && q < kMaxQ1; ++q) {
        bw = fma(c.w[q], kPtC0Bias, bw);
        bwx = fma(c.wx[q], kPtC0Bias, bwx);
        bwd = fma(c.wd[q], kPtC0Bias, bwd);
    }
    c.bw = bw; c.bwx = bwx; c.bwd = bwd;
}

#define BN_FOR_EACH_SITE(X)                                                                           \
    X(BN_LIK_GAUSSIAN, BN_METHOD_VI) X(BN_LIK_GAUSSIAN, BN_METHOD_EP)                                 \
    X(BN_LIK_GAUSSIAN, BN_METHOD_NEWTON) X(BN_LIK_GAUSSIAN, BN_METHOD_PL)                             \
    X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_VI) X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_EP)                 \
    X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_NEWTON) X(BN_LIK_BERNOULLI_PROBIT, BN_METHOD_PL)             \
    X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_VI) X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_EP)                   \
    X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_NEWTON) X(BN_LIK_BERNOULLI_LOGIT, BN_METHOD_PL)               \
    X(BN_LIK_POISSON_EXP, BN_METHOD_VI) X(BN_LIK_POISSON_EXP, BN_METHOD_EP)                           \
    X(BN_LIK_POISSON_EXP, BN_METHOD_NEWTON) X(BN_LIK_POISSON_EXP, BN_METHOD_PL)                       \
    X(BN_LIK_STUDENTS_T, BN_METHOD_VI) X(BN_LIK_STUDENTS_T, BN_METHOD_EP)                             \
    X(BN_LIK_STUDENTS_T, BN_METHOD_NEWTON) X(BN_LIK_STUDENTS_T, BN_METHOD_PL)                         \
    X(BN_LIK_GAMMA_EXP, BN_METHOD_VI) X(BN_LIK_GAMMA_EXP, BN_METHOD_EP)                               \
    X(BN_LIK_GAMMA_EXP, BN_METHOD_NEWTON) X(BN_LIK_GAMMA_EXP, BN_METHOD_PL)                           \
    X(BN_LIK_NEGBIN_EXP, BN_METHOD_VI) X(BN_LIK_NEGBIN_EXP, BN_METHOD_EP)                             \
    X(BN_LIK_NEGBIN_EXP, BN_METHOD_NEWTON) X(BN_LIK_NEGBIN_EXP, BN_METHOD_PL)                         \
    X(BN_LIK_BETA_PROBIT, BN_METHOD_VI) X(BN_LIK_BETA_PROBIT, BN_METHOD_EP)                           \
    X(BN_LIK_BETA_PROBIT, BN_METHOD_NEWTON) X(BN_LIK_BETA_PROBIT, BN_METHOD_PL)                       \
    X(BN_LIK_HETEROSCEDASTIC_SOFTPLUS, BN_METHOD_VI) X(BN_LIK_HETEROSCEDASTIC_SOFTPLUS, BN_METHOD_EP) \
    X(BN_LIK_HETEROSCEDASTIC_SOFTPLUS, BN_METHOD_NEWTON)                                              \
    X(BN_LIK_HETEROSCEDASTIC_EXP, BN_METHOD_VI) X(BN_LIK_HETEROSCEDASTIC_EXP, BN_METHOD_EP)           \
    X(BN_LIK_HETEROSCEDASTIC_EXP, BN_METHOD_NEWTON)


}  // namespace bn
