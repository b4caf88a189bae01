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
#include "site_math.cuh"

namespace bn {

// what a site kernel needs besides bn_site_args: the 1-D rule by value, the probit table (shared
// memory on the device, null = evaluate through erf/log), the multi-latent rule in device memory
struct SiteCtx {
    const Cub1* cub;
    const double* tab;
    const double* cx2;   // [2, Q]
    const double* cw2;   // [Q]
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
