/*
 * markov_c.c -- plain-C restatement of the reference's CPU path for one inference iteration of a
 * temporal Markov GP with a single-latent likelihood (ORACLE / CPU BASELINE; test infrastructure,
 * never linked into the product).
 *
 * It follows the reference AS WRITTEN (default CPU settings: parallel=False):
 *   - As = vmap(state_transition)(dt), Qs = Pinf - A Pinf A^T are MATERIALISED as [N,d,d] arrays
 *     before each filter and each smoother pass            bayesnewton/ops.py:274-278, 371-373
 *   - _sequential_kf: lax.scan body, Cholesky solve        ops.py:154-180, utils.py:14-19,376-396
 *   - _sequential_rts: reversed lax.scan body              ops.py:288-311
 *   - VI site update for a Bernoulli-probit / Gaussian likelihood with 20-point Gauss-Hermite
 *     cubature, ensure_psd, newton_update, damping, reparametrise
 *                              inference.py:21-39,65-90,170-195; cubature.py:198-246; basemodels.py:85-100
 *   - energy: expected log-lik, gaussian_expected_log_lik, filter log-lik
 *                              inference.py:197-222; basemodels.py:708-741; utils.py:510-531
 * The per-time-step ("vmapped") loops are OpenMP-parallel, as XLA's CPU backend parallelises
 * elementwise work; the two scans are sequential, as lax.scan is.
 * Matern-1/2, 3/2, 5/2, 7/2 (kernels.py:158-165,216-224,273-286,344-365), D = 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 4
static const double LOG2PI = 1.8378770664093453;
static const double INV2PI = 0.15915494309189535;

typedef struct {
    int family;        /* 1..4 = Matern 1/2, 3/2, 5/2, 7/2 */
    double variance, lengthscale;
} ckernel;

static int kdim(const ckernel* k) { return k->family; }

static void pinf(const ckernel* k, double* P) {
    int d = kdim(k);
    double v = k->variance, l = k->lengthscale;
    memset(P, 0, sizeof(double) * d * d);
    if (d == 1) { P[0] = v; }
    else if (d == 2) { P[0] = v; P[3] = 3.0 * v / (l * l); }
    else if (d == 3) {
        double kap = 5.0 / 3.0 * v / (l * l);
        P[0] = v; P[2] = -kap; P[4] = kap; P[6] = -kap; P[8] = 25.0 * v / (l * l * l * l);
    } else {
        double k1 = 7.0 / 5.0 * v / (l * l), k2 = 9.8 * v / (l * l * l * l);
        P[0] = v; P[2] = -k1; P[5] = k1; P[7] = -k2; P[8] = -k1; P[10] = k2; P[13] = -k2;
        P[15] = 343.0 * v / (l * l * l * l * l * l);
    }
}

static void transition(const ckernel* k, double dt, double* A) {
    int d = kdim(k);
    double l = k->lengthscale;
    if (d == 1) { A[0] = exp(-dt / l); return; }
    if (d == 2) {
        double lam = sqrt(3.0) / l, e = exp(-dt * lam);
        A[0] = e * (dt * lam + 1.0); A[1] = e * dt; A[2] = e * (dt * -lam * lam); A[3] = e * (dt * -lam + 1.0);
        return;
    }
    if (d == 3) {
        double lam = sqrt(5.0) / l, dl = dt * lam, e = exp(-dl), l2 = lam * lam;
        double M[9] = {lam * (0.5 * dl + 1.0), dl + 1.0, 0.5 * dt,
                       -0.5 * dl * l2, lam * (1.0 - dl), 1.0 - 0.5 * dl,
                       l2 * lam * (0.5 * dl - 1.0), l2 * (dl - 3.0), lam * (0.5 * dl - 2.0)};
        for (int i = 0; i < 9; ++i) A[i] = e * (dt * M[i] + ((i % 4 == 0) ? 1.0 : 0.0));
        return;
    }
    {
        double lam = sqrt(7.0) / l, l2 = lam * lam, l3 = l2 * lam, dl = dt * lam, dl2 = dl * dl, e = exp(-dl);
        double M[16] = {lam * (1.0 + 0.5 * dl + dl2 / 6.0), 1.0 + dl + 0.5 * dl2, 0.5 * dt * (1.0 + dl), dt * dt / 6,
                        -dl2 * l2 / 6.0, lam * (1.0 + 0.5 * dl - 0.5 * dl2), 1.0 + dl - 0.5 * dl2, dt * (0.5 - dl / 6.0),
                        l3 * dl * (dl / 6.0 - 0.5), dl * l2 * (0.5 * dl - 2.0), lam * (1.0 - 2.5 * dl + 0.5 * dl2),
                        1.0 - dl + dl2 / 6.0,
                        l2 * l2 * (dl - 1.0 - dl2 / 6.0), l3 * (3.5 * dl - 4.0 - 0.5 * dl2),
                        l2 * (4.0 * dl - 6.0 - 0.5 * dl2), lam * (1.5 * dl - 3.0 - dl2 / 6.0)};
        for (int i = 0; i < 16; ++i) A[i] = e * (dt * M[i] + ((i % 5 == 0) ? 1.0 : 0.0));
    }
}

static void matmul(int n, const double* A, const double* B, double* C) {  /* C = A B */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[k * n + j];
            C[i * n + j] = s;
        }
}
static void matmul_bt(int n, const double* A, const double* B, double* C) {  /* C = A B^T */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[j * n + k];
            C[i * n + j] = s;
        }
}

/* lower Cholesky of an n x n matrix (lower triangle read); NaN on non-PD */
static void chol(int n, const double* P, double* L) {
    memset(L, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double s = P[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        double ljj = sqrt(s);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double t = P[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / ljj;
        }
    }
}
/* X = (L L^T)^-1 B, B is n x c */
static void cho_solve(int n, int c, const double* L, const double* B, double* X) {
    for (int j = 0; j < c; ++j) {
        for (int i = 0; i < n; ++i) {
            double s = B[i * c + j];
            for (int k = 0; k < i; ++k) s -= L[i * n + k] * X[k * c + j];
            X[i * c + j] = s / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = X[i * c + j];
            for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * X[k * c + j];
            X[i * c + j] = s / L[i * n + i];
        }
    }
}

/* As[N,d,d], Qs[N,d,d]  (ops.py:274-278) */
void bnc_discretise(const ckernel* k, int64_t N, const double* dt, double* As, double* Qs) {
    int d = kdim(k);
    double P[MAXD * MAXD];
    pinf(k, P);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        double* A = As + n * d * d;
        double AP[MAXD * MAXD], APA[MAXD * MAXD];
        transition(k, dt[n], A);
        matmul(d, A, P, AP);
        matmul_bt(d, AP, A, APA);
        for (int i = 0; i < d * d; ++i) Qs[n * d * d + i] = P[i] - APA[i];
    }
}

/* _sequential_kf with H = e_0^T, D = 1 (ops.py:154-180).  mask may be NULL.  Returns ell. */
double bnc_sequential_kf(int d, int64_t N, const double* As, const double* Qs, const double* ys, const double* Rs,
                         const uint8_t* mask, const double* m0, const double* P0, double* fms, double* fPs) {
    double m[MAXD], P[MAXD * MAXD], ell = 0.0;
    memcpy(m, m0, sizeof(double) * d);
    memcpy(P, P0, sizeof(double) * d * d);
    for (int64_t n = 0; n < N; ++n) {
        const double *A = As + n * d * d, *Q = Qs + n * d * d;
        double mp[MAXD], AP[MAXD * MAXD], Pp[MAXD * MAXD];
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += A[i * d + k] * m[k];
            mp[i] = s;
        }
        matmul(d, A, P, AP);
        matmul_bt(d, AP, A, Pp);
        for (int i = 0; i < d * d; ++i) Pp[i] += Q[i];
        double obs_mean = mp[0];
        const double* HP = Pp;          /* first row of P_ */
        double S = HP[0] + Rs[n];
        /* mvn_logpdf (utils.py:376-396) */
        {
            int mk = mask && mask[n];
            double x = mk ? 0.0 : ys[n], mu = mk ? 0.0 : obs_mean, cov = mk ? INV2PI : S;
            double L = sqrt(cov), diff = x - mu;
            double sd = (diff / L) / L;
            ell += -0.5 * (diff * sd + LOG2PI + 2.0 * log(fabs(L)));
        }
        double L = sqrt(S), K[MAXD];
        for (int i = 0; i < d; ++i) K[i] = (HP[i] / L) / L;   /* solve(S, HP).T */
        double e = ys[n] - obs_mean;
        for (int i = 0; i < d; ++i) m[i] = mp[i] + K[i] * e;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) P[i * d + j] = Pp[i * d + j] - K[i] * HP[j];
        memcpy(fms + n * d, m, sizeof(double) * d);
        memcpy(fPs + n * d * d, P, sizeof(double) * d * d);
    }
    return ell;
}

/* _sequential_rts, return_full = False, H = e_0^T (ops.py:288-311): sms[N], sPs[N] */
void bnc_sequential_rts(int d, int64_t N, const double* fms, const double* fPs, const double* As, const double* Qs,
                        double* sms, double* sPs) {
    double sm[MAXD], sP[MAXD * MAXD];
    memcpy(sm, fms + (N - 1) * d, sizeof(double) * d);
    memcpy(sP, fPs + (N - 1) * d * d, sizeof(double) * d * d);
    for (int64_t n = N - 1; n >= 0; --n) {
        const double *A = As + n * d * d, *Q = Qs + n * d * d, *fm = fms + n * d, *fP = fPs + n * d * d;
        double pm[MAXD], AfP[MAXD * MAXD], pP[MAXD * MAXD], L[MAXD * MAXD], X[MAXD * MAXD], C[MAXD * MAXD];
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += A[i * d + k] * fm[k];
            pm[i] = s;
        }
        matmul(d, A, fP, AfP);
        matmul_bt(d, AfP, A, pP);
        for (int i = 0; i < d * d; ++i) pP[i] += Q[i];
        chol(d, pP, L);
        cho_solve(d, d, L, AfP, X);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) C[i * d + j] = X[j * d + i];
        double dm[MAXD], dP[MAXD * MAXD], T1[MAXD * MAXD], T2[MAXD * MAXD];
        for (int i = 0; i < d; ++i) dm[i] = sm[i] - pm[i];
        for (int i = 0; i < d * d; ++i) dP[i] = sP[i] - pP[i];
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += C[i * d + k] * dm[k];
            sm[i] = fm[i] + s;
        }
        matmul(d, C, dP, T1);
        matmul_bt(d, T1, C, T2);
        for (int i = 0; i < d * d; ++i) sP[i] = fP[i] + T2[i];
        sms[n] = sm[0];
        sPs[n] = sP[0];
    }
}

/* likelihoods: 1 = Gaussian(param), 2 = Bernoulli probit (likelihoods.py:828-852) */
static double log_lik(int lik, double param, double y, double f) {
    if (lik == 1) return -0.5 * log(2.0 * M_PI * param) - 0.5 * (y - f) * (y - f) / param;
    double p = 0.5 * (1.0 + erf(f / sqrt(2.0))) * (1.0 - 2e-3) + 1e-3;
    return log(y == 1.0 ? p : 1.0 - p);
}

/* VI statistics at one step (likelihoods.py:363-383, 727-753; cubature.py:198-246) */
static void var_exp(int lik, double param, double y, double m, double v, int Q, const double* gx, const double* gw,
                    double* E, double* dE, double* d2E) {
    int missing = isnan(y);
    if (missing) y = m;
    if (lik == 1) {
        *E = -0.5 * log(2.0 * M_PI) - 0.5 * log(param) - 0.5 * ((y - m) * (y - m) + v) / param;
        *dE = (y - m) / param;
        *d2E = -1.0 / param;
    } else {
        double sd = sqrt(v), iv = 1.0 / v, e = 0.0, d1 = 0.0, dv = 0.0;
        for (int q = 0; q < Q; ++q) {
            double f = sd * gx[q] + m, wl = gw[q] * log_lik(lik, param, y, f), df = f - m;
            e += wl;
            d1 += iv * df * wl;
            dv += (0.5 * (iv * iv * df * df) - 0.5 * iv) * wl;
        }
        *E = e; *dE = d1; *d2E = 2.0 * dv;
    }
    if (missing) { *E = 0.0; *dE = NAN; *d2E = NAN; }
}

/* the VI site update over all steps (inference.py:65-90,170-195): updates nat1, nat2, site mean/cov in place */
void bnc_vi_site_update(int lik, double param, int64_t N, const double* y, const double* pm, const double* pv,
                        int Q, const double* gx, const double* gw, double lr, int ensure_psd,
                        double* nat1, double* nat2, double* site_mean, double* site_cov) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        double E, j, h;
        var_exp(lik, param, y[n], pm[n], pv[n], Q, gx, gw, &E, &j, &h);
        if (ensure_psd) { double k = -h; k = (k < 0.0) ? 1e-2 : k; h = -k; }
        if (isnan(h)) h = -1e-6;
        if (isnan(j)) j = h * pm[n];
        double n1 = j - h * pm[n], n2 = -h;
        double r1 = (1.0 - lr) * nat1[n] + lr * n1, r2 = (1.0 - lr) * nat2[n] + lr * n2;
        nat1[n] = r1; nat2[n] = r2;
        double L = sqrt(r2);
        site_mean[n] = (r1 / L) / L;
        site_cov[n] = (1.0 / L) / L;
    }
}

/* sum_n E_q[log p(y_n|f_n)]  (inference.py:209-218) */
double bnc_vi_expected_density(int lik, double param, int64_t N, const double* y, const double* pm, const double* pv,
                               int Q, const double* gx, const double* gw) {
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t n = 0; n < N; ++n) {
        double E, j, h;
        var_exp(lik, param, y[n], pm[n], pv[n], Q, gx, gw, &E, &j, &h);
        if (!isnan(E)) tot += E;
    }
    return tot;
}

/* sum_n gaussian_expected_log_lik (utils.py:510-531), D = 1 */
double bnc_gaussian_expected_log_lik(int64_t N, const double* py, const double* pm, const double* pv,
                                     const double* pR, const uint8_t* mask) {
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t n = 0; n < N; ++n) {
        int mk = mask && mask[n];
        double mu = mk ? py[n] : pm[n], R = mk ? INV2PI : pR[n], V = mk ? 1e-20 : pv[n];
        double L = sqrt(R), diff = py[n] - mu;
        double ml = -0.5 * (diff * ((diff / L) / L) + LOG2PI + 2.0 * log(fabs(L)));
        tot += ml - 0.5 * ((V / L) / L);
    }
    return tot;
}

/* One train_op-equivalent iteration for MarkovVariationalGP, returning the energy:
 * inference() = F,S,U,F,S then energy() = V,L,X.  Scratch arrays are caller-provided:
 * As,Qs [N,d,d]; fms [N,d]; fPs [N,d,d]; post_mean,post_var [N].  */
double bnc_vi_iteration(const ckernel* k, int lik, double param, int64_t N, const double* dt, const double* dts,
                        const double* y, const uint8_t* mask, int Q, const double* gx, const double* gw, double lr,
                        double* nat1, double* nat2, double* site_mean, double* site_cov,
                        double* As, double* Qs, double* fms, double* fPs, double* post_mean, double* post_var) {
    int d = kdim(k);
    double P0[MAXD * MAXD], m0[MAXD] = {0, 0, 0, 0};
    pinf(k, P0);
    for (int pass = 0; pass < 2; ++pass) {
        bnc_discretise(k, N, dt, As, Qs);
        bnc_sequential_kf(d, N, As, Qs, site_mean, site_cov, mask, m0, P0, fms, fPs);
        bnc_discretise(k, N, dts, As, Qs);
        bnc_sequential_rts(d, N, fms, fPs, As, Qs, post_mean, post_var);
        if (pass == 0)
            bnc_vi_site_update(lik, param, N, y, post_mean, post_var, Q, gx, gw, lr, 1, nat1, nat2, site_mean, site_cov);
    }
    double ed = bnc_vi_expected_density(lik, param, N, y, post_mean, post_var, Q, gx, gw);
    bnc_discretise(k, N, dt, As, Qs);
    double ell = bnc_sequential_kf(d, N, As, Qs, site_mean, site_cov, mask, m0, P0, fms, fPs);
    double edp = bnc_gaussian_expected_log_lik(N, site_mean, post_mean, post_var, site_cov, mask);
    return -(ed - (edp - ell));
}
