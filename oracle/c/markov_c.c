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

/* the body of _sequential_rts (ops.py:288-311) over `count` steps from a given carry (sm, sP): sms[count], sPs[count] */
static void rts_from(int d, int64_t count, const double* fms, const double* fPs, const double* As, const double* Qs,
                     double* sm, double* sP, double* sms, double* sPs) {
    for (int64_t n = count - 1; n >= 0; --n) {
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

/* _sequential_rts, return_full = False, H = e_0^T (ops.py:288-311): sms[N], sPs[N] */
void bnc_sequential_rts(int d, int64_t N, const double* fms, const double* fPs, const double* As, const double* Qs,
                        double* sms, double* sPs) {
    double sm[MAXD], sP[MAXD * MAXD];
    memcpy(sm, fms + (N - 1) * d, sizeof(double) * d);
    memcpy(sP, fPs + (N - 1) * d * d, sizeof(double) * d * d);
    rts_from(d, N, fms, fPs, As, Qs, sm, sP, sms, sPs);
}

/* ---------------------------------------------------------------------------------------------------------------
 * The temporally parallel form (parallel=True; ops.py:183-253, 314-354) on the host cores: lax.associative_scan is
 * restated as a time-blocked three-phase scan -- (1) every block folds its elements with the reference's operator,
 * in parallel over blocks; (2) the block aggregates are combined in sequence; (3) every block re-runs the plain
 * recursion from its incoming state, in parallel.  Same elements and operators as the reference, a different (but
 * associative-equivalent) bracketing of the combines.  H = e_0^T, D = 1, m0 = 0.
 */
typedef struct { double A[MAXD * MAXD], b[MAXD], C[MAXD * MAXD], J[MAXD * MAXD], eta[MAXD]; } felem;
typedef struct { double E[MAXD * MAXD], g[MAXD], L[MAXD * MAXD]; } selem;

/* parallel_filtering_element_ (ops.py:183-197) */
static void filt_element(int d, const double* A, const double* Q, double R, double y, felem* e) {
    double S = Q[0] + R, K[MAXD];
    for (int i = 0; i < d; ++i) K[i] = Q[i * d] / S;                  /* Q H^T S^-1 */
    for (int i = 0; i < d; ++i) {
        e->b[i] = K[i] * y;
        e->eta[i] = A[i] * y / S;                                     /* A^T H^T S^-1 y, A[0][i] */
        for (int j = 0; j < d; ++j) {
            e->A[i * d + j] = A[i * d + j] - K[i] * A[j];             /* A - K H A */
            e->C[i * d + j] = Q[i * d + j] - K[i] * Q[j];             /* Q - K H Q */
            e->J[i * d + j] = A[i] * A[j] / S;                        /* A^T H^T S^-1 H A */
        }
    }
}

static void mat_inv_chol(int d, const double* P, double* Pinv) {      /* inv (utils.py:22-27): cho_solve(chol(P), I) */
    double L[MAXD * MAXD], I[MAXD * MAXD];
    memset(I, 0, sizeof(I));
    for (int i = 0; i < d; ++i) I[i * d + i] = 1.0;
    chol(d, P, L);
    cho_solve(d, d, L, I, Pinv);
}

/* parallel_filtering_operator (ops.py:203-219) */
static void filt_combine(int d, const felem* e1, const felem* e2, felem* o) {
    double C1inv[MAXD * MAXD], M[MAXD * MAXD], L[MAXD * MAXD], temp[MAXD * MAXD], A2t[MAXD * MAXD], T[MAXD * MAXD];
    mat_inv_chol(d, e1->C, C1inv);
    for (int i = 0; i < d * d; ++i) M[i] = C1inv[i] + e2->J[i];
    chol(d, M, L);
    cho_solve(d, d, L, C1inv, temp);                                  /* solve(C1inv + J2, C1inv) */
    matmul(d, e2->A, temp, A2t);                                      /* A2 temp */
    matmul(d, A2t, e1->A, o->A);
    double v[MAXD];
    for (int i = 0; i < d; ++i) {
        double s = e1->b[i];
        for (int k = 0; k < d; ++k) s += e1->C[i * d + k] * e2->eta[k];
        v[i] = s;
    }
    for (int i = 0; i < d; ++i) {
        double s = e2->b[i];
        for (int k = 0; k < d; ++k) s += A2t[i * d + k] * v[k];
        o->b[i] = s;
    }
    matmul(d, A2t, e1->C, T);
    matmul_bt(d, T, e2->A, o->C);
    for (int i = 0; i < d * d; ++i) o->C[i] += e2->C[i];
    double A1t[MAXD * MAXD];                                          /* A1^T temp^T */
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += e1->A[k * d + i] * temp[j * d + k];
            A1t[i * d + j] = s;
        }
    for (int i = 0; i < d; ++i) {
        double s = e2->eta[i];
        for (int k = 0; k < d; ++k) s -= e2->J[i * d + k] * e1->b[k];
        v[i] = s;
    }
    for (int i = 0; i < d; ++i) {
        double s = e1->eta[i];
        for (int k = 0; k < d; ++k) s += A1t[i * d + k] * v[k];
        o->eta[i] = s;
    }
    matmul(d, A1t, e2->J, T);
    matmul(d, T, e1->A, o->J);
    for (int i = 0; i < d * d; ++i) o->J[i] += e1->J[i];
}

/* _parallel_kf (ops.py:237-253), blocked.  Returns ell; fms[N,d], fPs[N,d,d]. */
double bnc_parallel_kf_blocked(int d, int64_t N, const double* As, const double* Qs, const double* ys, const double* Rs,
                               const uint8_t* mask, const double* P0, double* fms, double* fPs, int nblocks) {
    if (nblocks > N) nblocks = (int)N;
    felem* agg = (felem*)malloc(sizeof(felem) * nblocks);
    felem* pre = (felem*)malloc(sizeof(felem) * nblocks);
    double* ells = (double*)calloc(nblocks, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nblocks; ++b) {
        int64_t n0 = N * b / nblocks, n1 = N * (b + 1) / nblocks;
        felem acc, e, r;
        for (int64_t n = n0; n < n1; ++n) {
            filt_element(d, As + n * d * d, n == 0 ? P0 : Qs + n * d * d, Rs[n], ys[n], &e);   /* Qs[0] := P0 (:223) */
            if (n == n0) acc = e; else { filt_combine(d, &acc, &e, &r); acc = r; }
        }
        agg[b] = acc;
    }
    pre[0] = agg[0];
    for (int b = 1; b < nblocks; ++b) filt_combine(d, &pre[b - 1], &agg[b], &pre[b]);
    double m0[MAXD] = {0, 0, 0, 0};
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nblocks; ++b) {
        int64_t n0 = N * b / nblocks, n1 = N * (b + 1) / nblocks;
        const double *mi = b ? pre[b - 1].b : m0, *Pi = b ? pre[b - 1].C : P0;
        ells[b] = bnc_sequential_kf(d, n1 - n0, As + n0 * d * d, Qs + n0 * d * d, ys + n0, Rs + n0, mask ? mask + n0 : NULL,
                                    mi, Pi, fms + n0 * d, fPs + n0 * d * d);
    }
    double ell = 0.0;
    for (int b = 0; b < nblocks; ++b) ell += ells[b];
    free(agg); free(pre); free(ells);
    return ell;
}

/* parallel_smoothing_element (ops.py:318-325) / last element (:314-315) */
static void smooth_element(int d, const double* A, const double* Q, const double* m, const double* P, int last, selem* e) {
    if (last) {
        memset(e->E, 0, sizeof(e->E));
        memcpy(e->g, m, sizeof(double) * d);
        memcpy(e->L, P, sizeof(double) * d * d);
        return;
    }
    double AP[MAXD * MAXD], Pp[MAXD * MAXD], Lc[MAXD * MAXD], X[MAXD * MAXD], EPp[MAXD * MAXD], T[MAXD * MAXD], Am[MAXD];
    matmul(d, A, P, AP);
    matmul_bt(d, AP, A, Pp);
    for (int i = 0; i < d * d; ++i) Pp[i] += Q[i];
    chol(d, Pp, Lc);
    cho_solve(d, d, Lc, AP, X);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) e->E[i * d + j] = X[j * d + i];
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += A[i * d + k] * m[k];
        Am[i] = s;
    }
    for (int i = 0; i < d; ++i) {
        double s = m[i];
        for (int k = 0; k < d; ++k) s -= e->E[i * d + k] * Am[k];
        e->g[i] = s;
    }
    matmul(d, e->E, Pp, EPp);
    matmul_bt(d, EPp, e->E, T);
    for (int i = 0; i < d * d; ++i) e->L[i] = P[i] - T[i];
}

/* parallel_smoothing_operator (ops.py:328-335): elem1 = the later steps, elem2 = the earlier one */
static void smooth_combine(int d, const selem* e1, const selem* e2, selem* o) {
    double T[MAXD * MAXD];
    matmul(d, e2->E, e1->E, o->E);
    for (int i = 0; i < d; ++i) {
        double s = e2->g[i];
        for (int k = 0; k < d; ++k) s += e2->E[i * d + k] * e1->g[k];
        o->g[i] = s;
    }
    matmul(d, e2->E, e1->L, T);
    matmul_bt(d, T, e2->E, o->L);
    for (int i = 0; i < d * d; ++i) o->L[i] += e2->L[i];
}

/* _parallel_rts (ops.py:338-354), blocked, return_full = False: sms[N], sPs[N] */
void bnc_parallel_rts_blocked(int d, int64_t N, const double* fms, const double* fPs, const double* As, const double* Qs,
                              double* sms, double* sPs, int nblocks) {
    if (nblocks > N) nblocks = (int)N;
    selem* agg = (selem*)malloc(sizeof(selem) * nblocks);
    selem* suf = (selem*)malloc(sizeof(selem) * nblocks);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nblocks; ++b) {
        int64_t n0 = N * b / nblocks, n1 = N * (b + 1) / nblocks;
        selem acc, e, r;
        for (int64_t n = n1 - 1; n >= n0; --n) {
            smooth_element(d, As + n * d * d, Qs + n * d * d, fms + n * d, fPs + n * d * d, n == N - 1, &e);
            if (n == n1 - 1) acc = e; else { smooth_combine(d, &acc, &e, &r); acc = r; }
        }
        agg[b] = acc;
    }
    suf[nblocks - 1] = agg[nblocks - 1];
    for (int b = nblocks - 2; b >= 0; --b) smooth_combine(d, &suf[b + 1], &agg[b], &suf[b]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nblocks; ++b) {
        int64_t n0 = N * b / nblocks, n1 = N * (b + 1) / nblocks;
        double sm[MAXD], sP[MAXD * MAXD];
        if (b == nblocks - 1) {
            memcpy(sm, fms + (N - 1) * d, sizeof(double) * d);
            memcpy(sP, fPs + (N - 1) * d * d, sizeof(double) * d * d);
        } else {  /* the smoothed state of the first step of the next block */
            memcpy(sm, suf[b + 1].g, sizeof(double) * d);
            memcpy(sP, suf[b + 1].L, sizeof(double) * d * d);
        }
        rts_from(d, n1 - n0, fms + n0 * d, fPs + n0 * d * d, As + n0 * d * d, Qs + n0 * d * d, sm, sP, sms + n0, sPs + n0);
    }
    free(agg); free(suf);
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

/* the same iteration with parallel=True: the filter and the smoother in the blocked temporally parallel form */
double bnc_vi_iteration_blocked(const ckernel* k, int lik, double param, int64_t N, const double* dt, const double* dts,
                                const double* y, const uint8_t* mask, int Q, const double* gx, const double* gw, double lr,
                                double* nat1, double* nat2, double* site_mean, double* site_cov,
                                double* As, double* Qs, double* fms, double* fPs, double* post_mean, double* post_var,
                                int nblocks) {
    int d = kdim(k);
    double P0[MAXD * MAXD];
    pinf(k, P0);
    for (int pass = 0; pass < 2; ++pass) {
        bnc_discretise(k, N, dt, As, Qs);
        bnc_parallel_kf_blocked(d, N, As, Qs, site_mean, site_cov, mask, P0, fms, fPs, nblocks);
        bnc_discretise(k, N, dts, As, Qs);
        bnc_parallel_rts_blocked(d, N, fms, fPs, As, Qs, post_mean, post_var, nblocks);
        if (pass == 0)
            bnc_vi_site_update(lik, param, N, y, post_mean, post_var, Q, gx, gw, lr, 1, nat1, nat2, site_mean, site_cov);
    }
    double ed = bnc_vi_expected_density(lik, param, N, y, post_mean, post_var, Q, gx, gw);
    bnc_discretise(k, N, dt, As, Qs);
    double ell = bnc_parallel_kf_blocked(d, N, As, Qs, site_mean, site_cov, mask, P0, fms, fPs, nblocks);
    double edp = bnc_gaussian_expected_log_lik(N, site_mean, post_mean, post_var, site_cov, mask);
    return -(ed - (edp - ell));
}
