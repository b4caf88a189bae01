// Warp-cooperative Kalman filter / RTS smoother for any small state dimension d <= 16 (observation dimension <= d):
// ONE WARP per chunk of the time axis, the state and every temporary matrix of a step in shared memory, entries
// distributed over the lanes (lane l owns entries l, l + 32, ...), __syncwarp between dependent stages.  This is the
// route for the state dimensions the register-resident instantiations (core.cuh: d <= 6, one THREAD per chunk) cannot
// reach -- stacks of Independent Matern kernels, the pairs filter of larger sparse models (ops.py:154-180, 237-253,
// 288-311, 338-354 on caller arrays As, Qs).
//
// The chunk bodies are __host__ __device__: on the host they run with (lane, lanes) = (0, 1), which is what
// tests/hostemu executes against the oracle on the CPU.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "core.cuh"

namespace bn {

constexpr int kGdMaxD = 16;
constexpr int kGdWarps = 4;            // warps (chunks) per CTA
constexpr int kGdMinChunk = 16;        // steps per chunk at least

struct GdW { int lane, nl; };
#define GD_FOR(i, n) for (int i = w.lane; i < (n); i += w.nl)

BN_DEV void gd_sync() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}

BN_DEV void gd_copy(GdW w, double* dst, const double* src, int n) {
    GD_FOR(i, n) dst[i] = src[i];
    gd_sync();
}
BN_DEV void gd_fill(GdW w, double* dst, double v, int n) {
    GD_FOR(i, n) dst[i] = v;
    gd_sync();
}
BN_DEV void gd_eye(GdW w, double* dst, int n) {
    GD_FOR(i, n * n) dst[i] = (i / n == i % n) ? 1.0 : 0.0;
    gd_sync();
}

// C[m x n] = op(A) op(B) (+ add): A is m x k (k x m when ta), B is k x n (n x k when tb).  C must not alias A or B.
BN_DEV void gd_mm(GdW w, double* C, const double* A, const double* B, int m, int k, int n, bool ta, bool tb,
                  const double* add = nullptr, double scale = 1.0) {
    GD_FOR(idx, m * n) {
        const int i = idx / n, j = idx % n;
        double s = 0.0;
        for (int l = 0; l < k; ++l) s = fma(ta ? A[l * m + i] : A[i * k + l], tb ? B[j * k + l] : B[l * n + j], s);
        C[idx] = (add ? add[idx] : 0.0) + scale * s;
    }
    gd_sync();
}

// symmetric result: the lower triangle is computed, the upper mirrored (keeps covariances exactly symmetric)
BN_DEV void gd_mm_sym(GdW w, double* C, const double* A, const double* B, int n, int k, bool tb, const double* add,
                      double scale = 1.0) {
    GD_FOR(idx, n * n) {
        const int i = idx / n, j = idx % n;
        if (j > i) continue;
        double s = 0.0;
        for (int l = 0; l < k; ++l) s = fma(A[i * k + l], tb ? B[j * k + l] : B[l * n + j], s);
        const double v = (add ? add[i * n + j] : 0.0) + scale * s;
        C[i * n + j] = v;
        C[j * n + i] = v;
    }
    gd_sync();
}

// in-place lower Cholesky factor of the n x n matrix S (upper triangle left as it is); a non-PD input gives NaN
BN_DEV void gd_chol(GdW w, double* S, int n) {
    for (int j = 0; j < n; ++j) {
        if (w.lane == 0) {
            double s = S[j * n + j];
            for (int k = 0; k < j; ++k) s = fma(-S[j * n + k], S[j * n + k], s);
            S[j * n + j] = sqrt(s);
        }
        gd_sync();
        const double dj = S[j * n + j];
        GD_FOR(r, n - j - 1) {
            const int i = j + 1 + r;
            double s = S[i * n + j];
            for (int k = 0; k < j; ++k) s = fma(-S[i * n + k], S[j * n + k], s);
            S[i * n + j] = s / dj;
        }
        gd_sync();
    }
}

// B (n x c) <- (L L^T)^-1 B, one lane per column
BN_DEV void gd_chol_solve(GdW w, const double* L, double* B, int n, int c) {
    GD_FOR(col, c) {
        for (int i = 0; i < n; ++i) {
            double s = B[i * c + col];
            for (int k = 0; k < i; ++k) s = fma(-L[i * n + k], B[k * c + col], s);
            B[i * c + col] = s / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = B[i * c + col];
            for (int k = i + 1; k < n; ++k) s = fma(-L[k * n + i], B[k * c + col], s);
            B[i * c + col] = s / L[i * n + i];
        }
    }
    gd_sync();
}

// B (n x c) <- M^-1 B by Gaussian elimination with partial pivoting; M (n x n) is destroyed
BN_DEV void gd_lu_solve(GdW w, double* M, double* B, int n, int c) {
    for (int k = 0; k < n; ++k) {
        int p = k;  // every lane finds the same pivot row
        double best = fabs(M[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(M[i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        gd_sync();
        if (p != k) {
            GD_FOR(j, n + c) {
                double* a = j < n ? &M[k * n + j] : &B[k * c + (j - n)];
                double* b = j < n ? &M[p * n + j] : &B[p * c + (j - n)];
                const double t = *a; *a = *b; *b = t;
            }
            gd_sync();
        }
        const double pinv = 1.0 / M[k * n + k];
        GD_FOR(r, n - k - 1) {
            const int i = k + 1 + r;
            const double f = M[i * n + k] * pinv;
            for (int j = k + 1; j < n; ++j) M[i * n + j] = fma(-f, M[k * n + j], M[i * n + j]);
            for (int j = 0; j < c; ++j) B[i * c + j] = fma(-f, B[k * c + j], B[i * c + j]);
        }
        gd_sync();
    }
    GD_FOR(col, c) {
        for (int i = n - 1; i >= 0; --i) {
            double s = B[i * c + col];
            for (int k = i + 1; k < n; ++k) s = fma(-M[i * n + k], B[k * c + col], s);
            B[i * c + col] = s / M[i * n + i];
        }
    }
    gd_sync();
}

// ---------------------------------------------------------------------------------------------- shared-memory pool
// doubles a warp needs: the state, one filtering element, and the temporaries of the heaviest routine (the combine)
BN_DEV constexpr int gd_pool_doubles(int d) { return 20 * d * d + 16 * d + 64; }

struct GdPool {
    double* p;
    BN_DEV double* take(int n) { double* r = p; p += n; return r; }
};

// ---------------------------------------------------------------------------------------------- filter step
// (m, P) <- Kalman predict + update (ops.py:156-175); (mp, Pp) = predicted state; returns the log-likelihood term
// (utils.py:376-396 with the reference's mask rule) when want_ell.  H: D x d; R: D x D (lower triangle read).
BN_DEV double gd_kf_step(GdW w, int d, int D, double* m, double* P, const double* A, const double* Q, const double* H,
                         const double* y, const double* R, const unsigned char* msk, double* mp, double* Pp,
                         GdPool pool, bool want_ell) {
    double* T1 = pool.take(d * d);
    double* HP = pool.take(D * d);
    double* S = pool.take(D * D);
    double* Sm = pool.take(D * D);
    double* e = pool.take(D);
    double* x = pool.take(D);
    double* Kt = pool.take(D * d);
    gd_mm(w, mp, A, m, d, d, 1, false, false);
    gd_mm(w, T1, A, P, d, d, d, false, false);
    gd_mm_sym(w, Pp, T1, A, d, d, true, Q);
    gd_mm(w, HP, H, Pp, D, d, d, false, false);
    GD_FOR(idx, D * D) {
        const int i = idx / D, j = idx % D;
        if (j > i) continue;
        double s = R[i * D + j];
        for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], H[j * d + l], s);
        S[i * D + j] = s;
        S[j * D + i] = s;
    }
    GD_FOR(i, D) {
        double s = 0.0;
        for (int l = 0; l < d; ++l) s = fma(H[i * d + l], mp[l], s);
        e[i] = y[i] - s;
    }
    gd_sync();
    double ell = 0.0;
    if (want_ell) {
        GD_FOR(idx, D * D) {
            const int i = idx / D, j = idx % D;
            const bool mi = msk && msk[i], mj = msk && msk[j];
            double v = (mi || mj) ? 0.0 : S[idx];
            if (i == j && mi) v = kInv2Pi;
            Sm[idx] = v;
        }
        GD_FOR(i, D) x[i] = (msk && msk[i]) ? 0.0 : e[i];
        gd_sync();
        gd_chol(w, Sm, D);
        double logdet = 0.0, dist = 0.0;
        for (int i = 0; i < D; ++i) logdet += log(fabs(Sm[i * D + i]));
        // dist = em^T Sm^-1 em = |L^-1 em|^2: forward substitution, every lane the same (D is small)
        double z[kGdMaxD];
        for (int i = 0; i < D; ++i) {
            double s = x[i];
            for (int k = 0; k < i; ++k) s = fma(-Sm[i * D + k], z[k], s);
            z[i] = s / Sm[i * D + i];
            dist = fma(z[i], z[i], dist);
        }
        ell = -0.5 * (dist + D * kLog2Pi + 2.0 * logdet);
    }
    gd_chol(w, S, D);
    gd_copy(w, Kt, HP, D * d);
    gd_chol_solve(w, S, Kt, D, d);
    GD_FOR(i, d) {
        double s = mp[i];
        for (int a = 0; a < D; ++a) s = fma(Kt[a * d + i], e[a], s);
        m[i] = s;
    }
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (j > i) continue;
        double s = Pp[idx];
        for (int a = 0; a < D; ++a) s = fma(-Kt[a * d + i], HP[a * d + j], s);
        P[i * d + j] = s;
        P[j * d + i] = s;
    }
    gd_sync();
    return ell;
}

// ---------------------------------------------------------------------------------------------- filtering element
// Element (A, b, C, J, eta) with full d x d storage, contiguous: [A | b | C | J | eta] = 3 d^2 + 2 d doubles.
BN_DEV constexpr int gd_felem(int d) { return 3 * d * d + 2 * d; }
struct GdFElem {
    double *A, *b, *C, *J, *eta;
    BN_DEV GdFElem(double* p, int d) : A(p), b(p + d * d), C(p + d * d + d), J(p + 2 * d * d + d), eta(p + 3 * d * d + d) {}
};

BN_DEV void gd_felem_identity(GdW w, GdFElem g, int d) {
    gd_eye(w, g.A, d);
    gd_fill(w, g.b, 0.0, d * d + d + d * d + d);  // b, C, J, eta are contiguous
}

// absorb one step into the running aggregate of a chunk (core.cuh: filter_absorb; ops.py:183-219)
BN_DEV void gd_filter_absorb(GdW w, int d, int D, GdFElem g, const double* A, const double* Q, const double* H,
                             const double* y, const double* R, bool first, const double* m0, GdPool pool) {
    double* mp = pool.take(d);
    double* Pp = pool.take(d * d);
    double* T1 = pool.take(d * d);
    double* Phi = pool.take(d * d);
    double* HP = pool.take(D * d);
    double* HPhi = pool.take(D * d);
    double* S = pool.take(D * D);
    double* e = pool.take(D);
    double* Kt = pool.take(D * d);
    double* V = pool.take(D * (d + 1));
    gd_mm(w, mp, A, g.b, d, d, 1, false, false);
    gd_mm(w, T1, A, g.C, d, d, d, false, false);
    gd_mm_sym(w, Pp, T1, A, d, d, true, Q);
    gd_mm(w, Phi, A, g.A, d, d, d, false, false);
    gd_mm(w, HP, H, Pp, D, d, d, false, false);
    gd_mm(w, HPhi, H, Phi, D, d, d, false, false);
    GD_FOR(idx, D * D) {
        const int i = idx / D, j = idx % D;
        if (j > i) continue;
        double s = R[i * D + j];
        for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], H[j * d + l], s);
        S[i * D + j] = s;
        S[j * D + i] = s;
    }
    GD_FOR(i, D) {
        double s = 0.0, s0 = 0.0;
        for (int l = 0; l < d; ++l) {
            s = fma(H[i * d + l], mp[l], s);
            if (first) s0 = fma(H[i * d + l], m0[l], s0);
        }
        V[i * (d + 1) + d] = y[i] - s;            // innovation seen by eta: y - H A b
        e[i] = first ? (y[i] - s0) : (y[i] - s);  // innovation seen by b on the first step: y - H m0
    }
    GD_FOR(idx, D * d) V[(idx / d) * (d + 1) + idx % d] = HPhi[idx];
    gd_sync();
    gd_chol(w, S, D);
    gd_copy(w, Kt, HP, D * d);
    gd_chol_solve(w, S, Kt, D, d);
    gd_chol_solve(w, S, V, D, d + 1);
    GD_FOR(i, d) {
        double s = g.eta[i];
        for (int a = 0; a < D; ++a) s = fma(HPhi[a * d + i], V[a * (d + 1) + d], s);
        g.eta[i] = s;
        double bb = first ? m0[i] : mp[i];
        for (int a = 0; a < D; ++a) bb = fma(Kt[a * d + i], e[a], bb);
        g.b[i] = bb;
    }
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        double t = Phi[idx];
        for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], HPhi[a * d + j], t);
        g.A[idx] = t;
        if (j <= i) {
            double u = g.J[idx], c = Pp[idx];
            for (int a = 0; a < D; ++a) {
                u = fma(HPhi[a * d + i], V[a * (d + 1) + j], u);
                c = fma(-Kt[a * d + i], HP[a * d + j], c);
            }
            g.J[i * d + j] = u; g.J[j * d + i] = u;
            g.C[i * d + j] = c; g.C[j * d + i] = c;
        }
    }
    gd_sync();
}

// B (d x c) <- (I + C J)^-1 B
BN_DEV void gd_solve_ipcj(GdW w, int d, const double* C, const double* J, double* B, int c, double* M) {
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        double s = (i == j) ? 1.0 : 0.0;
        for (int l = 0; l < d; ++l) s = fma(C[i * d + l], J[l * d + j], s);
        M[idx] = s;
    }
    gd_sync();
    gd_lu_solve(w, M, B, d, c);
}

// out = op(e1 = earlier, e2 = later)   (ops.py:203-219).  out may alias e1 or e2.
BN_DEV void gd_filter_combine(GdW w, int d, GdFElem e1, GdFElem e2, GdFElem out, GdPool pool) {
    const int c = 2 * d + 1;
    double* B = pool.take(d * c);      // [A1 | C1 | b1 + C1 eta2] -> T [...]
    double* M = pool.take(d * d);
    double* TA1 = pool.take(d * d);
    double* W = pool.take(d * d);
    double* tb = pool.take(d);
    double* u = pool.take(d);
    double* JA = pool.take(d * d);
    double* rA = pool.take(d * d);
    double* rb = pool.take(d);
    double* rC = pool.take(d * d);
    double* rJ = pool.take(d * d);
    double* reta = pool.take(d);
    GD_FOR(i, d) {
        double s = e1.b[i];
        for (int l = 0; l < d; ++l) s = fma(e1.C[i * d + l], e2.eta[l], s);
        B[i * c + 2 * d] = s;
        for (int j = 0; j < d; ++j) {
            B[i * c + j] = e1.A[i * d + j];
            B[i * c + d + j] = e1.C[i * d + j];
        }
    }
    gd_sync();
    gd_solve_ipcj(w, d, e1.C, e2.J, B, c, M);
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        TA1[idx] = B[i * c + j];
        W[idx] = 0.5 * (B[i * c + d + j] + B[j * c + d + i]);
    }
    GD_FOR(i, d) {
        tb[i] = B[i * c + 2 * d];
        double s = e2.eta[i];
        for (int l = 0; l < d; ++l) s = fma(-e2.J[i * d + l], e1.b[l], s);
        u[i] = s;
    }
    gd_sync();
    gd_mm(w, rA, e2.A, TA1, d, d, d, false, false);
    gd_mm(w, rb, e2.A, tb, d, d, 1, false, false, e2.b);
    gd_mm(w, M, e2.A, W, d, d, d, false, false);
    gd_mm_sym(w, rC, M, e2.A, d, d, true, e2.C);
    gd_mm(w, reta, TA1, u, d, d, 1, true, false, e1.eta);
    gd_mm(w, JA, e2.J, e1.A, d, d, d, false, false);
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (j > i) continue;
        double s = e1.J[idx];
        for (int l = 0; l < d; ++l) s = fma(TA1[l * d + i], JA[l * d + j], s);
        rJ[i * d + j] = s;
        rJ[j * d + i] = s;
    }
    gd_sync();
    gd_copy(w, out.A, rA, d * d);
    gd_copy(w, out.b, rb, d);
    gd_copy(w, out.C, rC, d * d);
    gd_copy(w, out.J, rJ, d * d);
    gd_copy(w, out.eta, reta, d);
}

// (m, P) <- the state after the block e, given the state (m, P) before it
BN_DEV void gd_filter_apply(GdW w, int d, GdFElem e, double* m, double* P, GdPool pool) {
    const int c = d + 1;
    double* B = pool.take(d * c);  // [P | m + P eta]
    double* M = pool.take(d * d);
    double* W = pool.take(d * d);
    double* tb = pool.take(d);
    double* T1 = pool.take(d * d);
    GD_FOR(i, d) {
        double s = m[i];
        for (int l = 0; l < d; ++l) s = fma(P[i * d + l], e.eta[l], s);
        B[i * c + d] = s;
        for (int j = 0; j < d; ++j) B[i * c + j] = P[i * d + j];
    }
    gd_sync();
    gd_solve_ipcj(w, d, P, e.J, B, c, M);
    GD_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        W[idx] = 0.5 * (B[i * c + j] + B[j * c + i]);
    }
    GD_FOR(i, d) tb[i] = B[i * c + d];
    gd_sync();
    gd_mm(w, m, e.A, tb, d, d, 1, false, false, e.b);
    gd_mm(w, T1, e.A, W, d, d, d, false, false);
    gd_mm_sym(w, P, T1, e.A, d, d, true, e.C);
}

// ---------------------------------------------------------------------------------------------- smoother
// G = (pP^-1 A fP)^T, pm = A fm, pP = A fP A^T + Q   (ops.py:294-299)
BN_DEV void gd_rts_gain(GdW w, int d, const double* fm, const double* fP, const double* A, const double* Q, double* G,
                        double* pm, double* pP, GdPool pool) {
    double* AfP = pool.take(d * d);
    double* Lc = pool.take(d * d);
    gd_mm(w, pm, A, fm, d, d, 1, false, false);
    gd_mm(w, AfP, A, fP, d, d, d, false, false);
    gd_mm_sym(w, pP, AfP, A, d, d, true, Q);
    gd_copy(w, Lc, pP, d * d);
    gd_chol(w, Lc, d);
    gd_chol_solve(w, Lc, AfP, d, d);
    GD_FOR(idx, d * d) G[idx] = AfP[(idx % d) * d + idx / d];
    gd_sync();
}

// (sm, sP) <- fm + G (sm - pm), fP + G (sP - pP) G^T   (ops.py:300-301)
BN_DEV void gd_rts_step(GdW w, int d, double* sm, double* sP, const double* fm, const double* fP, const double* G,
                        const double* pm, const double* pP, GdPool pool) {
    double* dm = pool.take(d);
    double* dP = pool.take(d * d);
    double* T1 = pool.take(d * d);
    GD_FOR(i, d) dm[i] = sm[i] - pm[i];
    GD_FOR(i, d * d) dP[i] = sP[i] - pP[i];
    gd_sync();
    gd_mm(w, sm, G, dm, d, d, 1, false, false, fm);
    gd_mm(w, T1, G, dP, d, d, d, false, false);
    gd_mm_sym(w, sP, T1, G, d, d, true, fP);
}

// smoothing element [E | g | L] = 2 d^2 + d doubles
BN_DEV constexpr int gd_selem(int d) { return 2 * d * d + d; }
struct GdSElem {
    double *E, *g, *L;
    BN_DEV GdSElem(double* p, int d) : E(p), g(p + d * d), L(p + d * d + d) {}
};

// element of one step (ops.py:318-325): E = G, g = fm - E A fm, L = fP - E pP E^T
BN_DEV void gd_rts_element(GdW w, int d, const double* fm, const double* fP, const double* A, const double* Q, GdSElem e,
                           GdPool pool) {
    double* pm = pool.take(d);
    double* pP = pool.take(d * d);
    double* X = pool.take(d * d);
    gd_rts_gain(w, d, fm, fP, A, Q, e.E, pm, pP, pool);
    gd_mm(w, e.g, e.E, pm, d, d, 1, false, false, fm, -1.0);
    gd_mm(w, X, e.E, pP, d, d, d, false, false);
    gd_mm_sym(w, e.L, X, e.E, d, d, true, fP, -1.0);
}

// out = combine(e1 = the LATER part already accumulated, e2 = the EARLIER element)  (ops.py:328-335); out may alias
BN_DEV void gd_smoother_combine(GdW w, int d, GdSElem e1, GdSElem e2, GdSElem out, GdPool pool) {
    double* rE = pool.take(d * d);
    double* rg = pool.take(d);
    double* rL = pool.take(d * d);
    double* T1 = pool.take(d * d);
    gd_mm(w, rE, e2.E, e1.E, d, d, d, false, false);
    gd_mm(w, rg, e2.E, e1.g, d, d, 1, false, false, e2.g);
    gd_mm(w, T1, e2.E, e1.L, d, d, d, false, false);
    gd_mm_sym(w, rL, T1, e2.E, d, d, true, e2.L);
    gd_copy(w, out.E, rE, d * d);
    gd_copy(w, out.g, rg, d);
    gd_copy(w, out.L, rL, d * d);
}

BN_DEV void gd_smoother_apply(GdW w, int d, GdSElem e, double* m, double* P, GdPool pool) {
    double* tm = pool.take(d);
    double* T1 = pool.take(d * d);
    gd_copy(w, tm, m, d);
    gd_mm(w, m, e.E, tm, d, d, 1, false, false, e.g);
    gd_mm(w, T1, e.E, P, d, d, d, false, false);
    gd_mm_sym(w, P, T1, e.E, d, d, true, e.L);
}

// ---------------------------------------------------------------------------------------------- problem description
struct GdKf {
    long long N;
    int d, D;
    const double *As, *Qs, *H, *ys, *Rs, *m0, *P0;
    const unsigned char* masks;
    int return_predict;
    double *fms, *fPs;   // nullable together
};

struct GdRts {
    long long N;
    int d, Df;
    const double *fms, *fPs, *As, *Qs, *H;
    int return_full;
    double *sms, *sPs, *gains;
};

struct GdPlan { int L; long long nchunks; };
// one resident wave of warps: as many CTAs of kGdWarps warps per SM as their shared-memory pools allow (32 warps per SM at
// d <= 6, 16 at d = 8, 4 at d = 16)
inline long long gd_max_chunks(int d) {
    const long long warp_bytes = (long long)gd_pool_doubles(d) * 8;
    long long ctas = (220 * 1024) / (warp_bytes * kGdWarps);
    if (ctas < 1) ctas = 1;
    if (ctas > 8) ctas = 8;
    static const long long env_cap = [] {  // BN_B200_GD_CTAS=<n>: cap on resident CTAs per SM (tuning aid)
        const char* e = getenv("BN_B200_GD_CTAS");
        return e ? atoll(e) : 0LL;
    }();
    if (env_cap > 0 && ctas > env_cap) ctas = env_cap;
    return 148LL * ctas * kGdWarps;
}
inline GdPlan gd_plan(long long N, int d) {
    const long long maxc = gd_max_chunks(d);
    long long L = (N + maxc - 1) / maxc;
    if (L < kGdMinChunk) L = kGdMinChunk;
    GdPlan p;
    p.L = (int)L;
    p.nchunks = (N + L - 1) / L;
    return p;
}

BN_DEV void gd_write_state(GdW w, const GdKf& a, long long k, const double* m, const double* P) {
    if (!a.fms) return;
    GD_FOR(i, a.d) a.fms[k * a.d + i] = m[i];
    GD_FOR(i, a.d * a.d) a.fPs[k * (a.d * a.d) + i] = P[i];
}

// filter over steps [k0, k1) from the state (m, P): the body of the sequential form and of phase 3
BN_DEV double gd_kf_run(GdW w, const GdKf& a, long long k0, long long k1, double* m, double* P, bool want_ell, GdPool pool) {
    const int d = a.d, D = a.D;
    double* mp = pool.take(d);
    double* Pp = pool.take(d * d);
    double ell = 0.0;
    for (long long k = k0; k < k1; ++k) {
        const unsigned char* mk = a.masks ? a.masks + k * D : nullptr;
        ell += gd_kf_step(w, d, D, m, P, a.As + k * d * d, a.Qs + k * d * d, a.H, a.ys + k * D, a.Rs + k * D * D, mk, mp, Pp,
                          pool, want_ell);
        if (a.return_predict) gd_write_state(w, a, k, mp, Pp);
        else gd_write_state(w, a, k, m, P);
        gd_sync();
    }
    return ell;
}

// phase 1: chunk c -> its filtering element in agg[c]
BN_DEV void gd_kf_reduce_chunk(GdW w, const GdKf& a, int L, long long c, double* agg, GdPool pool) {
    const int d = a.d, D = a.D;
    GdFElem g(pool.take(gd_felem(d)), d);
    gd_felem_identity(w, g, d);
    const long long k0 = c * L, k1 = (k0 + L < a.N) ? k0 + L : a.N;
    for (long long k = k0; k < k1; ++k) {
        const bool first = (k == 0);
        gd_filter_absorb(w, d, D, g, a.As + k * d * d, first ? a.P0 : a.Qs + k * d * d, a.H, a.ys + k * D, a.Rs + k * D * D,
                         first, a.m0, pool);
    }
    gd_copy(w, agg + c * gd_felem(d), g.A, gd_felem(d));
}

// phase 3: chunk c filtered from its incoming state.  The first step of the scan form starts its update from (m0, P0)
// itself (ops.py:222-229, 245-248); its log-likelihood / predicted outputs use A_0 m0, A_0 P0 A_0^T + Q_0.
BN_DEV double gd_kf_apply_chunk(GdW w, const GdKf& a, int L, long long c, const double* prefix, bool want_ell, GdPool pool) {
    const int d = a.d, D = a.D;
    double* m = pool.take(d);
    double* P = pool.take(d * d);
    long long k0 = c * L;
    const long long k1 = (k0 + L < a.N) ? k0 + L : a.N;
    double ell = 0.0;
    if (c == 0) {
        double* mt = pool.take(d);
        double* Pt = pool.take(d * d);
        double* mp = pool.take(d);
        double* Pp = pool.take(d * d);
        double* I = pool.take(d * d);
        double* Z = pool.take(d * d);
        gd_copy(w, mt, a.m0, d);
        gd_copy(w, Pt, a.P0, d * d);
        const unsigned char* mk = a.masks ? a.masks : nullptr;
        if (want_ell || a.return_predict)
            ell += gd_kf_step(w, d, D, mt, Pt, a.As, a.Qs, a.H, a.ys, a.Rs, mk, mp, Pp, pool, want_ell);
        if (a.return_predict) gd_write_state(w, a, 0, mp, Pp);
        gd_sync();
        gd_copy(w, m, a.m0, d);
        gd_copy(w, P, a.P0, d * d);
        gd_eye(w, I, d);
        gd_fill(w, Z, 0.0, d * d);
        gd_kf_step(w, d, D, m, P, I, Z, a.H, a.ys, a.Rs, mk, mt, Pt, pool, false);
        if (!a.return_predict) gd_write_state(w, a, 0, m, P);
        gd_sync();
        k0 = 1;
    } else {
        gd_fill(w, m, 0.0, d);
        gd_fill(w, P, 0.0, d * d);
        gd_filter_apply(w, d, GdFElem(const_cast<double*>(prefix) + (c - 1) * gd_felem(d), d), m, P, pool);
    }
    return ell + gd_kf_run(w, a, k0, k1, m, P, want_ell, pool);
}

// ---- smoother
BN_DEV void gd_write_smoothed(GdW w, const GdRts& a, long long k, const double* sm, const double* sP, const double* G,
                              GdPool pool) {
    const int d = a.d, Df = a.Df;
    if (a.return_full) {
        GD_FOR(i, d) a.sms[k * d + i] = sm[i];
        GD_FOR(i, d * d) a.sPs[k * (d * d) + i] = sP[i];
    } else {
        double* HP = pool.take(Df * d);
        gd_mm(w, HP, a.H, sP, Df, d, d, false, false);
        GD_FOR(i, Df) {
            double s = 0.0;
            for (int l = 0; l < d; ++l) s = fma(a.H[i * d + l], sm[l], s);
            a.sms[k * Df + i] = s;
        }
        GD_FOR(idx, Df * Df) {
            const int i = idx / Df, j = idx % Df;
            double s = 0.0;
            for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], a.H[j * d + l], s);
            a.sPs[k * (Df * Df) + idx] = s;
        }
    }
    if (a.gains) GD_FOR(i, d * d) a.gains[k * (d * d) + i] = G[i];
    gd_sync();
}

// RTS recursion over steps k1-1 .. k0 from the smoothed state (sm, sP) of step k1 (or, with `terminal`, starting AT the
// last step of the series, whose smoothed state is its filtered state)
BN_DEV void gd_rts_run(GdW w, const GdRts& a, long long k0, long long k1, double* sm, double* sP, bool terminal, GdPool pool) {
    const int d = a.d;
    double* G = pool.take(d * d);
    double* pm = pool.take(d);
    double* pP = pool.take(d * d);
    for (long long k = k1 - 1; k >= k0; --k) {
        const double* fm = a.fms + k * d;
        const double* fP = a.fPs + k * d * d;
        gd_rts_gain(w, d, fm, fP, a.As + k * d * d, a.Qs + k * d * d, G, pm, pP, pool);
        if (terminal && k == a.N - 1) {
            gd_copy(w, sm, fm, d);
            gd_copy(w, sP, fP, d * d);
        } else {
            gd_rts_step(w, d, sm, sP, fm, fP, G, pm, pP, pool);
        }
        gd_write_smoothed(w, a, k, sm, sP, G, pool);
    }
}

// phase 1: chunk c -> its smoothing element at scan position nchunks - 1 - c
BN_DEV void gd_rts_reduce_chunk(GdW w, const GdRts& a, int L, long long nchunks, long long c, double* agg, GdPool pool) {
    const int d = a.d, ne = gd_selem(d);
    GdSElem acc(pool.take(ne), d), e(pool.take(ne), d);
    gd_eye(w, acc.E, d);
    gd_fill(w, acc.g, 0.0, d + d * d);
    const long long k0 = c * L, k1 = (k0 + L < a.N) ? k0 + L : a.N;
    for (long long k = k1 - 1; k >= k0; --k) {
        const double* fm = a.fms + k * d;
        const double* fP = a.fPs + k * d * d;
        if (k == a.N - 1) {  // last_parallel_smoothing_element, ops.py:314-315
            gd_fill(w, e.E, 0.0, d * d);
            gd_copy(w, e.g, fm, d);
            gd_copy(w, e.L, fP, d * d);
        } else {
            gd_rts_element(w, d, fm, fP, a.As + k * d * d, a.Qs + k * d * d, e, pool);
        }
        gd_smoother_combine(w, d, acc, e, acc, pool);
    }
    gd_copy(w, agg + (nchunks - 1 - c) * ne, acc.E, ne);
}

BN_DEV void gd_rts_apply_chunk(GdW w, const GdRts& a, int L, long long nchunks, long long c, const double* prefix, GdPool pool) {
    const int d = a.d;
    double* sm = pool.take(d);
    double* sP = pool.take(d * d);
    const long long p = nchunks - 1 - c;
    gd_fill(w, sm, 0.0, d);
    gd_fill(w, sP, 0.0, d * d);
    if (p > 0) gd_smoother_apply(w, d, GdSElem(const_cast<double*>(prefix) + (p - 1) * gd_selem(d), d), sm, sP, pool);
    const long long k0 = c * L, k1 = (k0 + L < a.N) ? k0 + L : a.N;
    gd_rts_run(w, a, k0, k1, sm, sP, true, pool);
}


// ---------------------------------------------------------------------------------------------- phase 2 in groups
// The scan over the chunk elements, organised for warps: a level cuts its n elements into groups of kGdScanGroup
// consecutive ones; ONE WARP walks a group in order (31 warp-cooperative combines), leaving within-group inclusive prefixes
// and the group total; the totals are the next level (n / 32 elements) and so on until one group is left; going back down,
// every element of a group > 0 takes one combine with the prefix of the groups before it.  1 184 elements -> 37 -> 2 -> 1:
// about 100 combine latencies on the critical path instead of 1 184.
constexpr int kGdScanGroup = 32;

template <bool FILTER>
BN_DEV constexpr int gd_elem(int d) { return FILTER ? gd_felem(d) : gd_selem(d); }

template <bool FILTER>
BN_DEV void gd_combine(GdW w, int d, double* e1, double* e2, double* out, GdPool pool) {
    if constexpr (FILTER) gd_filter_combine(w, d, GdFElem(e1, d), GdFElem(e2, d), GdFElem(out, d), pool);
    else gd_smoother_combine(w, d, GdSElem(e1, d), GdSElem(e2, d), GdSElem(out, d), pool);
}

// group g of a level: in[n] -> prefix[n] (within-group inclusive), totals[g]
template <bool FILTER>
BN_DEV void gd_scan_group(GdW w, int d, long long n, const double* in, double* prefix, double* totals, long long g, GdPool pool) {
    const int ne = gd_elem<FILTER>(d);
    double* acc = pool.take(ne);
    double* cur = pool.take(ne);
    const long long i0 = g * kGdScanGroup, i1 = (i0 + kGdScanGroup < n) ? i0 + kGdScanGroup : n;
    gd_copy(w, acc, in + i0 * ne, ne);
    gd_copy(w, prefix + i0 * ne, acc, ne);
    for (long long i = i0 + 1; i < i1; ++i) {
        gd_copy(w, cur, in + i * ne, ne);
        gd_combine<FILTER>(w, d, acc, cur, acc, pool);
        gd_copy(w, prefix + i * ne, acc, ne);
    }
    if (totals) gd_copy(w, totals + g * ne, acc, ne);
}

// element i of a level, after the level above has been scanned: prefix[i] <- combine(upper[group(i) - 1], prefix[i])
template <bool FILTER>
BN_DEV void gd_scan_down(GdW w, int d, double* prefix, const double* upper, long long i, GdPool pool) {
    const long long g = i / kGdScanGroup;
    if (g == 0) return;
    const int ne = gd_elem<FILTER>(d);
    double* a = pool.take(ne);
    double* b = pool.take(ne);
    gd_copy(w, a, upper + (g - 1) * ne, ne);
    gd_copy(w, b, prefix + i * ne, ne);
    gd_combine<FILTER>(w, d, a, b, b, pool);
    gd_copy(w, prefix + i * ne, b, ne);
}

// elements of scratch the upper levels need (totals of every level above the first)
inline long long gd_scan_upper_elems(long long n) {
    long long tot = 0;
    while (n > kGdScanGroup) {
        n = (n + kGdScanGroup - 1) / kGdScanGroup;
        tot += n;
    }
    return tot + 1;
}

}  // namespace bn
