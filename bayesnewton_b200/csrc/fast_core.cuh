// Per-step arithmetic of the fused posterior update (filter + smoother) for stationary Matern
// stacks, written for the fp64 pipe: the measurement model is a 0/1 selector (kernels.py:212-214,
// 269-271, 1535-1541), A / Pinf / Q are block diagonal (kernels.py:1543-1583), Pinf has the
// checkerboard sparsity of every Matern family (kernels.py:207-210, 288-293, 367-382), and the
// predicted covariance is formed as  Pinf + A (P - Pinf) A^T  (== A P A^T + Q with
// Q = Pinf - A Pinf A^T, ops.py:149-151,159) so Q is never built in the filter.
// Same recursions as core.cuh (ops.py:156-175, 183-219, 290-335); only the order of rounding differs.
#pragma once
#include "core.cuh"
#include "gen.cuh"
#include "dual.cuh"

namespace BN_NS {

// -------------------------------------------------------------------------------- prepared generator
template <int FAMILY, int NC_>
struct FastGen {
    static constexpr int family = FAMILY;
    static constexpr int n = FamilyDim<FAMILY>::value;
    static constexpr int NC = NC_;
    static constexpr int d = NC * n;
    static constexpr int D = NC;
    static constexpr int kBlockA = NC * n * n;
    static constexpr int kBlockS = NC * symn(n);
    static __host__ __device__ constexpr int sel(int a) { return a * n; }

    real lam[NC];          // sqrt(2 nu) / lengthscale
    real Pb[kBlockS];      // Pinf blocks, packed

    BN_DEV void prepare(const bn_kernel_spec& s) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            lam[c] = MaternBlock<FAMILY, real>::rate(s.lengthscale[c]);
            MaternBlock<FAMILY, real>::pinf(s.variance[c], s.lengthscale[c], Pb + c * symn(n));
        }
    }
    // A blocks for a step of length h
    BN_DEV void trans(real h, real* Ab) const {
#pragma unroll
        for (int c = 0; c < NC; ++c) MaternBlock<FAMILY, real>::transition_rate(lam[c], h, Ab + c * n * n);
    }
    // Q blocks = Pinf - A Pinf A^T, skipping the structural zeros of Pinf ((i + j) odd)
    BN_DEV void noise(const real* Ab, real* Qb) const {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const real* A = Ab + c * n * n;
            const real* P = Pb + c * symn(n);
            real X[n * n];
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real s = 0.0;
                    bool started = false;
#pragma unroll
                    for (int l = 0; l < n; ++l)
                        if (((l + j) & 1) == 0) {
                            s = started ? fma(A[i * n + l], P[sidx(l, j)], s) : A[i * n + l] * P[sidx(l, j)];
                            started = true;
                        }
                    X[i * n + j] = s;
                }
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    real s = (((i + j) & 1) == 0) ? P[sidx(i, j)] : 0.0;
#pragma unroll
                    for (int l = 0; l < n; ++l) s = fma(-X[i * n + l], A[j * n + l], s);
                    Qb[c * symn(n) + sidx(i, j)] = s;
                }
        }
    }
    // A blocks and their derivative with respect to the rate lam_c (dual-number evaluation of the same closed form)
    BN_DEV void trans_d(real h, real* Ab, real* dAb) const {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            Dual A[n * n];
            MaternBlock<FAMILY, Dual>::transition_rate(Dual(lam[c], 1.0), Dual(h), A);
#pragma unroll
            for (int i = 0; i < n * n; ++i) {
                Ab[c * n * n + i] = A[i].v;
                dAb[c * n * n + i] = A[i].d;
            }
        }
    }
    // full packed Pinf (d x d)
    BN_DEV void pinf_full(real* P) const {
#pragma unroll
        for (int i = 0; i < symn(d); ++i) P[i] = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(c * n + i, c * n + j)] = Pb[c * symn(n) + sidx(i, j)];
    }
};

// -------------------------------------------------------------------------------- block-diagonal products
// y = A x, A = blockdiag(Ab)
template <class G>
BN_DEV void bd_matvec(const real* Ab, const real* x, real* y) {
    constexpr int n = G::n;
#pragma unroll
    for (int c = 0; c < G::NC; ++c)
#pragma unroll
        for (int i = 0; i < n; ++i) {
            real s = 0.0;
#pragma unroll
            for (int l = 0; l < n; ++l) s = fma(Ab[c * n * n + i * n + l], x[c * n + l], s);
            y[c * n + i] = s;
        }
}

// X (d x d full) = A S, S symmetric packed
template <class G>
BN_DEV void bd_mat_sym(const real* Ab, const real* S, real* X) {
    constexpr int n = G::n, d = G::d;
#pragma unroll
    for (int c = 0; c < G::NC; ++c)
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) {
                real s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(Ab[c * n * n + i * n + l], S[sidx(c * n + l, j)], s);
                X[(c * n + i) * d + j] = s;
            }
}

// C (d x cc full) = A B, B (d x cc full)
template <class G, int cc>
BN_DEV void bd_matmul(const real* Ab, const real* B, real* C) {
    constexpr int n = G::n;
#pragma unroll
    for (int c = 0; c < G::NC; ++c)
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < cc; ++j) {
                real s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(Ab[c * n * n + i * n + l], B[(c * n + l) * cc + j], s);
                C[(c * n + i) * cc + j] = s;
            }
}

// out (packed, lower) = X A^T + blockdiag(Sb)   (Sb nullable)
template <class G>
BN_DEV void bd_abt_sym(const real* X, const real* Ab, const real* Sb, real* out) {
    constexpr int n = G::n, d = G::d;
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const int cj = j / n, jj = j % n, ci = i / n, ii = i % n;
            real s = (Sb && ci == cj) ? Sb[cj * symn(n) + sidx(ii, jj)] : 0.0;
#pragma unroll
            for (int l = 0; l < n; ++l) s = fma(X[i * d + cj * n + l], Ab[cj * n * n + jj * n + l], s);
            out[sidx(i, j)] = s;
        }
}

// Pp = Pinf + A (P - Pinf) A^T
template <class G>
BN_DEV void predict_cov(const G& g, const real* Ab, const real* P, real* Pp) {
    constexpr int n = G::n, d = G::d;
    real Dl[symn(d)];
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool in_block = (i / n == j / n) && ((((i % n) + (j % n)) & 1) == 0);
            Dl[sidx(i, j)] = in_block ? P[sidx(i, j)] - g.Pb[(i / n) * symn(n) + sidx(i % n, j % n)] : P[sidx(i, j)];
        }
    real X[d * d];
    bd_mat_sym<G>(Ab, Dl, X);
    // + Pinf, skipping its structural zeros
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const int cj = j / n, jj = j % n, ci = i / n, ii = i % n;
            const bool nz = (ci == cj) && (((ii + jj) & 1) == 0);
            real s = nz ? g.Pb[cj * symn(n) + sidx(ii, jj)] : 0.0;
            bool started = nz;
#pragma unroll
            for (int l = 0; l < n; ++l) {
                s = started ? fma(X[i * d + cj * n + l], Ab[cj * n * n + jj * n + l], s)
                            : X[i * d + cj * n + l] * Ab[cj * n * n + jj * n + l];
                started = true;
            }
            Pp[sidx(i, j)] = s;
        }
}

// a Cholesky pivot that is not positive poisons the result exactly as sqrt() of it would (utils.py:14-19)
BN_DEV real pd_guard(real s) { return s > 0.0 ? s : nan(""); }

// -------------------------------------------------------------------------------- filter step
// Innovation covariance S (packed D), innovation e, HP rows -- shared by the step and the absorb.
template <class G>
BN_DEV void innovation(const real* mp, const real* Pp, const real* y, const real* R, real* S, real* e) {
    constexpr int D = G::D;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        e[a] = y[a] - mp[G::sel(a)];
#pragma unroll
        for (int b = 0; b <= a; ++b) S[sidx(a, b)] = Pp[sidx(G::sel(a), G::sel(b))] + R[a * D + b];
    }
}

// Kt (D x d) = S^-1 HP, with HP[a][:] = Pp[sel(a)][:]; returns through Kt.  Sf receives what the
// caller needs to apply S^-1 again: D == 1: Sf[0] = 1/S; D > 1: the Cholesky factor.
template <class G>
BN_DEV void gain(const real* Pp, const real* S, real* Sf, real* Kt) {
    constexpr int d = G::d, D = G::D;
    if constexpr (D == 1) {
        const real r = 1.0 / pd_guard(S[0]);
        Sf[0] = r;
#pragma unroll
        for (int i = 0; i < d; ++i) Kt[i] = Pp[sidx(G::sel(0), i)] * r;
    } else {
#pragma unroll
        for (int i = 0; i < symn(D); ++i) Sf[i] = S[i];
        chol<D>(Sf);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int i = 0; i < d; ++i) Kt[a * d + i] = Pp[sidx(G::sel(a), i)];
        chol_solve<D, d>(Sf, Kt);
    }
}

template <class G>
BN_DEV real step_logpdf(const real* S, const real* Sf, const real* e, const unsigned char* msk) {
    constexpr int D = G::D;
    if constexpr (D == 1) {
        if (msk && msk[0]) return 0.0;  // the masked 1x1 density is exactly 1 (utils.py:376-396)
        return -0.5 * (e[0] * e[0] * Sf[0] + kLog2Pi + log(S[0]));
    } else {
        return mvn_logpdf_masked<D>(S, e, msk);
    }
}

// One predict + update (ops.py:156-175).  (m, P) in: filtered state of the previous step; out:
// of this step.  mp / Pp receive the prediction.  Returns the log-likelihood increment.
template <class G, bool WANT_ELL>
BN_DEV real fkf_step(const G& g, real* m, real* P, const real* Ab, const real* y, const real* R,
                       const unsigned char* msk, real* mp, real* Pp) {
    constexpr int d = G::d, D = G::D;
    bd_matvec<G>(Ab, m, mp);
    predict_cov<G>(g, Ab, P, Pp);
    real S[symn(D)], Sf[symn(D)], e[D], Kt[D * d];
    innovation<G>(mp, Pp, y, R, S, e);
    gain<G>(Pp, S, Sf, Kt);
    real ell = 0.0;
    if constexpr (WANT_ELL) ell = step_logpdf<G>(S, Sf, e, msk);
#pragma unroll
    for (int i = 0; i < d; ++i) {
        real s = mp[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(Kt[a * d + i], e[a], s);
        m[i] = s;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real t = Pp[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], Pp[sidx(G::sel(a), j)], t);
            P[sidx(i, j)] = t;
        }
    }
    return ell;
}

// Fold one step into a chunk aggregate (A, b, C, J, eta) -- ops.py:183-219 evaluated as a
// zero-prior Kalman step that also tracks the sensitivity to, and the information about, the
// state entering the chunk.  `first`: global step 0 of the scan form (Q_0 := P_0, m0 = 0;
// ops.py:222-229), valid only on a fresh aggregate.
template <class G>
BN_DEV void fkf_absorb(const G& g, typename FilterAlg<G::d>::Elem& el, const real* Ab, const real* y,
                       const real* R, bool first) {
    constexpr int d = G::d, D = G::D, n = G::n;
    real mp[d], Pp[symn(d)], Phi[d * d];
    if (first) {
#pragma unroll
        for (int i = 0; i < d; ++i) mp[i] = 0.0;
        g.pinf_full(Pp);
#pragma unroll
        for (int i = 0; i < d * d; ++i) Phi[i] = 0.0;
#pragma unroll
        for (int c = 0; c < G::NC; ++c)
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j < n; ++j) Phi[(c * n + i) * d + c * n + j] = Ab[c * n * n + i * n + j];
    } else {
        bd_matvec<G>(Ab, el.b, mp);
        predict_cov<G>(g, Ab, el.C, Pp);
        bd_matmul<G, d>(Ab, el.A, Phi);
    }
    real S[symn(D)], Sf[symn(D)], e[D], Kt[D * d];
    innovation<G>(mp, Pp, y, R, S, e);
    gain<G>(Pp, S, Sf, Kt);
    // V = S^-1 [H Phi | e]  (D x (d+1)), H Phi = selected rows of Phi
    real V[D * (d + 1)];
    if constexpr (D == 1) {
#pragma unroll
        for (int j = 0; j < d; ++j) V[j] = Phi[G::sel(0) * d + j] * Sf[0];
        V[d] = e[0] * Sf[0];
    } else {
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
            for (int j = 0; j < d; ++j) V[a * (d + 1) + j] = Phi[G::sel(a) * d + j];
            V[a * (d + 1) + d] = e[a];
        }
        chol_solve<D, d + 1>(Sf, V);
    }
#pragma unroll
    for (int i = 0; i < d; ++i) {
        real s = el.eta[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(Phi[G::sel(a) * d + i], V[a * (d + 1) + d], s);
        el.eta[i] = s;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real t = el.J[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(Phi[G::sel(a) * d + i], V[a * (d + 1) + j], t);
            el.J[sidx(i, j)] = t;
        }
    }
#pragma unroll
    for (int i = 0; i < d; ++i) {
        real s = mp[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(Kt[a * d + i], e[a], s);
        el.b[i] = s;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            real t = Phi[i * d + j];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], Phi[G::sel(a) * d + j], t);
            el.A[i * d + j] = t;
        }
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real t = Pp[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], Pp[sidx(G::sel(a), j)], t);
            el.C[sidx(i, j)] = t;
        }
    }
}

// -------------------------------------------------------------------------------- LDL^T
// S (packed, destroyed) -> unit lower factor in the strict lower triangle, 1/d_i on the diagonal.
template <int n>
BN_DEV void ldlt(real* S) {
#pragma unroll
    for (int j = 0; j < n; ++j) {
        real w[n];  // w[k] = L[j][k] d_k
        real dj = S[sidx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) {
            w[k] = S[sidx(j, k)];                 // still holds L[j][k] d_k (scaled below)
            dj = fma(-w[k], w[k] * S[sidx(k, k)], dj);
        }
        const real inv = 1.0 / pd_guard(dj);
        // rows below: S[i][j] <- (S[i][j] - sum_k (L[i][k] d_k) L[j][k])   kept as L[i][j] d_j
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            real t = S[sidx(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-S[sidx(i, k)], w[k] * S[sidx(k, k)], t);
            S[sidx(i, j)] = t;
        }
        S[sidx(j, j)] = inv;
    }
    // here S[i][j] (i > j) = L[i][j] d_j and S[j][j] = 1 / d_j
}

// solve (L D L^T) X = B in place, B (n x c) row-major; S as left by ldlt()
template <int n, int c>
BN_DEV void ldlt_solve(const real* S, real* B) {
#pragma unroll
    for (int j = 0; j < c; ++j) {
        // forward: z = L^-1 b with L[i][k] = S[i][k] * S[k][k]
        real z[n];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            real s = B[i * c + j];
#pragma unroll
            for (int k = 0; k < i; ++k) s = fma(-S[sidx(i, k)], z[k], s);   // z[k] already scaled by 1/d_k
            z[i] = s * S[sidx(i, i)];                                         // z_i / d_i
        }
        // backward: x = L^-T (D^-1 z);  L^T[i][k] = L[k][i] = S[k][i] * S[i][i]
#pragma unroll
        for (int i = n - 1; i >= 0; --i) {
            real s = 0.0;
#pragma unroll
            for (int k = i + 1; k < n; ++k) s = fma(S[sidx(k, i)], z[k], s);
            z[i] = fma(-s, S[sidx(i, i)], z[i]);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) B[i * c + j] = z[i];
    }
}

// -------------------------------------------------------------------------------- hyper-gradient of the filter log-likelihood
// d ell / d theta, ell = the filter log-likelihood compute_log_lik() returns (basemodels.py:726-741), the only
// route from the kernel hyper-parameters to energy() in a temporal model (sites and posterior are StateVars;
// the reference differentiates it with objax.GradValues, README.md:56-70).  The adjoint of the PREDICTED state of
// step k is available in closed form from the smoothed state the RTS sweep holds at that moment:
//     d ell / d m^-_k = v_k = (P^-_k)^-1 delta_k,                      delta_k = sm_k - m^-_k
//     d ell / d P^-_k = M_k = 1/2 (P^-_k)^-1 (sP_k - P^-_k + delta_k delta_k^T) (P^-_k)^-1
// (the prediction is the prior of the observations k..N-1 and the smoothed state their posterior).  With
// m^-_k = A_k m_{k-1},  P^-_k = A_k P_{k-1} A_k^T + Q_k,  Q_k = Pinf - A_k Pinf A_k^T  (ops.py:149-151):
//     d ell = < Gamma, dPinf > + sum_k < 2 M_k A_k (P_{k-1} - Pinf) + v_k m_{k-1}^T , dA_k >,
//     Gamma = sum_k (M_k - A_k^T M_k A_k)   (+ A_0^T M_0 A_0: the prior entering step 0 is Pinf itself).
// Everything on the right is in registers inside frts_step (the LDL^T factor of P^-, delta, sP - P^-), so the
// adjoint pass costs no HBM traffic at all.  A and Pinf are block diagonal: only the diagonal blocks of Gamma
// and of the dA coefficient are needed.  Not valid with a mask (the reference's mask rule removes the masked
// densities from ell but keeps their updates, so ell is no longer a marginal likelihood).
template <class G>
struct GradAcc {
    static constexpr int kFields = G::NC * (symn(G::n) + 1);
    real Gam[G::NC * symn(G::n)];  // diagonal blocks of Gamma, packed
    real gl[G::NC];                // sum_k < dA-coefficient block , dA_k / d lam_c >
    BN_DEV void zero() {
#pragma unroll
        for (int i = 0; i < G::NC * symn(G::n); ++i) Gam[i] = 0.0;
#pragma unroll
        for (int i = 0; i < G::NC; ++i) gl[i] = 0.0;
    }
};

// F: P^- factorised by ldlt(); dm = sm - m^-; dP = sP - P^- (packed); (pm_, pP_) = (m_{k-1}, P_{k-1}).
// first: the step whose incoming state is the stationary prior (Gamma takes the whole of M, no dA term).
template <class G>
BN_DEV void grad_accumulate(const G& g, const real* Ab, real h, const real* F, const real* dm,
                            const real* dP, const real* pm_, const real* pP_, bool first, GradAcc<G>& acc) {
    constexpr int d = G::d, n = G::n, c1 = d + 1;
    // [X | v] = (P^-)^-1 [dP + dm dm^T | dm]
    real B[d * c1];
#pragma unroll
    for (int i = 0; i < d; ++i) {
#pragma unroll
        for (int j = 0; j < d; ++j) B[i * c1 + j] = fma(dm[i], dm[j], dP[sidx(i, j)]);
        B[i * c1 + d] = dm[i];
    }
    ldlt_solve<d, c1>(F, B);
    // M2 = (P^-)^-1 X^T = 2 M
    real M2[d * d], v[d];
#pragma unroll
    for (int i = 0; i < d; ++i) {
        v[i] = B[i * c1 + d];
#pragma unroll
        for (int j = 0; j < d; ++j) M2[i * d + j] = B[j * c1 + i];
    }
    ldlt_solve<d, d>(F, M2);
    if (first) {
#pragma unroll
        for (int c = 0; c < G::NC; ++c)
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    acc.Gam[c * symn(n) + sidx(i, j)] += 0.25 * (M2[(c * n + i) * d + c * n + j] + M2[(c * n + j) * d + c * n + i]);
        return;
    }
    // XA = M2 A  (d x d)
    real XA[d * d];
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int c = 0; c < G::NC; ++c)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(M2[i * d + c * n + l], Ab[c * n * n + l * n + j], s);
                XA[i * d + c * n + j] = s;
            }
#pragma unroll
    for (int c = 0; c < G::NC; ++c) {
        // Gamma_c += 1/2 (M2_cc - A_c^T (M2 A)_cc)
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                real s = 0.5 * (M2[(c * n + i) * d + c * n + j] + M2[(c * n + j) * d + c * n + i]);
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(-Ab[c * n * n + l * n + i], XA[(c * n + l) * d + c * n + j], s);
                acc.Gam[c * symn(n) + sidx(i, j)] = fma(0.5, s, acc.Gam[c * symn(n) + sidx(i, j)]);
            }
        // < (M2 A (P_{k-1} - Pinf))_cc + v_c m_c^T , dA_c / d lam_c >; the derivative of the transition is formed
        // here, at its only use, by the dual-number evaluation of the closed form (nothing extra stays live
        // across the smoother step)
        Dual Ad[n * n];
        MaternBlock<G::family, Dual>::transition_rate(Dual(g.lam[c], 1.0), Dual(h), Ad);
        real s = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real z = v[c * n + i] * pm_[c * n + j];
#pragma unroll
                for (int l = 0; l < d; ++l) {
                    const bool nz = (l / n == c) && ((((l % n) + j) & 1) == 0);
                    const real dl = nz ? pP_[sidx(l, c * n + j)] - g.Pb[c * symn(n) + sidx(l % n, j)]
                                         : pP_[sidx(l, c * n + j)];
                    z = fma(XA[(c * n + i) * d + l], dl, z);
                }
                s = fma(z, Ad[i * n + j].d, s);
            }
        acc.gl[c] += s;
    }
}

// (Gamma blocks, gl) -> d ell / d variance_c, d ell / d lengthscale_c.  Pinf is linear in the variance and A
// does not depend on it; the lengthscale acts through Pinf and through lam = sqrt(2 nu) / lengthscale.
template <class G>
BN_DEV void grad_finish(const bn_kernel_spec& s, const real* fields, real* dvar, real* dlen) {
    constexpr int n = G::n, NC = G::NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        Dual P[symn(n)];
        MaternBlock<G::family, Dual>::pinf(Dual(s.variance[c]), Dual(s.lengthscale[c], 1.0), P);
        real a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const real w = (i == j ? 1.0 : 2.0) * fields[c * symn(n) + sidx(i, j)];
                a = fma(w, P[sidx(i, j)].v, a);
                b = fma(w, P[sidx(i, j)].d, b);
            }
        const real lam = MaternBlock<G::family, real>::rate(s.lengthscale[c]);
        dvar[c] = a / s.variance[c];
        dlen[c] = b - fields[NC * symn(n) + c] * lam / s.lengthscale[c];
    }
}

// -------------------------------------------------------------------------------- smoother step
// (sm, sP) at step k+1 in, at step k out (ops.py:290-301); Ab, Qb belong to the step k -> k+1.
// With GRAD the step also accumulates the hyper-gradient terms of the transition k -> k+1 (see below).
template <class G, bool GRAD = false>
BN_DEV void frts_step(const real* Ab, const real* Qb, const real* fm, const real* fP, real* sm,
                      real* sP, const G* g = nullptr, real h = 0.0, GradAcc<G>* acc = nullptr) {
    constexpr int d = G::d;
    real pm[d], AfP[d * d], pP[symn(d)];
    bd_matvec<G>(Ab, fm, pm);
    bd_mat_sym<G>(Ab, fP, AfP);
    bd_abt_sym<G>(AfP, Ab, Qb, pP);
    real dm[d], dP[symn(d)];
#pragma unroll
    for (int i = 0; i < d; ++i) dm[i] = sm[i] - pm[i];
#pragma unroll
    for (int i = 0; i < symn(d); ++i) dP[i] = sP[i] - pP[i];
    ldlt<d>(pP);
    if constexpr (GRAD) grad_accumulate<G>(*g, Ab, h, pP, dm, dP, fm, fP, false, *acc);
    ldlt_solve<d, d>(pP, AfP);  // AfP <- pP^-1 A fP = G^T
    // sm = fm + G dm,  G[i][j] = AfP[j][i]
#pragma unroll
    for (int i = 0; i < d; ++i) {
        real s = fm[i];
#pragma unroll
        for (int l = 0; l < d; ++l) s = fma(AfP[l * d + i], dm[l], s);
        sm[i] = s;
    }
    // sP = fP + G dP G^T
    real X[d * d];  // X = G dP
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            real s = 0.0;
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(AfP[l * d + i], dP[sidx(l, j)], s);
            X[i * d + j] = s;
        }
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real s = fP[sidx(i, j)];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(X[i * d + l], AfP[l * d + j], s);
            sP[sidx(i, j)] = s;
        }
}

// -------------------------------------------------------------------------------- smoothing element of a chunk
// The chunk of filter steps (k_a, k_b] is one generalised step: its filtering element
// (A, b, C, J, eta) gives x_b | x_a ~ N(A x_a + b, C) and the information (eta, J) its observations
// carry about x_a.  With the filtered states at both ends the RTS recursion over the whole chunk is
//     (m', P') = posterior of x_a given the data up to k_b     [(I + P_a J)^-1 (m_a + P_a eta), (I + P_a J)^-1 P_a]
//     E = P' A^T P_b^-1,   g = m' - E m_b,   L = P' - E P_b E^T
// i.e. the composition of the per-step smoothing elements of ops.py:318-335 over the chunk,
// obtained in O(d^3) per CHUNK instead of a second O(d^3) pass per STEP.
template <int d>
BN_DEV void chunk_smoothing_element(const typename FilterAlg<d>::Elem& fe, const real* ma, const real* Pa,
                                    const real* mb, const real* Pb, typename SmootherAlg<d>::Elem& se) {
    using FA = FilterAlg<d>;
    constexpr int c = d + 1;
    real B[d * c], v[d];
    symvec<d>(Pa, fe.eta, v);
#pragma unroll
    for (int i = 0; i < d; ++i) {
#pragma unroll
        for (int j = 0; j < d; ++j) B[i * c + j] = Pa[sidx(i, j)];
        B[i * c + d] = ma[i] + v[i];
    }
    FA::template solve_ipcj<c>(Pa, fe.J, B);
    real Pq[symn(d)], mq[d];
#pragma unroll
    for (int i = 0; i < d; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) Pq[sidx(i, j)] = 0.5 * (B[i * c + j] + B[j * c + i]);
        mq[i] = B[i * c + d];
    }
    real X[d * d];  // A P'
    mat_sym<d, d>(fe.A, Pq, X);
    real Lc[symn(d)];
#pragma unroll
    for (int i = 0; i < symn(d); ++i) Lc[i] = Pb[i];
    chol<d>(Lc);
    chol_solve<d, d>(Lc, X);  // X <- P_b^-1 A P' = E^T
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) se.E[i * d + j] = X[j * d + i];
    real t[d];
    matvec<d, d>(se.E, mb, t);
#pragma unroll
    for (int i = 0; i < d; ++i) se.g[i] = mq[i] - t[i];
    real Y[d * d];  // E P_b
    mat_sym<d, d>(se.E, Pb, Y);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            real s = 0.0;
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(Y[i * d + l], se.E[j * d + l], s);
            se.L[sidx(i, j)] = Pq[sidx(i, j)] - s;
        }
}

}  // namespace BN_NS
