// Per-step Kalman / RTS arithmetic and the two associative algebras of the temporally
// parallel forms, all in registers.
//   filter step            bayesnewton/ops.py:156-175   (+ mvn_logpdf, utils.py:376-396)
//   filtering elements     ops.py:183-200, 222-229;  operator ops.py:203-219
//   smoother step          ops.py:290-311
//   smoothing elements     ops.py:314-325;           operator ops.py:328-335
#pragma once
#include "smallmat.cuh"

namespace BN_NS {

constexpr double kLog2Pi = 1.8378770664093453;
constexpr double kInv2Pi = 0.15915494309189535;

// ------------------------------------------------------------------------------------------
// mvn_logpdf with the reference's mask rule.  S packed D (the innovation covariance, intact),
// e = y - obs_mean.  Returns log N.  msk: D bytes or null.
template <int D, typename T>
BN_DEV T mvn_logpdf_masked(const T* S, const T* e, const unsigned char* msk) {
    T Sm[symn(D)];
    T em[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        bool mi = msk && msk[i];
        em[i] = mi ? T(0) : e[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            bool mj = msk && msk[j];
            T v = (mi || mj) ? T(0) : S[sidx(i, j)];
            if (i == j && mi) v = T(kInv2Pi);
            Sm[sidx(i, j)] = v;
        }
    }
    chol<D>(Sm);
    T logdet = T(2) * chol_logdiag<D>(Sm);
    T x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = em[i];
    chol_solve<D, 1>(Sm, x);
    T dist = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) dist = fma(em[i], x[i], dist);
    return T(-0.5) * (dist + T(D) * T(kLog2Pi) + logdet);
}

// ------------------------------------------------------------------------------------------
// One Kalman predict+update (ops.py:156-175).  State (m, P packed) is overwritten with the
// updated state; (mp, Pp) receive the predicted state.  R is the D x D site covariance as stored
// (row-major full; the lower triangle is what the Cholesky reads).  Returns the log-likelihood
// increment when WANT_ELL.
template <int d, int D, bool WANT_ELL, typename T>
BN_DEV T kf_step(T* m, T* P, const T* A, const T* Q, const T* H, const T* y, const T* R,
                 const unsigned char* msk, T* mp, T* Pp) {
    matvec<d, d>(A, m, mp);
    asat_sym<d>(A, P, Q, Pp);
    T HP[D * d];
    mat_sym<D, d>(H, Pp, HP);
    T S[symn(D)];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = R[i * D + j];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], H[j * d + l], s);
            S[sidx(i, j)] = s;
        }
    T e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < d; ++l) s = fma(H[i * d + l], mp[l], s);
        e[i] = y[i] - s;
    }
    T ell = T(0);
    if constexpr (WANT_ELL) {
        if (D == 1 && !(msk && msk[0])) {
            // same arithmetic as the Cholesky path for a 1x1: L = sqrt(S), dist = e (e / L / L)
            T L = sqrt(S[0]);
            T x = (e[0] / L) / L;
            ell = T(-0.5) * (e[0] * x + T(kLog2Pi) + T(2) * log(fabs(L)));
        } else {
            ell = mvn_logpdf_masked<D>(S, e, msk);
        }
    }
    // K^T = S^-1 HP  (Cholesky solve, utils.py:14-19), unmasked S
    chol<D>(S);
    T Kt[D * d];
#pragma unroll
    for (int i = 0; i < D * d; ++i) Kt[i] = HP[i];
    chol_solve<D, d>(S, Kt);
#pragma unroll
    for (int i = 0; i < d; ++i) {
        T s = mp[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(Kt[a * d + i], e[a], s);
        m[i] = s;
    }
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = Pp[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) s = fma(-Kt[a * d + i], HP[a * d + j], s);
            P[sidx(i, j)] = s;
        }
    return ell;
}

// ------------------------------------------------------------------------------------------
// Filtering algebra.  Element = (A, b, C, J, eta): x_end | x_start, y ~ N(A x_start + b, C) and the
// information (eta, J) the block's observations carry about x_start.
template <int d, typename T = real>
struct FilterAlg {
    static constexpr int kElem = d * d + d + symn(d) + symn(d) + d;
    static constexpr int kState = d + symn(d);
    static constexpr int kCarry = 3 * d * d + 2 * d;  // full-storage layout of the public carry
    struct Elem { T A[d * d], b[d], C[symn(d)], J[symn(d)], eta[d]; };
    struct State { T m[d], P[symn(d)]; };

    static BN_DEV void identity(Elem& e) {
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) e.A[i * d + j] = (i == j) ? T(1) : T(0);
#pragma unroll
        for (int i = 0; i < d; ++i) { e.b[i] = T(0); e.eta[i] = T(0); }
#pragma unroll
        for (int i = 0; i < symn(d); ++i) { e.C[i] = T(0); e.J[i] = T(0); }
    }
    static BN_DEV void zero_state(State& s) {
#pragma unroll
        for (int i = 0; i < d; ++i) s.m[i] = T(0);
#pragma unroll
        for (int i = 0; i < symn(d); ++i) s.P[i] = T(0);
    }

    // M = I + C J (full), pivoted Gaussian elimination on [M | B], B (d x c).  C, J PSD => M is
    // non-singular (eigenvalues >= 1); equals solve(inv(C)+J, inv(C)) of ops.py:208-209 without
    // inverting C.
    template <int c>
    static BN_DEV void solve_ipcj(const T* C, const T* J, T* B) {
        T M[d * d];
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) {
                T s = (i == j) ? T(1) : T(0);
#pragma unroll
                for (int l = 0; l < d; ++l) s = fma(C[sidx(i, l)], J[sidx(l, j)], s);
                M[i * d + j] = s;
            }
#pragma unroll
        for (int k = 0; k < d; ++k) {
            // partial pivoting with predicated row exchanges (no dynamic register indexing)
#pragma unroll
            for (int i = k + 1; i < d; ++i) {
                bool sw = fabs(M[i * d + k]) > fabs(M[k * d + k]);
#pragma unroll
                for (int j = k; j < d; ++j) {
                    T a = M[k * d + j], bb = M[i * d + j];
                    M[k * d + j] = sw ? bb : a;
                    M[i * d + j] = sw ? a : bb;
                }
#pragma unroll
                for (int j = 0; j < c; ++j) {
                    T a = B[k * c + j], bb = B[i * c + j];
                    B[k * c + j] = sw ? bb : a;
                    B[i * c + j] = sw ? a : bb;
                }
            }
            T pinv = T(1) / M[k * d + k];
#pragma unroll
            for (int i = k + 1; i < d; ++i) {
                T f = M[i * d + k] * pinv;
#pragma unroll
                for (int j = k + 1; j < d; ++j) M[i * d + j] = fma(-f, M[k * d + j], M[i * d + j]);
#pragma unroll
                for (int j = 0; j < c; ++j) B[i * c + j] = fma(-f, B[k * c + j], B[i * c + j]);
            }
        }
#pragma unroll
        for (int i = d - 1; i >= 0; --i) {
            T pinv = T(1) / M[i * d + i];
#pragma unroll
            for (int j = 0; j < c; ++j) {
                T s = B[i * c + j];
#pragma unroll
                for (int k = i + 1; k < d; ++k) s = fma(-M[i * d + k], B[k * c + j], s);
                B[i * c + j] = s * pinv;
            }
        }
    }

    // out = op(e1 = earlier, e2 = later)   (ops.py:203-219)
    static BN_DEV void combine(const Elem& e1, const Elem& e2, Elem& out) {
        constexpr int c = 2 * d + 1;
        T B[d * c];  // [A1 | C1 | b1 + C1 eta2]
        T v[d];
        symvec<d>(e1.C, e2.eta, v);
#pragma unroll
        for (int i = 0; i < d; ++i) {
#pragma unroll
            for (int j = 0; j < d; ++j) {
                B[i * c + j] = e1.A[i * d + j];
                B[i * c + d + j] = e1.C[sidx(i, j)];
            }
            B[i * c + 2 * d] = e1.b[i] + v[i];
        }
        solve_ipcj<c>(e1.C, e2.J, B);  // B = T [A1 | C1 | b1 + C1 eta2]
        T TA1[d * d], W[symn(d)], tb[d];
#pragma unroll
        for (int i = 0; i < d; ++i) {
#pragma unroll
            for (int j = 0; j < d; ++j) TA1[i * d + j] = B[i * c + j];
#pragma unroll
            for (int j = 0; j <= i; ++j) W[sidx(i, j)] = T(0.5) * (B[i * c + d + j] + B[j * c + d + i]);
            tb[i] = B[i * c + 2 * d];
        }
        Elem r;
        matmul<d, d, d>(e2.A, TA1, r.A);
        matvec<d, d>(e2.A, tb, r.b);
#pragma unroll
        for (int i = 0; i < d; ++i) r.b[i] += e2.b[i];
        asat_sym<d>(e2.A, W, e2.C, r.C);
        // eta = (T A1)^T (eta2 - J2 b1) + eta1
        T u[d];
        symvec<d>(e2.J, e1.b, u);
#pragma unroll
        for (int i = 0; i < d; ++i) u[i] = e2.eta[i] - u[i];
        matTvec<d, d>(TA1, u, r.eta);
#pragma unroll
        for (int i = 0; i < d; ++i) r.eta[i] += e1.eta[i];
        // J = (T A1)^T J2 A1 + J1
        T JA[d * d];
        sym_mat<d, d>(e2.J, e1.A, JA);
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T s = e1.J[sidx(i, j)];
#pragma unroll
                for (int l = 0; l < d; ++l) s = fma(TA1[l * d + i], JA[l * d + j], s);
                r.J[sidx(i, j)] = s;
            }
        out = r;
    }

    // state after the block = element applied to the state before it
    static BN_DEV void apply(const Elem& e, const State& s, State& out) {
        constexpr int c = d + 1;
        T B[d * c];  // [P | m + P eta]
        T v[d];
        symvec<d>(s.P, e.eta, v);
#pragma unroll
        for (int i = 0; i < d; ++i) {
#pragma unroll
            for (int j = 0; j < d; ++j) B[i * c + j] = s.P[sidx(i, j)];
            B[i * c + d] = s.m[i] + v[i];
        }
        solve_ipcj<c>(s.P, e.J, B);
        T W[symn(d)], tb[d];
#pragma unroll
        for (int i = 0; i < d; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) W[sidx(i, j)] = T(0.5) * (B[i * c + j] + B[j * c + i]);
            tb[i] = B[i * c + d];
        }
        State r;
        matvec<d, d>(e.A, tb, r.m);
#pragma unroll
        for (int i = 0; i < d; ++i) r.m[i] += e.b[i];
        asat_sym<d>(e.A, W, e.C, r.P);
        out = r;
    }

    // SoA storage: field f of element i at base[f * stride + i]
    static BN_DEV void load(const T* base, long long stride, long long i, Elem& e) {
        const T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d * d; ++k) e.A[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < d; ++k) e.b[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) e.C[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) e.J[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < d; ++k) e.eta[k] = p[(f++) * stride];
    }
    static BN_DEV void store(T* base, long long stride, long long i, const Elem& e) {
        T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d * d; ++k) p[(f++) * stride] = e.A[k];
#pragma unroll
        for (int k = 0; k < d; ++k) p[(f++) * stride] = e.b[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(f++) * stride] = e.C[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(f++) * stride] = e.J[k];
#pragma unroll
        for (int k = 0; k < d; ++k) p[(f++) * stride] = e.eta[k];
    }
    static __device__ __forceinline__ void shfl_up(Elem& e, int delta) {
        T* p = reinterpret_cast<T*>(&e);
#pragma unroll
        for (int k = 0; k < kElem; ++k) p[k] = __shfl_up_sync(0xffffffffu, p[k], delta);
    }
    // public carry layout (full matrices): A[d,d], b[d], C[d,d], J[d,d], eta[d]
    static BN_DEV void to_carry(const Elem& e, T* c) {
        int f = 0;
        for (int k = 0; k < d * d; ++k) c[f++] = e.A[k];
        for (int k = 0; k < d; ++k) c[f++] = e.b[k];
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) c[f++] = e.C[sidx(i, j)];
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) c[f++] = e.J[sidx(i, j)];
        for (int k = 0; k < d; ++k) c[f++] = e.eta[k];
    }
    static BN_DEV void from_carry(const T* c, Elem& e) {
        int f = 0;
        for (int k = 0; k < d * d; ++k) e.A[k] = c[f++];
        for (int k = 0; k < d; ++k) e.b[k] = c[f++];
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { if (j <= i) e.C[sidx(i, j)] = c[f]; ++f; }
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { if (j <= i) e.J[sidx(i, j)] = c[f]; ++f; }
        for (int k = 0; k < d; ++k) e.eta[k] = c[f++];
    }
    static BN_DEV void load_state(const T* base, long long stride, long long i, State& s) {
        const T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d; ++k) s.m[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) s.P[k] = p[(f++) * stride];
    }
    static BN_DEV void store_state(T* base, long long stride, long long i, const State& s) {
        T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d; ++k) p[(f++) * stride] = s.m[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(f++) * stride] = s.P[k];
    }
};

// Absorb one time step into a block aggregate (the element of ops.py:183-200 combined onto the
// running aggregate by ops.py:203-219, evaluated as a zero-prior Kalman step that also tracks the
// sensitivity A of the mean to x_start and the information (eta, J) about x_start).
// `first`: global step 0 of the scan form, where Q_0 := P_0 and b absorbs m0 (ops.py:222-229);
// the caller passes Q = P0 and m0.
template <int d, int D, typename T>
BN_DEV void filter_absorb(typename FilterAlg<d, T>::Elem& g, const T* A, const T* Q, const T* H,
                          const T* y, const T* R, bool first, const T* m0) {
    T mp[d], Pp[symn(d)], Phi[d * d];
    matvec<d, d>(A, g.b, mp);
    asat_sym<d>(A, g.C, Q, Pp);
    matmul<d, d, d>(A, g.A, Phi);
    T HP[D * d], HPhi[D * d];
    mat_sym<D, d>(H, Pp, HP);
    matmul<D, d, d>(H, Phi, HPhi);
    T S[symn(D)];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = R[i * D + j];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(HP[i * d + l], H[j * d + l], s);
            S[sidx(i, j)] = s;
        }
    T e[D], ey[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T s = T(0), s0 = T(0);
#pragma unroll
        for (int l = 0; l < d; ++l) {
            s = fma(H[i * d + l], mp[l], s);
            if (first) s0 = fma(H[i * d + l], m0[l], s0);
        }
        ey[i] = y[i] - s;                 // innovation seen by (eta): y - H A b
        e[i] = first ? (y[i] - s0) : ey[i];  // innovation seen by b on the first step: y - H m0
    }
    chol<D>(S);
    // Kt = S^-1 HP ;  V = S^-1 [H Phi | ey]
    T Kt[D * d];
#pragma unroll
    for (int i = 0; i < D * d; ++i) Kt[i] = HP[i];
    chol_solve<D, d>(S, Kt);
    T V[D * (d + 1)];
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < d; ++j) V[i * (d + 1) + j] = HPhi[i * d + j];
        V[i * (d + 1) + d] = ey[i];
    }
    chol_solve<D, d + 1>(S, V);
    // eta += (H Phi)^T S^-1 ey ;  J += (H Phi)^T S^-1 (H Phi)
#pragma unroll
    for (int i = 0; i < d; ++i) {
        T s = g.eta[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(HPhi[a * d + i], V[a * (d + 1) + d], s);
        g.eta[i] = s;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T t = g.J[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(HPhi[a * d + i], V[a * (d + 1) + j], t);
            g.J[sidx(i, j)] = t;
        }
    }
    // b = (first ? m0 : mp) + K e ; C = Pp - K HP ; A = Phi - K H Phi
#pragma unroll
    for (int i = 0; i < d; ++i) {
        T s = first ? m0[i] : mp[i];
#pragma unroll
        for (int a = 0; a < D; ++a) s = fma(Kt[a * d + i], e[a], s);
        g.b[i] = s;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            T t = Phi[i * d + j];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], HPhi[a * d + j], t);
            g.A[i * d + j] = t;
        }
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T t = Pp[sidx(i, j)];
#pragma unroll
            for (int a = 0; a < D; ++a) t = fma(-Kt[a * d + i], HP[a * d + j], t);
            g.C[sidx(i, j)] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Smoothing algebra.  Element (E, g, L): x_k | x_{k+1} ~ N(E x_{k+1} + g, L).  combine(e1, e2)
// takes e1 = the already accumulated LATER part and e2 = the EARLIER element (ops.py:328-335).
template <int d, typename T = real>
struct SmootherAlg {
    static constexpr int kElem = d * d + d + symn(d);
    static constexpr int kState = d + symn(d);
    static constexpr int kCarry = 2 * d * d + d;
    struct Elem { T E[d * d], g[d], L[symn(d)]; };
    struct State { T m[d], P[symn(d)]; };

    static BN_DEV void identity(Elem& e) {
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) e.E[i * d + j] = (i == j) ? T(1) : T(0);
#pragma unroll
        for (int i = 0; i < d; ++i) e.g[i] = T(0);
#pragma unroll
        for (int i = 0; i < symn(d); ++i) e.L[i] = T(0);
    }
    static BN_DEV void zero_state(State& s) {
#pragma unroll
        for (int i = 0; i < d; ++i) s.m[i] = T(0);
#pragma unroll
        for (int i = 0; i < symn(d); ++i) s.P[i] = T(0);
    }
    static BN_DEV void combine(const Elem& e1, const Elem& e2, Elem& out) {
        Elem r;
        matmul<d, d, d>(e2.E, e1.E, r.E);
        matvec<d, d>(e2.E, e1.g, r.g);
#pragma unroll
        for (int i = 0; i < d; ++i) r.g[i] += e2.g[i];
        asat_sym<d>(e2.E, e1.L, e2.L, r.L);
        out = r;
    }
    static BN_DEV void apply(const Elem& e, const State& s, State& out) {
        State r;
        matvec<d, d>(e.E, s.m, r.m);
#pragma unroll
        for (int i = 0; i < d; ++i) r.m[i] += e.g[i];
        asat_sym<d>(e.E, s.P, e.L, r.P);
        out = r;
    }
    static BN_DEV void load(const T* base, long long stride, long long i, Elem& e) {
        const T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d * d; ++k) e.E[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < d; ++k) e.g[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) e.L[k] = p[(f++) * stride];
    }
    static BN_DEV void store(T* base, long long stride, long long i, const Elem& e) {
        T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d * d; ++k) p[(f++) * stride] = e.E[k];
#pragma unroll
        for (int k = 0; k < d; ++k) p[(f++) * stride] = e.g[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(f++) * stride] = e.L[k];
    }
    static __device__ __forceinline__ void shfl_up(Elem& e, int delta) {
        T* p = reinterpret_cast<T*>(&e);
#pragma unroll
        for (int k = 0; k < kElem; ++k) p[k] = __shfl_up_sync(0xffffffffu, p[k], delta);
    }
    static BN_DEV void to_carry(const Elem& e, T* c) {
        int f = 0;
        for (int k = 0; k < d * d; ++k) c[f++] = e.E[k];
        for (int k = 0; k < d; ++k) c[f++] = e.g[k];
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) c[f++] = e.L[sidx(i, j)];
    }
    static BN_DEV void from_carry(const T* c, Elem& e) {
        int f = 0;
        for (int k = 0; k < d * d; ++k) e.E[k] = c[f++];
        for (int k = 0; k < d; ++k) e.g[k] = c[f++];
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { if (j <= i) e.L[sidx(i, j)] = c[f]; ++f; }
    }
    static BN_DEV void load_state(const T* base, long long stride, long long i, State& s) {
        const T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d; ++k) s.m[k] = p[(f++) * stride];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) s.P[k] = p[(f++) * stride];
    }
    static BN_DEV void store_state(T* base, long long stride, long long i, const State& s) {
        T* p = base + i;
        int f = 0;
#pragma unroll
        for (int k = 0; k < d; ++k) p[(f++) * stride] = s.m[k];
#pragma unroll
        for (int k = 0; k < symn(d); ++k) p[(f++) * stride] = s.P[k];
    }
};

// smoother gain and predicted moments for one step: G = (pP^-1 A fP)^T  (ops.py:294-299)
template <int d, typename T>
BN_DEV void rts_gain(const T* fm, const T* fP, const T* A, const T* Q, T* G, T* pm, T* pP) {
    matvec<d, d>(A, fm, pm);
    T AfP[d * d];
    mat_sym<d, d>(A, fP, AfP);
    abt_sym<d, d>(AfP, A, Q, pP);
    T Lc[symn(d)];
#pragma unroll
    for (int i = 0; i < symn(d); ++i) Lc[i] = pP[i];
    chol<d>(Lc);
    chol_solve<d, d>(Lc, AfP);  // AfP <- pP^-1 A fP
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) G[i * d + j] = AfP[j * d + i];
}

// sequential RTS step (ops.py:300-301): (sm, sP) <- fm + G (sm - pm), fP + G (sP - pP) G^T
template <int d, typename T>
BN_DEV void rts_step(T* sm, T* sP, const T* fm, const T* fP, const T* G, const T* pm, const T* pP) {
    T dm[d], dP[symn(d)];
#pragma unroll
    for (int i = 0; i < d; ++i) dm[i] = sm[i] - pm[i];
#pragma unroll
    for (int i = 0; i < symn(d); ++i) dP[i] = sP[i] - pP[i];
    T t[d];
    matvec<d, d>(G, dm, t);
#pragma unroll
    for (int i = 0; i < d; ++i) sm[i] = fm[i] + t[i];
    asat_sym<d>(G, dP, fP, sP);
}

// smoothing element of one step (ops.py:318-325): E = G, g = fm - E A fm, L = fP - E pP E^T
template <int d, typename T>
BN_DEV void rts_element(const T* fm, const T* fP, const T* A, const T* Q, typename SmootherAlg<d, T>::Elem& e) {
    T pm[d], pP[symn(d)];
    rts_gain<d>(fm, fP, A, Q, e.E, pm, pP);
    T t[d];
    matvec<d, d>(e.E, pm, t);
#pragma unroll
    for (int i = 0; i < d; ++i) e.g[i] = fm[i] - t[i];
    T X[d * d];
    mat_sym<d, d>(e.E, pP, X);
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(X[i * d + l], e.E[j * d + l], s);
            e.L[sidx(i, j)] = fP[sidx(i, j)] - s;
        }
}

}  // namespace BN_NS
