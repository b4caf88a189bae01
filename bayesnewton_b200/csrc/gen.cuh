// Per-step state-space model generators: (A_k, Q_k) either read from caller arrays or
// generated in registers from dt_k with the closed-form Matern discretisation
// (reference: bayesnewton/kernels.py:158-165, 216-224, 273-286, 344-365; Q = Pinf - A Pinf A^T,
//  ops.py:149-151; Independent stacking = block diagonal, kernels.py:1535-1583).
// Generating in-kernel removes the As/Qs arrays (2 d^2 doubles per step) from HBM entirely.
#pragma once
#include "smallmat.cuh"
#include "../../include/bn_b200.h"

namespace BN_NS {

template <int FAMILY> struct FamilyDim;
template <> struct FamilyDim<BN_MATERN12> { static constexpr int value = 1; };
template <> struct FamilyDim<BN_MATERN32> { static constexpr int value = 2; };
template <> struct FamilyDim<BN_MATERN52> { static constexpr int value = 3; };
template <> struct FamilyDim<BN_MATERN72> { static constexpr int value = 4; };

// one component: block A (n x n full), Pinf (packed)
template <int FAMILY, typename T>
struct MaternBlock {
    static constexpr int n = FamilyDim<FAMILY>::value;

    static BN_DEV void pinf(T var, T ell, T* P) {
        if constexpr (FAMILY == BN_MATERN12) {
            P[0] = var;
        } else if constexpr (FAMILY == BN_MATERN32) {
            P[sidx(0, 0)] = var; P[sidx(1, 0)] = T(0); P[sidx(1, 1)] = T(3) * var / (ell * ell);
        } else if constexpr (FAMILY == BN_MATERN52) {
            T l2 = ell * ell;
            T kappa = T(5) / T(3) * var / l2;
            P[sidx(0, 0)] = var; P[sidx(1, 0)] = T(0); P[sidx(1, 1)] = kappa;
            P[sidx(2, 0)] = -kappa; P[sidx(2, 1)] = T(0); P[sidx(2, 2)] = T(25) * var / (l2 * l2);
        } else {
            T l2 = ell * ell;
            T k1 = T(7) / T(5) * var / l2;
            T k2 = T(9.8) * var / (l2 * l2);
            P[sidx(0, 0)] = var; P[sidx(1, 0)] = T(0); P[sidx(1, 1)] = k1;
            P[sidx(2, 0)] = -k1; P[sidx(2, 1)] = T(0); P[sidx(2, 2)] = k2;
            P[sidx(3, 0)] = T(0); P[sidx(3, 1)] = -k2; P[sidx(3, 2)] = T(0);
            P[sidx(3, 3)] = T(343) * var / (l2 * l2 * l2);
        }
    }

    // A = exp(-lam dt) (dt M + I)
    static BN_DEV T rate(T ell) {
        if constexpr (FAMILY == BN_MATERN12) return T(1) / ell;
        else if constexpr (FAMILY == BN_MATERN32) return sqrt(T(3)) / ell;
        else if constexpr (FAMILY == BN_MATERN52) return sqrt(T(5)) / ell;
        else return sqrt(T(7)) / ell;
    }
    static BN_DEV void transition(T ell, T dt, T* A) {
        if constexpr (FAMILY == BN_MATERN12) A[0] = exp(-dt / ell);
        else transition_rate(rate(ell), dt, A);
    }
    // the same closed forms given lam = sqrt(2 nu) / ell (hoisted out of per-step loops)
    static BN_DEV void transition_rate(T lam, T dt, T* A) {
        if constexpr (FAMILY == BN_MATERN12) {
            A[0] = exp(-dt * lam);
        } else if constexpr (FAMILY == BN_MATERN32) {
            T e = exp(-dt * lam);
            A[0] = e * (dt * lam + T(1));           A[1] = e * dt;
            A[2] = e * (dt * (-lam * lam));         A[3] = e * (dt * (-lam) + T(1));
        } else if constexpr (FAMILY == BN_MATERN52) {
            T dl = dt * lam;
            T e = exp(-dl);
            T l2 = lam * lam;
            A[0] = e * (dt * (lam * (T(0.5) * dl + T(1))) + T(1));
            A[1] = e * (dt * (dl + T(1)));
            A[2] = e * (dt * (T(0.5) * dt));
            A[3] = e * (dt * (T(-0.5) * dl * l2));
            A[4] = e * (dt * (lam * (T(1) - dl)) + T(1));
            A[5] = e * (dt * (T(1) - T(0.5) * dl));
            A[6] = e * (dt * (l2 * lam * (T(0.5) * dl - T(1))));
            A[7] = e * (dt * (l2 * (dl - T(3))));
            A[8] = e * (dt * (lam * (T(0.5) * dl - T(2))) + T(1));
        } else {
            T l2 = lam * lam, l3 = l2 * lam;
            T dl = dt * lam;
            T dl2 = dl * dl;
            T e = exp(-dl);
            A[0] = e * (dt * (lam * (T(1) + T(0.5) * dl + dl2 / T(6))) + T(1));
            A[1] = e * (dt * (T(1) + dl + T(0.5) * dl2));
            A[2] = e * (dt * (T(0.5) * dt * (T(1) + dl)));
            A[3] = e * (dt * (dt * dt / T(6)));
            A[4] = e * (dt * (-dl2 * l2 / T(6)));
            A[5] = e * (dt * (lam * (T(1) + T(0.5) * dl - T(0.5) * dl2)) + T(1));
            A[6] = e * (dt * (T(1) + dl - T(0.5) * dl2));
            A[7] = e * (dt * (dt * (T(0.5) - dl / T(6))));
            A[8] = e * (dt * (l3 * dl * (dl / T(6) - T(0.5))));
            A[9] = e * (dt * (dl * l2 * (T(0.5) * dl - T(2))));
            A[10] = e * (dt * (lam * (T(1) - T(2.5) * dl + T(0.5) * dl2)) + T(1));
            A[11] = e * (dt * (T(1) - dl + dl2 / T(6)));
            A[12] = e * (dt * (l2 * l2 * (dl - T(1) - dl2 / T(6))));
            A[13] = e * (dt * (l3 * (T(3.5) * dl - T(4) - T(0.5) * dl2)));
            A[14] = e * (dt * (l2 * (T(4) * dl - T(6) - T(0.5) * dl2)));
            A[15] = e * (dt * (lam * (T(1.5) * dl - T(3) - dl2 / T(6))) + T(1));
        }
    }
};

// Generator: stack of NC components of one family.  d = NC * n, D = NC.
template <int FAMILY, int NC, typename T = double>
struct MaternGen {
    static constexpr int n = FamilyDim<FAMILY>::value;
    static constexpr int d = NC * n;
    static constexpr int D = NC;
    static constexpr bool kArrays = false;
    bn_kernel_spec spec;
    const T* dt;

    BN_DEV void pinf(T* P) const {  // packed d
#pragma unroll
        for (int i = 0; i < symn(d); ++i) P[i] = T(0);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            T Pb[symn(n)];
            MaternBlock<FAMILY, T>::pinf(T(spec.variance[c]), T(spec.lengthscale[c]), Pb);
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(c * n + i, c * n + j)] = Pb[sidx(i, j)];
        }
    }
    BN_DEV void H(T* Hm) const {  // D x d
#pragma unroll
        for (int i = 0; i < D * d; ++i) Hm[i] = T(0);
#pragma unroll
        for (int c = 0; c < NC; ++c) Hm[c * d + c * n] = T(1);
    }
    BN_DEV void m0(T* m) const {
#pragma unroll
        for (int i = 0; i < d; ++i) m[i] = T(0);
    }
    // A (d x d full) and Q (packed) at step k
    BN_DEV void step(long long k, T* A, T* Q) const { step_dt(dt[k], A, Q); }
    BN_DEV void step_dt(T h, T* A, T* Q) const {
#pragma unroll
        for (int i = 0; i < d * d; ++i) A[i] = T(0);
#pragma unroll
        for (int i = 0; i < symn(d); ++i) Q[i] = T(0);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            T Ab[n * n], Pb[symn(n)], Qb[symn(n)], X[n * n];
            MaternBlock<FAMILY, T>::transition(T(spec.lengthscale[c]), h, Ab);
            MaternBlock<FAMILY, T>::pinf(T(spec.variance[c]), T(spec.lengthscale[c]), Pb);
            // Q = Pinf - A Pinf A^T   (ops.py:149-151)
            mat_sym<n, n>(Ab, Pb, X);
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    T s = T(0);
#pragma unroll
                    for (int l = 0; l < n; ++l) s = fma(X[i * n + l], Ab[j * n + l], s);
                    Qb[sidx(i, j)] = Pb[sidx(i, j)] - s;
                }
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j < n; ++j) A[(c * n + i) * d + c * n + j] = Ab[i * n + j];
#pragma unroll
                for (int j = 0; j <= i; ++j) Q[sidx(c * n + i, c * n + j)] = Qb[sidx(i, j)];
            }
        }
    }
};

// Generator: caller-provided As[N,d,d], Qs[N,d,d] (the generic entry every other kernel of the
// reference reaches through: ops.py:154,237,288,338).  Q is read from the lower triangle.
template <int d_, int D_, typename T = double>
struct ArrayGen {
    static constexpr int d = d_;
    static constexpr int D = D_;
    static constexpr bool kArrays = true;
    const T* As;
    const T* Qs;
    const T* Hm;   // [D,d]
    const T* m0p;  // [d]
    const T* P0p;  // [d,d]

    BN_DEV void pinf(T* P) const {
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) P[sidx(i, j)] = P0p[i * d + j];
    }
    BN_DEV void H(T* h) const {
#pragma unroll
        for (int i = 0; i < D * d; ++i) h[i] = Hm[i];
    }
    BN_DEV void m0(T* m) const {
#pragma unroll
        for (int i = 0; i < d; ++i) m[i] = m0p[i];
    }
    BN_DEV void step(long long k, T* A, T* Q) const {
        const T* a = As + k * (d * d);
        const T* q = Qs + k * (d * d);
#pragma unroll
        for (int i = 0; i < d * d; ++i) A[i] = a[i];
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Q[sidx(i, j)] = q[i * d + j];
    }
};

}  // namespace BN_NS
