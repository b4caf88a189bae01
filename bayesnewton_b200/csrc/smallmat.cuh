// Small dense linear algebra held entirely in registers (fully unrolled, compile-time sizes).
// Symmetric matrices are stored packed lower-triangular: (i,j), i>=j  ->  i(i+1)/2 + j.
// Cholesky-based solve/inverse everywhere, mirroring bayesnewton/utils.py:14-35: a non-PD
// input yields NaN (sqrt of a negative), never a trap or a clamp.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "real.cuh"

namespace BN_NS {

__host__ __device__ constexpr int symn(int d) { return d * (d + 1) / 2; }
__host__ __device__ constexpr int sidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// y = A x            A[d*d] row-major
template <int r, int c, typename T>
BN_DEV void matvec(const T* A, const T* x, T* y) {
#pragma unroll
    for (int i = 0; i < r; ++i) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < c; ++k) s = fma(A[i * c + k], x[k], s);
        y[i] = s;
    }
}

// y = A^T x
template <int r, int c, typename T>
BN_DEV void matTvec(const T* A, const T* x, T* y) {
#pragma unroll
    for (int j = 0; j < c; ++j) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(A[k * c + j], x[k], s);
        y[j] = s;
    }
}

// y = S x, S symmetric packed
template <int d, typename T>
BN_DEV void symvec(const T* S, const T* x, T* y) {
#pragma unroll
    for (int i = 0; i < d; ++i) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < d; ++k) s = fma(S[sidx(i, k)], x[k], s);
        y[i] = s;
    }
}

// C = A B   (r x k)(k x c), all full row-major
template <int r, int k, int c, typename T>
BN_DEV void matmul(const T* A, const T* B, T* C) {
#pragma unroll
    for (int i = 0; i < r; ++i)
#pragma unroll
        for (int j = 0; j < c; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < k; ++l) s = fma(A[i * k + l], B[l * c + j], s);
            C[i * c + j] = s;
        }
}

// C = A^T B   A (k x r), B (k x c)
template <int r, int k, int c, typename T>
BN_DEV void matTmul(const T* A, const T* B, T* C) {
#pragma unroll
    for (int i = 0; i < r; ++i)
#pragma unroll
        for (int j = 0; j < c; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < k; ++l) s = fma(A[l * r + i], B[l * c + j], s);
            C[i * c + j] = s;
        }
}

// X = A S   A (r x d) full, S symmetric packed d  -> X (r x d) full
template <int r, int d, typename T>
BN_DEV void mat_sym(const T* A, const T* S, T* X) {
#pragma unroll
    for (int i = 0; i < r; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(A[i * d + l], S[sidx(l, j)], s);
            X[i * d + j] = s;
        }
}

// X = S A   S symmetric packed d, A (d x c) full -> X (d x c)
template <int d, int c, typename T>
BN_DEV void sym_mat(const T* S, const T* A, T* X) {
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < c; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(S[sidx(i, l)], A[l * c + j], s);
            X[i * c + j] = s;
        }
}

// out(sym, r) = X A^T + Q      X (r x k), A (r x k), lower triangle only; Q packed or null
template <int r, int k, typename T>
BN_DEV void abt_sym(const T* X, const T* A, const T* Q, T* out) {
#pragma unroll
    for (int i = 0; i < r; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = Q ? Q[sidx(i, j)] : T(0);
#pragma unroll
            for (int l = 0; l < k; ++l) s = fma(X[i * k + l], A[j * k + l], s);
            out[sidx(i, j)] = s;
        }
}

// out(sym, c) = A^T S A + Q   A (d x c) full, S packed d
template <int d, int c, typename T>
BN_DEV void atsa_sym(const T* A, const T* S, const T* Q, T* out) {
    T X[d * c];
    sym_mat<d, c>(S, A, X);
#pragma unroll
    for (int i = 0; i < c; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T s = Q ? Q[sidx(i, j)] : T(0);
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(A[l * c + i], X[l * c + j], s);
            out[sidx(i, j)] = s;
        }
}

// out(sym) = A S A^T + Q
template <int d, typename T>
BN_DEV void asat_sym(const T* A, const T* S, const T* Q, T* out) {
    T X[d * d];
    mat_sym<d, d>(A, S, X);
    abt_sym<d, d>(X, A, Q, out);
}

// in-place lower Cholesky of packed symmetric S (n x n).  NaN on non-PD input.
template <int n, typename T>
BN_DEV void chol(T* S) {
#pragma unroll
    for (int j = 0; j < n; ++j) {
        T s = S[sidx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-S[sidx(j, k)], S[sidx(j, k)], s);
        T ljj = sqrt(s);
        S[sidx(j, j)] = ljj;
        T inv = T(1) / ljj;
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            T t = S[sidx(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-S[sidx(i, k)], S[sidx(j, k)], t);
            S[sidx(i, j)] = t * inv;
        }
    }
}

// solve (L L^T) X = B in place; B is (n x c) row-major, L packed lower Cholesky factor
template <int n, int c, typename T>
BN_DEV void chol_solve(const T* L, T* B) {
    T dinv[n];
#pragma unroll
    for (int i = 0; i < n; ++i) dinv[i] = T(1) / L[sidx(i, i)];
#pragma unroll
    for (int j = 0; j < c; ++j) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T s = B[i * c + j];
#pragma unroll
            for (int k = 0; k < i; ++k) s = fma(-L[sidx(i, k)], B[k * c + j], s);
            B[i * c + j] = s * dinv[i];
        }
#pragma unroll
        for (int i = n - 1; i >= 0; --i) {
            T s = B[i * c + j];
#pragma unroll
            for (int k = i + 1; k < n; ++k) s = fma(-L[sidx(k, i)], B[k * c + j], s);
            B[i * c + j] = s * dinv[i];
        }
    }
}

// sum of log |L_ii|
template <int n, typename T>
BN_DEV T chol_logdiag(const T* L) {
    T s = T(0);
#pragma unroll
    for (int i = 0; i < n; ++i) s += log(fabs(L[sidx(i, i)]));
    return s;
}

// symmetric inverse through Cholesky: S (packed, destroyed -> factor), out packed inverse
template <int n, typename T>
BN_DEV void sym_inverse(T* S, T* out) {
    chol<n>(S);
    T B[n * n];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) B[i * n + j] = (i == j) ? T(1) : T(0);
    chol_solve<n, n>(S, B);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) out[sidx(i, j)] = B[i * n + j];
}

// general (non-symmetric) n x n inverse-apply by Gaussian elimination without pivoting on
// M = I + C J (C, J PSD => eigenvalues >= 1, pivots safe).  Solves M X = B in place, B (n x c).
template <int n, int c, typename T>
BN_DEV void lu_solve_nopivot(T* M, T* B) {
#pragma unroll
    for (int k = 0; k < n; ++k) {
        T pinv = T(1) / M[k * n + k];
#pragma unroll
        for (int i = k + 1; i < n; ++i) {
            T f = M[i * n + k] * pinv;
#pragma unroll
            for (int j = k + 1; j < n; ++j) M[i * n + j] = fma(-f, M[k * n + j], M[i * n + j]);
#pragma unroll
            for (int j = 0; j < c; ++j) B[i * c + j] = fma(-f, B[k * c + j], B[i * c + j]);
        }
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {
        T pinv = T(1) / M[i * n + i];
#pragma unroll
        for (int j = 0; j < c; ++j) {
            T s = B[i * c + j];
#pragma unroll
            for (int k = i + 1; k < n; ++k) s = fma(-M[i * n + k], B[k * c + j], s);
            B[i * c + j] = s * pinv;
        }
    }
}

}  // namespace BN_NS
