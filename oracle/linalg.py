"""Small dense linear algebra with JAX semantics (oracle; test infrastructure).

The reference never uses LU: ``solve``/``inv`` are Cholesky based
(``bayesnewton/utils.py:14-35``) and a non-PD input silently yields NaN
instead of raising.  ``np.linalg.cholesky`` raises, so the factorisation is
restated here as explicit column loops, batched over leading axes, and valid
for any float dtype including ``np.longdouble``.
"""
import numpy as np


def T(P):
    """swap the last two axes (utils.py:50-51)"""
    return np.swapaxes(P, -1, -2)


def chol(P):
    """lower Cholesky factor of P[..., n, n]; NaN (not an exception) when P is not PD."""
    P = np.asarray(P)
    n = P.shape[-1]
    L = np.zeros_like(P)
    with np.errstate(invalid='ignore', divide='ignore'):
        for j in range(n):
            s = P[..., j, j] - np.sum(L[..., j, :j] ** 2, axis=-1)
            ljj = np.sqrt(s)  # sqrt(negative) -> NaN, as in XLA's potrf path
            L[..., j, j] = ljj
            for i in range(j + 1, n):
                s = P[..., i, j] - np.sum(L[..., i, :j] * L[..., j, :j], axis=-1)
                L[..., i, j] = s / ljj
    return L


def tri_solve_lower(L, B):
    """solve L X = B, L lower-triangular, batched"""
    n = L.shape[-1]
    X = np.zeros(np.broadcast_shapes(L.shape[:-2], B.shape[:-2]) + B.shape[-2:], dtype=np.result_type(L, B))
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(n):
            s = B[..., i, :] - np.sum(L[..., i, :i, None] * X[..., :i, :], axis=-2)
            X[..., i, :] = s / L[..., i, i, None]
    return X


def tri_solve_upper_from_lower(L, B):
    """solve L^T X = B, batched"""
    n = L.shape[-1]
    X = np.zeros(np.broadcast_shapes(L.shape[:-2], B.shape[:-2]) + B.shape[-2:], dtype=np.result_type(L, B))
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(n - 1, -1, -1):
            s = B[..., i, :] - np.sum(L[..., i + 1:, i, None] * X[..., i + 1:, :], axis=-2)
            X[..., i, :] = s / L[..., i, i, None]
    return X


def cho_solve(L, B):
    return tri_solve_upper_from_lower(L, tri_solve_lower(L, B))


def solve(P, Q):
    """P^-1 Q through the Cholesky factor (utils.py:14-19)"""
    return cho_solve(chol(P), Q)


def inv(P):
    """P^-1 through the Cholesky factor (utils.py:22-35)"""
    P = np.asarray(P)
    eye = np.broadcast_to(np.eye(P.shape[-1], dtype=P.dtype), P.shape)
    return cho_solve(chol(P), eye)


def block_diag(*mats):
    n = sum(m.shape[0] for m in mats)
    k = sum(m.shape[1] for m in mats)
    out = np.zeros((n, k), dtype=np.result_type(*mats))
    i = j = 0
    for m in mats:
        out[i:i + m.shape[0], j:j + m.shape[1]] = m
        i += m.shape[0]
        j += m.shape[1]
    return out
