"""jax.scipy.linalg on SciPy (LAPACK), batched over leading axes like jax.  A Cholesky of a non-PD matrix yields NaN for
THAT matrix, as in jax (no exception).  TEST INFRASTRUCTURE."""
import numpy as _np
import scipy.linalg as _sl

from .. import numpy as _jnp
from .._dual import Dual


def _batched(f, a, *rest):
    """apply f to the trailing 2-D matrices of `a` (and the matching slices of `rest`)"""
    a = _np.asarray(a, dtype=_np.float64)
    if a.ndim <= 2:
        return f(a, *rest)
    lead = a.shape[:-2]
    outs = [f(a[idx], *[r[idx] for r in rest]) for idx in _np.ndindex(*lead)]
    return _np.stack(outs).reshape(lead + outs[0].shape)


def _chol1(a, lower):
    try:
        return _sl.cholesky(a, lower=lower)
    except (_np.linalg.LinAlgError, ValueError):
        return _np.full(a.shape, _np.nan)


def cholesky(a, lower=False):
    return _batched(lambda m: _chol1(m, lower), a).view(_jnp.Arr)


def cho_factor(a, lower=False):
    # jax returns the triangular factor with the other triangle zeroed by its cholesky; consumers only use cho_solve
    return _batched(lambda m: _chol1(m, lower), a).view(_jnp.Arr), lower


def _solve1(c, b, lower):
    if _np.isnan(c).any():
        return _np.full(b.shape, _np.nan)
    return _sl.cho_solve((c, lower), b)


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    if isinstance(b, Dual):
        return Dual(cho_solve((c, lower), b.v), cho_solve((c, lower), b.t))
    c = _np.asarray(c, dtype=_np.float64)
    b = _np.asarray(b, dtype=_np.float64)
    if c.ndim > 2:
        b = _np.broadcast_to(b, c.shape[:-2] + b.shape[-2:]) if b.ndim >= 2 else b
        return _batched(lambda m, r: _solve1(m, r, lower), c, b).view(_jnp.Arr)
    return _solve1(c, b, lower).view(_jnp.Arr)


def solve(a, b, **kw):
    return _np.linalg.solve(_np.asarray(a), _np.asarray(b)).view(_jnp.Arr)


def inv(a):
    return _np.linalg.inv(_np.asarray(a)).view(_jnp.Arr)


def expm(a):
    return _sl.expm(_np.asarray(a)).view(_jnp.Arr)


def block_diag(*arrs):
    return _sl.block_diag(*[_np.asarray(a) for a in arrs]).view(_jnp.Arr)
