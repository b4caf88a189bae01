"""jax.scipy.linalg on SciPy (LAPACK).  A Cholesky of a non-PD matrix yields NaN, as in jax (no exception)."""
import numpy as _np
import scipy.linalg as _sl

from .. import numpy as _jnp
from .._dual import Dual


def _nan_like(a):
    return _np.full(_np.shape(a), _np.nan).view(_jnp.Arr)


def cholesky(a, lower=False):
    try:
        return _sl.cholesky(_np.asarray(a), lower=lower).view(_jnp.Arr)
    except (_np.linalg.LinAlgError, ValueError):
        return _nan_like(a)


def cho_factor(a, lower=False):
    try:
        c, low = _sl.cho_factor(_np.asarray(a), lower=lower)
        return c.view(_jnp.Arr), low
    except (_np.linalg.LinAlgError, ValueError):
        return _nan_like(a), lower


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    if isinstance(b, Dual):
        return Dual(cho_solve((c, lower), b.v), cho_solve((c, lower), b.t))
    if _np.isnan(c).any():
        return _nan_like(b)
    return _sl.cho_solve((_np.asarray(c), lower), _np.asarray(b)).view(_jnp.Arr)


def solve(a, b, **kw):
    return _sl.solve(_np.asarray(a), _np.asarray(b)).view(_jnp.Arr)


def inv(a):
    return _sl.inv(_np.asarray(a)).view(_jnp.Arr)


def expm(a):
    return _sl.expm(_np.asarray(a)).view(_jnp.Arr)


def block_diag(*arrs):
    return _sl.block_diag(*[_np.asarray(a) for a in arrs]).view(_jnp.Arr)
