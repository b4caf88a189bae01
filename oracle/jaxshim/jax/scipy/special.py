import numpy as _np
from scipy import special as _sps

from .._dual import Dual, lift as _lift


def erf(x):
    return _lift('erf', x) if isinstance(x, Dual) else _sps.erf(x)


def gammaln(x):
    return _lift('gammaln', x) if isinstance(x, Dual) else _sps.gammaln(x)


def logsumexp(a, axis=None, b=None, keepdims=False):
    return _sps.logsumexp(_np.asarray(a), axis=axis, b=b, keepdims=keepdims)
