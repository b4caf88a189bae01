import numpy as _np


def softmax(x, axis=-1):
    z = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
    return z / _np.sum(z, axis=axis, keepdims=True)


def softplus(x):
    return _np.logaddexp(x, 0.0)
