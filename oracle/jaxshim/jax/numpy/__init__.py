"""jax.numpy on NumPy (float64).  See ../../README.md.  TEST INFRASTRUCTURE."""
import numpy as _np

from .._dual import Dual, lift as _lift

pi, nan, inf, newaxis, e = _np.pi, _np.nan, _np.inf, _np.newaxis, _np.e
float = _np.float64          # `np.float` (removed from NumPy) is used by the reference as a dtype
float64, float32, int32, int64, bool_ = _np.float64, _np.float32, _np.int32, _np.int64, _np.bool_
ndarray = _np.ndarray
linalg = _np.linalg


class Arr(_np.ndarray):
    """ndarray with jax's functional update helper"""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        b = _np.array(self.a, copy=True).view(Arr)
        b[self.idx] = v
        return b

    def add(self, v):
        b = _np.array(self.a, copy=True).view(Arr)
        _np.add.at(b, self.idx, v)
        return b

    def multiply(self, v):
        b = _np.array(self.a, copy=True).view(Arr)
        b[self.idx] = b[self.idx] * v
        return b


def _conv(x):
    if isinstance(x, Dual):
        return x
    if isinstance(x, _np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, tuple):
        return tuple(_conv(i) for i in x)
    if isinstance(x, list):
        return [_conv(i) for i in x]
    if isinstance(x, _np.generic):
        return _np.asarray(x).view(Arr)
    return x


def _wrap(f):
    def g(*a, **k):
        return _conv(f(*a, **k))
    g.__name__ = getattr(f, '__name__', 'f')
    return g


def _unvar(x):
    # objax variables passed where arrays are expected
    return x.value if hasattr(x, 'value') and not isinstance(x, (_np.ndarray, Dual)) else x


def array(x, dtype=None, **kw):
    x = _unvar(x)
    if isinstance(x, Dual):
        return x
    if isinstance(x, (list, tuple)) and _has_dual(x):
        return _lift('stack_nested', x)
    out = _np.array(x, dtype=dtype)
    if out.dtype.kind in 'iu' and dtype is None and not _is_int_input(x):
        out = out.astype(_np.float64)
    return out.view(Arr)


def _is_int_input(x):
    try:
        return _np.asarray(x).dtype.kind in 'iub'
    except Exception:
        return False


def _has_dual(x):
    if isinstance(x, Dual):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_dual(i) for i in x)
    return False


asarray = array


def __getattr__(name):
    attr = getattr(_np, name)
    if callable(attr) and not isinstance(attr, type):
        def g(*a, **k):
            if any(_has_dual(i) for i in a):
                return _lift(name, *a, **k)
            return _conv(attr(*[_unvar(i) for i in a], **k))
        g.__name__ = name
        return g
    return attr
