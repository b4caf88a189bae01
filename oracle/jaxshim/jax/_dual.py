"""Forward-mode dual numbers over NumPy arrays (nestable) for the shim's grad / jacrev.  TEST INFRASTRUCTURE.

A Dual carries a value array `v` and a tangent array `t` of the same shape (one directional derivative); values and
tangents may themselves be Duals, which gives second derivatives by nesting.  Only the operations the reference
differentiates through are provided."""
import numpy as np
from scipy import special as sps


def _val(x):
    return x.v if isinstance(x, Dual) else x


def _tan(x):
    return x.t if isinstance(x, Dual) else (np.zeros_like(np.asarray(x, dtype=np.float64)) if not isinstance(x, Dual) else None)


def _zeros_like(x):
    if isinstance(x, Dual):
        return Dual(_zeros_like(x.v), _zeros_like(x.t))
    return np.zeros_like(np.asarray(x, dtype=np.float64))


def _broadcast(t, shape):
    """tangent broadcast to the shape its value was broadcast to"""
    if isinstance(t, Dual):
        return Dual(_broadcast(t.v, shape), _broadcast(t.t, shape))
    t = np.asarray(t, dtype=np.float64)
    return t if t.shape == tuple(shape) else np.broadcast_to(t, shape).copy()


class Dual:
    __array_priority__ = 1000

    def __init__(self, v, t):
        self.v = v
        shape = np.shape(_primal(v))
        self.t = t if np.shape(_primal(t)) == shape else _broadcast(t, shape)

    # ---- structure
    @property
    def shape(self):
        return np.shape(_primal(self))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def T(self):
        return Dual(_T(self.v), _T(self.t))

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return Dual(self.v[idx], self.t[idx])

    def reshape(self, *s):
        return Dual(self.v.reshape(*s), self.t.reshape(*s))

    def squeeze(self, axis=None):
        return lift('squeeze', self, axis=axis)

    def sum(self, axis=None):
        return lift('sum', self, axis=axis)

    # ---- arithmetic
    def __neg__(self):
        return Dual(-self.v, -self.t)

    def __add__(self, o):
        return Dual(self.v + _val(o), self.t + (o.t if isinstance(o, Dual) else 0.0))
    __radd__ = __add__

    def __sub__(self, o):
        return Dual(self.v - _val(o), self.t - (o.t if isinstance(o, Dual) else 0.0))

    def __rsub__(self, o):
        return Dual(o - self.v, -self.t)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.t * o.v + self.v * o.t)
        return Dual(self.v * o, self.t * o)
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.t - q * o.t) / o.v)
        return Dual(self.v / o, self.t / o)

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -q * self.t / self.v)

    def __pow__(self, p):
        if isinstance(p, Dual):
            raise NotImplementedError('dual exponent')
        return Dual(self.v ** p, p * self.v ** (p - 1) * self.t)

    def __matmul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v @ o.v, self.t @ o.v + self.v @ o.t)
        return Dual(self.v @ o, self.t @ o)

    def __rmatmul__(self, o):
        return Dual(o @ self.v, o @ self.t)

    # comparisons act on the primal value
    def __lt__(self, o): return _primal(self) < _primal(o)
    def __le__(self, o): return _primal(self) <= _primal(o)
    def __gt__(self, o): return _primal(self) > _primal(o)
    def __ge__(self, o): return _primal(self) >= _primal(o)
    def __eq__(self, o): return _primal(self) == _primal(o)
    __hash__ = None


def _primal(x):
    while isinstance(x, Dual):
        x = x.v
    return x


def _T(x):
    return x.T


def _unary(f, df):
    def g(x, **k):
        if isinstance(x, Dual):
            return Dual(g(x.v), df(x.v) * x.t)
        return f(x)
    return g


def _erf(x):
    return lift('erf', x) if isinstance(x, Dual) else sps.erf(x)


_exp = lambda x: lift('exp', x) if isinstance(x, Dual) else np.exp(x)
_TABLE = {}
_TABLE['exp'] = _unary(np.exp, lambda v: _TABLE['exp'](v))
_TABLE['log'] = _unary(np.log, lambda v: 1.0 / v)
_TABLE['sqrt'] = _unary(np.sqrt, lambda v: 0.5 / _TABLE['sqrt'](v))
_TABLE['square'] = _unary(np.square, lambda v: 2.0 * v)
_TABLE['erf'] = _unary(sps.erf, lambda v: 2.0 / np.sqrt(np.pi) * _TABLE['exp'](-(v * v)))
_TABLE['sin'] = _unary(np.sin, lambda v: _TABLE['cos'](v))
_TABLE['cos'] = _unary(np.cos, lambda v: -_TABLE['sin'](v))
_TABLE['tanh'] = _unary(np.tanh, lambda v: 1.0 - _TABLE['tanh'](v) * _TABLE['tanh'](v))
_TABLE['abs'] = _unary(np.abs, lambda v: np.sign(_primal(v)))
_TABLE['digamma'] = _unary(sps.digamma, lambda v: sps.polygamma(1, _primal(v)))
_TABLE['gammaln'] = _unary(sps.gammaln, lambda v: _TABLE['digamma'](v))
_TABLE['log1p'] = _unary(np.log1p, lambda v: 1.0 / (1.0 + v))


def _struct(name):
    """operations that only rearrange entries: applied to value and tangent alike"""
    def g(x, *a, **k):
        if isinstance(x, Dual):
            return Dual(g(x.v, *a, **k), g(x.t, *a, **k))
        return getattr(np, name)(x, *a, **k)
    return g


for _n in ('squeeze', 'reshape', 'sum', 'transpose', 'swapaxes', 'expand_dims', 'diag', 'diagonal', 'trace', 'mean',
           'atleast_1d', 'atleast_2d', 'tile', 'repeat', 'broadcast_to', 'ravel', 'nansum', 'cumsum', 'flip'):
    _TABLE[_n] = _struct(_n)


def _where(c, a, b):
    c = _primal(c)
    if isinstance(a, Dual) or isinstance(b, Dual):
        av, at = (a.v, a.t) if isinstance(a, Dual) else (a, _zeros_like(_val(b)) * 0 + 0.0)
        bv, bt = (b.v, b.t) if isinstance(b, Dual) else (b, _zeros_like(_val(a)) * 0 + 0.0)
        return Dual(_where(c, av, bv), _where(c, at, bt))
    return np.where(c, a, b)


def _seq(name):
    def g(xs, *a, **k):
        if any(isinstance(x, Dual) for x in xs):
            xs = [x if isinstance(x, Dual) else Dual(x, _zeros_like(x)) for x in xs]
            return Dual(g([x.v for x in xs], *a, **k), g([x.t for x in xs], *a, **k))
        return getattr(np, name)(xs, *a, **k)
    return g


for _n in ('concatenate', 'stack', 'hstack', 'vstack', 'block'):
    _TABLE[_n] = _seq(_n)


def _stack_nested(x):
    if isinstance(x, (list, tuple)):
        return _TABLE['stack']([_stack_nested(i) for i in x])
    return x


def _maximum(a, b):
    return _where(_primal(a) >= _primal(b), a, b)


def _minimum(a, b):
    return _where(_primal(a) <= _primal(b), a, b)


def _isnan(x):
    return np.isnan(_primal(x))


def _matmul(a, b):
    if isinstance(a, Dual):
        return a.__matmul__(b)
    if isinstance(b, Dual):
        return b.__rmatmul__(a)
    return np.matmul(a, b)


def _zeros_like_fn(x, **k):
    return _zeros_like(x)


def _ones_like_fn(x, **k):
    return _zeros_like(x) + 1.0


_TABLE.update(where=_where, stack_nested=_stack_nested, maximum=_maximum, minimum=_minimum, isnan=_isnan, matmul=_matmul,
              dot=_matmul, zeros_like=_zeros_like_fn, ones_like=_ones_like_fn)


def lift(name, *a, **k):
    if name not in _TABLE:
        raise NotImplementedError('jaxshim: numpy.%s has no dual-number rule' % name)
    return _TABLE[name](*a, **k)


def make_dual(x, seed):
    return Dual(x, seed)
