"""jax.lax.scan / associative_scan on NumPy.  TEST INFRASTRUCTURE."""
import numpy as _np

from . import numpy as _jnp


def _leaves_len(xs):
    if isinstance(xs, (tuple, list)):
        for x in xs:
            n = _leaves_len(x)
            if n is not None:
                return n
        return None
    return None if xs is None else _np.shape(xs)[0]


def _slice(xs, i):
    if isinstance(xs, (tuple, list)):
        return type(xs)(_slice(x, i) for x in xs)
    return None if xs is None else _np.asarray(xs)[i].view(_jnp.Arr) if _np.ndim(_np.asarray(xs)[i]) else _np.asarray(xs)[i]


def _stack(ys):
    first = ys[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([y[k] for y in ys]) for k in range(len(first)))
    if first is None:
        return None
    return _np.stack([_np.asarray(y) for y in ys]).view(_jnp.Arr)


def scan(f, init, xs, length=None, reverse=False):
    """carry, stacked ys: the body runs in time order (reverse: from the last element, ys stacked in input order)"""
    n = length if length is not None else _leaves_len(xs)
    carry = init
    ys = [None] * n
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        carry, y = f(carry, _slice(xs, i))
        ys[i] = y
    return carry, _stack(ys)


def associative_scan(fn, elems, reverse=False, axis=0):
    """the odd/even recursion of jax.lax.associative_scan (same bracketing of the combines); fn acts on batches"""
    if axis != 0:
        raise NotImplementedError
    elems = [_np.asarray(e) for e in elems]
    if reverse:
        elems = [_np.flip(e, 0) for e in elems]

    def combine(a, b):
        if a[0].shape[0] == 0:  # an empty batch combines to an empty batch
            return [x[:0] for x in a]
        out = fn(tuple(a), tuple(b))
        return [_np.asarray(o) for o in out]

    def _scan(es):
        n = es[0].shape[0]
        if n < 2:
            return es
        reduced = combine([e[0:-1:2] for e in es], [e[1::2] for e in es])
        odd = _scan(reduced)
        if n % 2 == 0:
            even = combine([e[:-1] for e in odd], [e[2::2] for e in es])
        else:
            even = combine(odd, [e[2::2] for e in es])
        even = [_np.concatenate([e[0:1], r], axis=0) for e, r in zip(es, even)]
        out = []
        for ev, od in zip(even, odd):
            z = _np.empty((n,) + ev.shape[1:], dtype=ev.dtype)
            z[0::2] = ev
            z[1::2] = od
            out.append(z)
        return out

    res = _scan(elems)
    if reverse:
        res = [_np.flip(r, 0) for r in res]
    return tuple(r.view(_jnp.Arr) for r in res)
