"""jax on NumPy: vmap (loop), grad / jacrev (forward-mode duals), Array, config.  See ../README.md.  TEST INFRASTRUCTURE."""
import numpy as _np

from . import numpy as _jnp
from ._dual import Dual, _primal, _zeros_like
from . import lax, nn, scipy, lib  # noqa: F401

__version__ = '0.4.14-shim'


class Array:  # annotation only
    pass


class _Config:
    def update(self, *a, **k):
        pass

    def __getattr__(self, name):
        return None


config = _Config()


# ---- pytrees: tuples / lists of arrays
def _is_leaf(x):
    return not isinstance(x, (tuple, list))


def _index(tree, axis, i):
    """slice i of `tree` along `axis` (None: not mapped); axis may mirror the tuple structure"""
    if axis is None:
        return tree
    if not _is_leaf(tree):
        axes = axis if isinstance(axis, (tuple, list)) else [axis] * len(tree)
        return type(tree)(_index(t, a, i) for t, a in zip(tree, axes))
    if hasattr(tree, 'value') and not isinstance(tree, (_np.ndarray, Dual)):
        tree = tree.value
    if isinstance(tree, Dual):
        return Dual(_index(tree.v, axis, i), _index(tree.t, axis, i))
    return _np.take(_np.asarray(tree), i, axis=axis).view(_jnp.Arr) if _np.ndim(tree) > 0 else tree


def _mapped_size(tree, axis):
    if axis is None:
        return None
    if not _is_leaf(tree):
        axes = axis if isinstance(axis, (tuple, list)) else [axis] * len(tree)
        for t, a in zip(tree, axes):
            n = _mapped_size(t, a)
            if n is not None:
                return n
        return None
    if hasattr(tree, 'value') and not isinstance(tree, (_np.ndarray, Dual)):
        tree = tree.value
    return _np.shape(_primal(tree))[axis]


def _stack(outs):
    first = outs[0]
    if not _is_leaf(first):
        return type(first)(_stack([o[k] for o in outs]) for k in range(len(first)))
    if any(isinstance(o, Dual) for o in outs):
        outs = [o if isinstance(o, Dual) else Dual(o, _zeros_like(o)) for o in outs]
        return Dual(_stack([o.v for o in outs]), _stack([o.t for o in outs]))
    return _np.stack([_np.asarray(o) for o in outs]).view(_jnp.Arr)


def vmap(fun, in_axes=0, out_axes=0):
    if out_axes != 0:
        raise NotImplementedError('jaxshim.vmap: out_axes != 0')

    def mapped(*args, **kwargs):
        axes = list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)
        if len(axes) != len(args):
            raise ValueError('vmap: in_axes does not match the arguments')
        n = None
        for a, ax in zip(args, axes):
            n = _mapped_size(a, ax)
            if n is not None:
                break
        outs = [fun(*[_index(a, ax, i) for a, ax in zip(args, axes)], **kwargs) for i in range(n)]
        return _stack(outs)
    return mapped


# ---- derivatives: one forward pass per input component
def _directional(fun, args, argnum, seed):
    a = list(args)
    a[argnum] = Dual(a[argnum], seed)
    return fun(*a)


def _tangent(out):
    if isinstance(out, (tuple, list)):
        return type(out)(_tangent(o) for o in out)
    return out.t if isinstance(out, Dual) else _zeros_like(out)


def jacfwd(fun, argnums=0):
    def jac(*args, **kwargs):
        f = (lambda *a: fun(*a, **kwargs))
        if isinstance(argnums, (tuple, list)):
            return tuple(_jac_one(f, args, k) for k in argnums)
        return _jac_one(f, args, argnums)
    return jac


def _jac_one(f, args, k):
    x = args[k]
    if hasattr(x, 'value') and not isinstance(x, (_np.ndarray, Dual)):
        x = x.value
    shape = _np.shape(_primal(x))
    cols = []
    for idx in _np.ndindex(*shape) if shape else [()]:
        seed = _np.zeros(shape)
        seed[idx] = 1.0
        if isinstance(x, Dual):  # nested differentiation: the seed is constant with respect to the outer variable
            seed = Dual(seed, _np.zeros(shape))
        a = list(args)
        a[k] = x
        cols.append(_tangent(_directional(f, a, k, seed)))
    return _assemble(cols, shape)


def _assemble(cols, in_shape):
    first = cols[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_assemble([c[i] for c in cols], in_shape) for i in range(len(first)))
    st = _stack(cols)  # [n_in, *out_shape]
    out_shape = _np.shape(_primal(first))
    # jax layout: out_shape + in_shape
    nd = len(out_shape)
    if isinstance(st, Dual):
        mv = lambda z: _np.moveaxis(z.reshape(in_shape + out_shape), list(range(len(in_shape))), list(range(nd, nd + len(in_shape))))
        return Dual(_deep(mv, st.v), _deep(mv, st.t))
    z = _np.asarray(st).reshape(in_shape + out_shape)
    return _np.moveaxis(z, list(range(len(in_shape))), list(range(nd, nd + len(in_shape)))).view(_jnp.Arr)


def _deep(f, x):
    if isinstance(x, Dual):
        return Dual(_deep(f, x.v), _deep(f, x.t))
    return f(_np.asarray(x))


jacrev = jacfwd


def grad(fun, argnums=0):
    j = jacfwd(fun, argnums)

    def g(*args, **kwargs):
        return j(*args, **kwargs)
    return g
