"""Cubature tables (host side).  Same construction as bayesnewton/cubature.py:56-84: Gauss-Hermite
nodes from numpy's hermgauss, product grid with the first coordinate slowest, nodes scaled by
sqrt(2) and weights by pi^(-dim/2).  The table is tiny (Q x (dim+1) doubles); it is built once on
the host, cached, and handed to the site kernels as HOST arrays x[dim,Q], w[Q] (the library passes
a 1-D rule to its kernels by value and stages a 2-D rule into the workspace)."""
import itertools

import numpy as np
from numpy.polynomial.hermite import hermgauss


_cache = {}


def gauss_hermite(dim=1, num_quad_pts=20):
    gh_x, gh_w = hermgauss(num_quad_pts)
    x = np.array(list(itertools.product(*(gh_x,) * dim)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * dim))), 1)
    return np.sqrt(2) * x.T, w.T * np.pi ** (-0.5 * dim)


def symmetric_cubature_third_order(dim=1, kappa=None):
    """2 dim + 1 sigma points, exact to order 3 (cubature.py:87-117; kappa = 0 is the cubature Kalman filter rule)"""
    kappa = 0 if kappa is None else kappa
    w0, wm, u = kappa / (dim + kappa), 1 / (2 * (dim + kappa)), np.sqrt(dim + kappa)
    x = u * np.concatenate([np.zeros((dim, 1)), np.eye(dim), -np.eye(dim)], axis=1)
    w = np.concatenate([[w0], wm * np.ones(2 * dim)])
    return (x[0] if dim == 1 else x), w


def symmetric_cubature_fifth_order(dim=1):
    """2 dim^2 + 1 sigma points, exact to order 5 (McNamee & Stenger; cubature.py:120-162), dim = 1 or 2"""
    u = np.sqrt(3.0)
    A0 = 1 - dim * (1 / 3) ** 2 * (3 - 0.5 * (dim - 1))
    A1 = 0.5 * (1 / 3) ** 2 * (3 - (dim - 1))
    A2 = 0.25 * (1 / 3) ** 2
    if dim == 1:
        return np.array([0., u, -u]), np.array([A0, A1, A1])
    if dim == 2:
        x = np.array([[0., u, -u, 0., 0., u, -u, u, -u],
                      [0., 0., 0., u, -u, u, -u, -u, u]])
        return x, np.array([A0, A1, A1, A1, A1, A2, A2, A2, A2])
    raise NotImplementedError('the site kernels take 1-D and 2-D rules')


class Cubature:
    """callable cubature object, like the reference's (cubature.py:12-30): cubature(dim) -> (x [dim, Q], w [Q])"""

    def __init__(self, dim=None):
        self._stored = None if dim is None else self.get_cubature_points_and_weights(dim)

    def __call__(self, dim):
        return self._stored if self._stored is not None else self.get_cubature_points_and_weights(dim)

    def get_cubature_points_and_weights(self, dim):
        raise NotImplementedError


class GaussHermite(Cubature):
    """cubature.py:33-40"""

    def __init__(self, dim=None, num_cub_points=20):
        self.num_cub_points = num_cub_points
        super().__init__(dim)

    def get_cubature_points_and_weights(self, dim):
        return gauss_hermite(dim, self.num_cub_points)


class UnscentedThirdOrder(Cubature):
    """cubature.py:43-46"""

    def get_cubature_points_and_weights(self, dim):
        return symmetric_cubature_third_order(dim)


class UnscentedFifthOrder(Cubature):
    """cubature.py:49-52"""

    def get_cubature_points_and_weights(self, dim):
        return symmetric_cubature_fifth_order(dim)


class Unscented(UnscentedFifthOrder):
    pass


def host_table(cubature, dim):
    """(x [dim,Q], w [Q], Q) as contiguous float64 numpy arrays (None = Gauss-Hermite 20, the reference default).
    The default rule is cached per dimension; a custom rule keeps its tables on the object itself (an id()-keyed
    cache would hand a new object the table of a collected one)."""
    if cubature is None:
        hit = _cache.get(('gh20', dim))
        if hit is None:
            hit = _cache[('gh20', dim)] = _as_table(*gauss_hermite(dim))
        return hit
    store = getattr(cubature, '_bn_tables', None)
    if store is None:
        store = {}
        try:
            cubature._bn_tables = store
        except AttributeError:  # e.g. a plain function with __slots__-like restrictions: no caching
            pass
    hit = store.get(dim)
    if hit is None:
        hit = store[dim] = _as_table(*cubature(dim))
    return hit


def _as_table(x, w):
    x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
    w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1))
    return x, w, int(w.shape[0])
