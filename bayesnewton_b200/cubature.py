"""Cubature tables (host side).  Same construction as bayesnewton/cubature.py:56-84: Gauss-Hermite
nodes from numpy's hermgauss, product grid with the first coordinate slowest, nodes scaled by
sqrt(2) and weights by pi^(-dim/2).  The table is tiny (Q x (dim+1) doubles); it is built once on
the host, cached on the device, and handed to the site kernels as x[dim,Q], w[Q]."""
import itertools

import numpy as np
from numpy.polynomial.hermite import hermgauss

from ._util import as_dev

_cache = {}


def gauss_hermite(dim=1, num_quad_pts=20):
    gh_x, gh_w = hermgauss(num_quad_pts)
    x = np.array(list(itertools.product(*(gh_x,) * dim)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * dim))), 1)
    return np.sqrt(2) * x.T, w.T * np.pi ** (-0.5 * dim)


class GaussHermite:
    """callable cubature object, like the reference's (cubature.py:12-19): cubature(dim) -> (x, w)"""

    def __init__(self, num_cub_points=20):
        self.num_cub_points = num_cub_points

    def __call__(self, dim):
        return gauss_hermite(dim, self.num_cub_points)


def device_table(cubature, dim):
    """(x_dev [dim,Q], w_dev [Q], Q) for `cubature` (None = Gauss-Hermite 20, the reference default)"""
    key = ('gh20', dim) if cubature is None else (id(cubature), dim)
    hit = _cache.get(key)
    if hit is None:
        x, w = gauss_hermite(dim) if cubature is None else cubature(dim)
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        w = np.asarray(w, dtype=np.float64).reshape(-1)
        hit = (as_dev(x), as_dev(w), int(w.shape[0]))
        _cache[key] = hit
    return hit
