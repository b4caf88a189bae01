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


class GaussHermite:
    """callable cubature object, like the reference's (cubature.py:12-19): cubature(dim) -> (x, w)"""

    def __init__(self, num_cub_points=20):
        self.num_cub_points = num_cub_points

    def __call__(self, dim):
        return gauss_hermite(dim, self.num_cub_points)


def host_table(cubature, dim):
    """(x [dim,Q], w [Q], Q) as contiguous float64 numpy arrays (None = Gauss-Hermite 20, the reference default)"""
    key = ('gh20', dim) if cubature is None else (id(cubature), dim)
    hit = _cache.get(key)
    if hit is None:
        x, w = gauss_hermite(dim) if cubature is None else cubature(dim)
        x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
        w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1))
        hit = (x, w, int(w.shape[0]))
        _cache[key] = hit
    return hit
