"""ctypes front-end of oracle/c/markov_c.c (plain-C restatement of the reference's CPU path).
Test infrastructure and the `cpu_baseline` / `--impl reference` legs of bench.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libbn_oracle_c.so')


class CKernel(C.Structure):
    _fields_ = [('family', C.c_int), ('variance', C.c_double), ('lengthscale', C.c_double)]


def build():
    """compiles oracle/c/markov_c.c when its content changed (hash stamp: a copied tree does not rebuild)"""
    import hashlib
    src = os.path.join(HERE, 'c', 'markov_c.c')
    dig = hashlib.sha256(open(src, 'rb').read()).hexdigest()
    stamp = LIB + '.sha256'
    if not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig):
        subprocess.run(['make', '-B', '-C', HERE], check=True, capture_output=True)
        open(stamp, 'w').write(dig)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        P, I, L64, D = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.bnc_discretise.argtypes = [C.POINTER(CKernel), L64, P, P, P]
        L.bnc_sequential_kf.argtypes = [I, L64, P, P, P, P, P, P, P, P, P]
        L.bnc_sequential_kf.restype = D
        L.bnc_sequential_rts.argtypes = [I, L64, P, P, P, P, P, P]
        L.bnc_vi_iteration.argtypes = [C.POINTER(CKernel), I, D, L64, P, P, P, P, I, P, P, D] + [P] * 10
        L.bnc_vi_iteration.restype = D
        L.bnc_vi_iteration_blocked.argtypes = [C.POINTER(CKernel), I, D, L64, P, P, P, P, I, P, P, D] + [P] * 10 + [I]
        L.bnc_vi_iteration_blocked.restype = D
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ViModel:
    """MarkovVariationalGP state for the C port: Matern family (1..4), likelihood 1 = Gaussian, 2 = probit"""

    def __init__(self, family, variance, lengthscale, lik, lik_param, dt, y, num_quad_pts=20):
        from .sites import gauss_hermite
        self.k = CKernel(family, variance, lengthscale)
        self.lik, self.lik_param = lik, float(lik_param)
        self.N, self.d = dt.shape[0], family
        self.dt = np.ascontiguousarray(dt, dtype=np.float64)
        self.dts = np.ascontiguousarray(np.concatenate([dt[1:], [0.0]]))
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        m = np.isnan(self.y)
        self.mask = np.ascontiguousarray(m.astype(np.uint8)) if m.any() else None
        x, w = gauss_hermite(1, num_quad_pts)
        self.gx, self.gw = np.ascontiguousarray(x[0]), np.ascontiguousarray(w)
        N, d = self.N, self.d
        self.nat1, self.nat2 = np.zeros(N), np.full(N, 1e-2)
        self.site_mean, self.site_cov = np.zeros(N), np.full(N, 1e2)
        self.As, self.Qs = np.empty((N, d, d)), np.empty((N, d, d))
        self.fms, self.fPs = np.empty((N, d)), np.empty((N, d, d))
        self.post_mean, self.post_var = np.zeros(N), np.ones(N)

    def iteration(self, lr=1.0):
        """inference(lr) then energy(): returns the energy"""
        return lib().bnc_vi_iteration(C.byref(self.k), self.lik, self.lik_param, self.N, _p(self.dt), _p(self.dts),
                                      _p(self.y), _p(self.mask), self.gw.shape[0], _p(self.gx), _p(self.gw), lr,
                                      _p(self.nat1), _p(self.nat2), _p(self.site_mean), _p(self.site_cov),
                                      _p(self.As), _p(self.Qs), _p(self.fms), _p(self.fPs), _p(self.post_mean),
                                      _p(self.post_var))

    def iteration_blocked(self, lr=1.0, nblocks=None):
        """the same iteration with the filter / smoother in the temporally parallel form (parallel=True), time-blocked
        over the host threads (4 blocks per thread)"""
        if nblocks is None:
            nblocks = 4 * int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1))
        return lib().bnc_vi_iteration_blocked(C.byref(self.k), self.lik, self.lik_param, self.N, _p(self.dt), _p(self.dts),
                                              _p(self.y), _p(self.mask), self.gw.shape[0], _p(self.gx), _p(self.gw), lr,
                                              _p(self.nat1), _p(self.nat2), _p(self.site_mean), _p(self.site_cov),
                                              _p(self.As), _p(self.Qs), _p(self.fms), _p(self.fPs), _p(self.post_mean),
                                              _p(self.post_var), int(nblocks))
