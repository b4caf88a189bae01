"""Host-side kernel objects: the state-space interface of the reference's kernel classes.

Mirrors the four methods the Markov hot path calls on a kernel (bayesnewton/kernels.py):
``stationary_covariance()``, ``measurement_model()``, ``state_transition(dt)`` and
``feedback_matrix()`` for Matern12/32/52/72 (:123-382) and ``Independent`` (:1499-1616).
They hold the hyper-parameters on the host and describe themselves to the CUDA library as a
``bn_kernel_spec`` so that A_k and Q_k are generated inside the filter/smoother kernels; the
matrix-returning methods exist for API parity and run ``bn_discretise`` on the GPU.
"""
import math

import numpy as np
import torch

from . import _lib
from ._util import as_dev, stream_ptr


def softplus(x):
    return math.log(1.0 + math.exp(x))  # utils.py:54-55


def softplus_inv(x):
    return math.log(math.exp(x) - 1.0)  # utils.py:67-74


class Kernel:
    """base: anything with the four state-space methods can be fed to ops.kalman_filter"""
    def spec(self):
        """bn_kernel_spec for in-kernel discretisation, or None (then ops materialises As, Qs)"""
        return None


class StationaryKernel(Kernel):
    family = None
    state_dim = None

    def __init__(self, variance=1.0, lengthscale=1.0, fix_variance=False, fix_lengthscale=False):
        # stored softplus-transformed like the reference (kernels.py:80-95); gradients are taken
        # w.r.t. the untransformed values and chained by the caller
        self.transformed_variance = softplus_inv(float(variance))
        self.transformed_lengthscale = softplus_inv(float(lengthscale))
        self.fix_variance, self.fix_lengthscale = fix_variance, fix_lengthscale

    @property
    def variance(self):
        return softplus(self.transformed_variance)

    @property
    def lengthscale(self):
        return softplus(self.transformed_lengthscale)

    def spec(self):
        return _lib.kernel_spec(self.family, [self.variance], [self.lengthscale])

    def chain_to_transformed(self, grad):
        """[2, 1] gradient w.r.t. (variance, lengthscale) -> w.r.t. the stored softplus-transformed variables
        (d softplus(x)/dx = sigmoid(x)), the quantities objax's optimiser steps on (kernels.py:80-95)"""
        sig = lambda x: 1.0 / (1.0 + math.exp(-x))
        return grad * torch.tensor([[sig(self.transformed_variance)], [sig(self.transformed_lengthscale)]],
                                   dtype=grad.dtype, device=grad.device)

    def measurement_model(self):
        H = np.zeros((1, self.state_dim))
        H[0, 0] = 1.0
        return H

    def K(self, X, X2):
        """covariance function on host arrays (kernels.py:97-104, clip at 1e-36); used for the per-hyper-parameter
        constants of the spatial conditional (spacetime.py), never per time step"""
        X = np.asarray(X, dtype=np.float64).reshape(-1, 1) / self.lengthscale
        X2 = np.asarray(X2, dtype=np.float64).reshape(-1, 1) / self.lengthscale
        return self.K_r(np.sqrt(np.maximum((X - X2.T) ** 2, 1e-36)))

    __call__ = K

    def state_transition(self, dt):
        """A = expm(F dt): [d,d] for a scalar dt, [N,d,d] for an array (the vmapped call of ops.py:277)"""
        As, _ = discretise(self, dt)
        return As[0] if np.ndim(dt) == 0 else As


class Matern12(StationaryKernel):
    family, state_dim = _lib.BN_MATERN12, 1

    def K_r(self, r):
        return self.variance * np.exp(-r)

    def stationary_covariance(self):
        return np.array([[self.variance]])

    def feedback_matrix(self):
        return np.array([[-1.0 / self.lengthscale]])


class Matern32(StationaryKernel):
    family, state_dim = _lib.BN_MATERN32, 2

    def K_r(self, r):
        s3 = 3.0 ** 0.5
        return self.variance * (1.0 + s3 * r) * np.exp(-s3 * r)

    def stationary_covariance(self):
        return np.array([[self.variance, 0.0], [0.0, 3.0 * self.variance / self.lengthscale ** 2]])

    def feedback_matrix(self):
        lam = 3.0 ** 0.5 / self.lengthscale
        return np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]])


class Matern52(StationaryKernel):
    family, state_dim = _lib.BN_MATERN52, 3

    def K_r(self, r):
        s5 = 5.0 ** 0.5
        return self.variance * (1.0 + s5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-s5 * r)

    def stationary_covariance(self):
        kappa = 5.0 / 3.0 * self.variance / self.lengthscale ** 2
        return np.array([[self.variance, 0.0, -kappa], [0.0, kappa, 0.0],
                         [-kappa, 0.0, 25.0 * self.variance / self.lengthscale ** 4]])

    def feedback_matrix(self):
        lam = 5.0 ** 0.5 / self.lengthscale
        return np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])


class Matern72(StationaryKernel):
    family, state_dim = _lib.BN_MATERN72, 4

    def K_r(self, r):
        s7 = 7.0 ** 0.5
        return self.variance * (1. + s7 * r + 14. / 5. * r ** 2 + 7. * s7 / 15. * r ** 3) * np.exp(-s7 * r)

    def stationary_covariance(self):
        k1 = 7.0 / 5.0 * self.variance / self.lengthscale ** 2
        k2 = 9.8 * self.variance / self.lengthscale ** 4
        return np.array([[self.variance, 0.0, -k1, 0.0], [0.0, k1, 0.0, -k2], [-k1, 0.0, k2, 0.0],
                         [0.0, -k2, 0.0, 343.0 * self.variance / self.lengthscale ** 6]])

    def feedback_matrix(self):
        lam = 7.0 ** 0.5 / self.lengthscale
        return np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                         [-lam ** 4, -4.0 * lam ** 3, -6.0 * lam ** 2, -4.0 * lam]])


Exponential = Matern12


def _block_diag(mats):
    n = sum(m.shape[0] for m in mats)
    k = sum(m.shape[1] for m in mats)
    out = np.zeros((n, k))
    i = j = 0
    for m in mats:
        out[i:i + m.shape[0], j:j + m.shape[1]] = m
        i, j = i + m.shape[0], j + m.shape[1]
    return out


class Independent(Kernel):
    """stack of independent priors, one latent each (kernels.py:1499-1616)"""

    def __init__(self, kernels):
        self.kernels = list(kernels)
        self.num_kernels = len(self.kernels)

    @property
    def state_dim(self):
        return sum(k.state_dim for k in self.kernels)

    def spec(self):
        fams = {getattr(k, 'family', None) for k in self.kernels}
        if len(fams) == 1 and None not in fams and self.num_kernels <= _lib.BN_MAX_COMPONENTS:
            return _lib.kernel_spec(self.kernels[0].family, [k.variance for k in self.kernels],
                                    [k.lengthscale for k in self.kernels])
        return None

    def measurement_model(self):
        return _block_diag([k.measurement_model() for k in self.kernels])

    def stationary_covariance(self):
        return _block_diag([k.stationary_covariance() for k in self.kernels])

    def feedback_matrix(self):
        return _block_diag([k.feedback_matrix() for k in self.kernels])

    def state_transition(self, dt):
        As, _ = discretise(self, dt)
        return As[0] if np.ndim(dt) == 0 else As


Separate = Independent


def discretise(kernel, dt):
    """(As[N,d,d], Qs[N,d,d]) on the GPU: vmap(state_transition)(dt) and Q = Pinf - A Pinf A^T (ops.py:274-278)"""
    dt = as_dev(np.atleast_1d(dt) if not torch.is_tensor(dt) else dt.reshape(-1)).reshape(-1)
    N = dt.shape[0]
    spec = kernel.spec()
    if spec is not None:
        d = _lib.lib().bn_state_dim(spec)
        As = torch.empty((N, d, d), dtype=torch.float64, device=dt.device)
        Qs = torch.empty((N, d, d), dtype=torch.float64, device=dt.device)
        _lib.check(_lib.lib().bn_discretise(spec, N, dt.data_ptr(), As.data_ptr(), Qs.data_ptr(), stream_ptr()))
        return As, Qs
    if isinstance(kernel, Independent):  # mixed families: discretise per component, assemble the blocks
        d = kernel.state_dim
        As = torch.zeros((N, d, d), dtype=torch.float64, device=dt.device)
        Qs = torch.zeros((N, d, d), dtype=torch.float64, device=dt.device)
        o = 0
        for k in kernel.kernels:
            a, q = discretise(k, dt)
            n = a.shape[1]
            As[:, o:o + n, o:o + n] = a
            Qs[:, o:o + n, o:o + n] = q
            o += n
        return As, Qs
    raise NotImplementedError(
        '%s has no closed-form discretisation compiled into libbn_b200: build As, Qs yourself and call the '
        'array-level entry points (ops._sequential_kf / _parallel_kf / _sequential_rts / _parallel_rts)'
        % type(kernel).__name__)
