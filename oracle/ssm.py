"""State-space (SDE) form of the Matern family + Independent stacking (oracle; test infrastructure).

Restates the discretisation generators the reference's filter consumes:
``state_transition`` / ``stationary_covariance`` / ``measurement_model`` of
Matern12/32/52/72 (``bayesnewton/kernels.py:123-382``) and ``Independent``
(``kernels.py:1499-1616``), the process noise ``Q = Pinf - A Pinf A^T``
(``ops.py:149-151``) and the covariance functions ``K_r`` used by the dense-GP
comparator (``kernels.py:97-104,139,191-193,249-251,320-322``).
Hyper-parameters are held untransformed; ``softplus`` / ``softplus_inv``
(``utils.py:54-74``) are provided for gradients w.r.t. the stored variables.
"""
import numpy as np
from .linalg import block_diag, T


def softplus(x):
    return np.log(1.0 + np.exp(x))  # the naive form, utils.py:54-55


def softplus_inv(x):
    return np.log(np.exp(x) - 1.0)  # utils.py:67-74


class _Stationary:
    state_dim = None

    def __init__(self, variance=1.0, lengthscale=1.0, dtype=np.float64):
        self.variance = dtype(variance)
        self.lengthscale = dtype(lengthscale)
        self.dtype = dtype

    def measurement_model(self):
        H = np.zeros((1, self.state_dim), dtype=self.dtype)
        H[0, 0] = 1.0
        return H

    def K(self, X, X2):
        # scaled distance with the 1e-36 clip of kernels.py:97-104
        X = np.asarray(X, dtype=self.dtype).reshape(-1, 1) / self.lengthscale
        X2 = np.asarray(X2, dtype=self.dtype).reshape(-1, 1) / self.lengthscale
        r2 = (X - X2.T) ** 2
        return self.K_r(np.sqrt(np.maximum(r2, 1e-36)))


class Matern12(_Stationary):
    state_dim = 1

    def K_r(self, r):
        return self.variance * np.exp(-r)

    def stationary_covariance(self):
        return np.array([[self.variance]], dtype=self.dtype)

    def state_transition(self, dt):
        return np.exp(-dt / self.lengthscale) * np.ones((1, 1), dtype=self.dtype)


class Matern32(_Stationary):
    state_dim = 2

    def K_r(self, r):
        s3 = np.sqrt(self.dtype(3.0))
        return self.variance * (1.0 + s3 * r) * np.exp(-s3 * r)

    def stationary_covariance(self):
        return np.array([[self.variance, 0.0],
                         [0.0, 3.0 * self.variance / self.lengthscale ** 2]], dtype=self.dtype)

    def state_transition(self, dt):
        lam = np.sqrt(self.dtype(3.0)) / self.lengthscale
        M = np.array([[lam, 1.0], [-lam ** 2, -lam]], dtype=self.dtype)
        return np.exp(-dt * lam) * (dt * M + np.eye(2, dtype=self.dtype))


class Matern52(_Stationary):
    state_dim = 3

    def K_r(self, r):
        s5 = np.sqrt(self.dtype(5.0))
        return self.variance * (1.0 + s5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-s5 * r)

    def stationary_covariance(self):
        kappa = 5.0 / 3.0 * self.variance / self.lengthscale ** 2
        return np.array([[self.variance, 0.0, -kappa],
                         [0.0, kappa, 0.0],
                         [-kappa, 0.0, 25.0 * self.variance / self.lengthscale ** 4]], dtype=self.dtype)

    def state_transition(self, dt):
        lam = np.sqrt(self.dtype(5.0)) / self.lengthscale
        dl = dt * lam
        M = np.array([[lam * (0.5 * dl + 1.0), dl + 1.0, 0.5 * dt],
                      [-0.5 * dl * lam ** 2, lam * (1.0 - dl), 1.0 - 0.5 * dl],
                      [lam ** 3 * (0.5 * dl - 1.0), lam ** 2 * (dl - 3), lam * (0.5 * dl - 2.0)]], dtype=self.dtype)
        return np.exp(-dl) * (dt * M + np.eye(3, dtype=self.dtype))


class Matern72(_Stationary):
    state_dim = 4

    def K_r(self, r):
        s7 = np.sqrt(self.dtype(7.0))
        return self.variance * (1. + s7 * r + 14. / 5. * r ** 2 + 7. * s7 / 15. * r ** 3) * np.exp(-s7 * r)

    def stationary_covariance(self):
        k1 = 7.0 / 5.0 * self.variance / self.lengthscale ** 2
        k2 = 9.8 * self.variance / self.lengthscale ** 4
        return np.array([[self.variance, 0.0, -k1, 0.0],
                         [0.0, k1, 0.0, -k2],
                         [-k1, 0.0, k2, 0.0],
                         [0.0, -k2, 0.0, 343.0 * self.variance / self.lengthscale ** 6]], dtype=self.dtype)

    def state_transition(self, dt):
        lam = np.sqrt(self.dtype(7.0)) / self.lengthscale
        l2, l3 = lam * lam, lam * lam * lam
        dl = dt * lam
        dl2 = dl ** 2
        M = np.array([
            [lam * (1.0 + 0.5 * dl + dl2 / 6.0), 1.0 + dl + 0.5 * dl2, 0.5 * dt * (1.0 + dl), dt ** 2 / 6],
            [-dl2 * lam ** 2 / 6.0, lam * (1.0 + 0.5 * dl - 0.5 * dl2), 1.0 + dl - 0.5 * dl2, dt * (0.5 - dl / 6.0)],
            [l3 * dl * (dl / 6.0 - 0.5), dl * l2 * (0.5 * dl - 2.0), lam * (1.0 - 2.5 * dl + 0.5 * dl2),
             1.0 - dl + dl2 / 6.0],
            [l2 ** 2 * (dl - 1.0 - dl2 / 6.0), l3 * (3.5 * dl - 4.0 - 0.5 * dl2), l2 * (4.0 * dl - 6.0 - 0.5 * dl2),
             lam * (1.5 * dl - 3.0 - dl2 / 6.0)]], dtype=self.dtype)
        return np.exp(-dl) * (dt * M + np.eye(4, dtype=self.dtype))


class Independent:
    """block-diagonal stack of priors, one latent per component (kernels.py:1499-1616)"""

    def __init__(self, kernels):
        self.kernels = list(kernels)
        self.dtype = self.kernels[0].dtype

    @property
    def state_dim(self):
        return sum(k.state_dim for k in self.kernels)

    def measurement_model(self):
        return block_diag(*[k.measurement_model() for k in self.kernels])

    def stationary_covariance(self):
        return block_diag(*[k.stationary_covariance() for k in self.kernels])

    def state_transition(self, dt):
        return block_diag(*[k.state_transition(dt) for k in self.kernels])

    def K(self, X, X2):
        # kron(K_i, e_i e_i^T) summed: latent index fastest (kernels.py:1512-1521)
        n = len(self.kernels)
        out = 0.0
        for i, k in enumerate(self.kernels):
            sel = np.zeros((n, n), dtype=self.dtype)
            sel[i, i] = 1.0
            out = out + np.kron(k.K(X, X2), sel)
        return out


def process_noise_covariance(A, Pinf):
    """ops.py:149-151"""
    return Pinf - A @ Pinf @ T(A)


def discretise(kernel, dt):
    """As[N,d,d], Qs[N,d,d] exactly as kalman_filter builds them (ops.py:274-278)"""
    dt = np.asarray(dt).reshape(-1)
    Pinf = kernel.stationary_covariance()
    As = np.stack([kernel.state_transition(d) for d in dt])
    Qs = Pinf - As @ Pinf @ T(As)
    return As, Qs
