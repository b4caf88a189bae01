"""Spatio-temporal Markov GP (oracle; test infrastructure -- only tests/, smoke() and bench.py's CPU leg import it).

NumPy restatement of the reference's spatio-temporal branch:
  ``Separable``                      kernels.py:1666-1683   product of 1-D kernels, one per spatial dimension
  ``SpatioTemporalKernel``           kernels.py:385-586     K, inducing_precision, spatial_conditional, the
                                                            Kronecker state-space model (I_M (x) temporal model)
  ``SpatioTemporalMarkovGP``         basemodels.py:625-764  compute_full_pseudo_lik (:676-687), update_posterior
                                                            (:689-706), compute_kl (:708-724),
                                                            conditional_posterior_to_data (:743-764)
                                     inference.py:65-90,170-222   VI iteration and energy
                                     likelihoods.py:363-383       per-observation variational expectation + NaN mask
  ``DenseSpatioTemporalGP``          basemodels.py:265-357 on K = k_t * k_s: the comparator of the reference's own
                                     test (tests/test_gp_vs_markovgp_spacetime.py:38-88)
Everything is explicit loops over time steps and dense d x d algebra: O(N_t d^3), small cases only.
"""
import math
import numpy as np
from . import kalman, sites
from .linalg import T, chol, cho_solve, solve, inv

LOG2PI = math.log(2 * math.pi)


class Separable:
    """kernels.py:1666-1683"""

    def __init__(self, kernels):
        self.kernels = list(kernels)

    def K(self, X, X2):
        X, X2 = np.asarray(X, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        out = self.kernels[0].K(X[:, :1], X2[:, :1])
        for i in range(1, len(self.kernels)):
            out = out * self.kernels[i].K(X[:, i:i + 1], X2[:, i:i + 1])
        return out

    __call__ = K


class _Spatial1D:
    """a plain stationary kernel used as the spatial kernel of 1-D space (Kernel.__call__ = K, kernels.py:76-78)"""

    def __init__(self, kernel):
        self.kernel = kernel

    def K(self, X, X2):
        return self.kernel.K(np.asarray(X)[:, :1], np.asarray(X2)[:, :1])

    __call__ = K


class SpatioTemporalKernel:
    """kernels.py:385-586 (sparse=True branch; conditional in {'Full', 'DTC', 'FIC'})"""

    def __init__(self, temporal_kernel, spatial_kernel, z, conditional='Full', sparse=True):
        self.temporal_kernel = temporal_kernel
        self.spatial_kernel = spatial_kernel if hasattr(spatial_kernel, 'kernels') else _Spatial1D(spatial_kernel)
        z = np.asarray(z, dtype=np.float64)
        self.z = z[:, None] if z.ndim < 2 else z
        self.M = self.z.shape[0]
        self.conditional = conditional.lower()
        self.sparse = sparse
        assert sparse, 'only the sparse branch is restated'

    @property
    def state_dim(self):
        return self.temporal_kernel.state_dim

    def K(self, X, X2):
        """kernels.py:463-468: product of the temporal and the spatial covariance"""
        X, X2 = np.asarray(X, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        return self.temporal_kernel.K(X[:, :1], X2[:, :1]) * self.spatial_kernel.K(X[:, 1:], X2[:, 1:])

    def inducing_precision(self):
        """kernels.py:508-515"""
        Kzz = self.spatial_kernel.K(self.z, self.z)
        Lzz = chol(Kzz)
        Qzz = cho_solve(Lzz, np.eye(self.M))
        return Qzz, Lzz

    def conditional_covariance(self, t, R, Krz, K):
        """kernels.py:470-484; t is one time stamp, so temporal_kernel.K(t, t) is the 1 x 1 prior variance"""
        if self.conditional == 'dtc':
            return np.array([[0.0]])
        Krr = self.spatial_kernel.K(R, R)
        kt = self.temporal_kernel.K(np.reshape(t, (-1, 1)), np.reshape(t, (-1, 1)))
        resid = Krr - K @ Krz.T
        if self.conditional in ('fic', 'fitc'):
            resid = np.diag(np.diag(resid))
        return kt * resid

    def spatial_conditional(self, X, R):
        """kernels.py:486-506: f(X,R) | u(t) ~ N(B u(t), C), per time step (the reference vmaps over time)"""
        Qzz, Lzz = self.inducing_precision()
        R = np.asarray(R, dtype=np.float64)
        R = R.reshape((R.shape[0], -1) + self.z.shape[1:])
        Bs, Cs = [], []
        for k in range(R.shape[0]):
            Krz = self.spatial_kernel.K(R[k], self.z)
            K = Krz @ Qzz
            Bs.append(K @ Lzz)
            Cs.append(self.conditional_covariance(np.asarray(X).reshape(-1)[k], R[k], Krz, K))
        return np.stack(Bs), np.stack(Cs)

    def stationary_covariance(self):
        return np.kron(np.eye(self.M), self.temporal_kernel.stationary_covariance())  # kernels.py:517-524

    def measurement_model(self):
        return np.kron(np.eye(self.M), self.temporal_kernel.measurement_model())  # kernels.py:534-541

    def state_transition(self, dt):
        return np.kron(np.eye(self.M), self.temporal_kernel.state_transition(dt))  # kernels.py:543-551


def st_input_admin(t, Y, R):
    """utils.py:234-265 with spatial inputs: sort by time; Y [N_t, N_s], R [N_t, N_s, n_spatial_dims]"""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    Y = np.asarray(Y, dtype=np.float64).reshape(t.shape[0], -1)
    R = np.asarray(R, dtype=np.float64)
    R = R.reshape(t.shape[0], Y.shape[1], -1)
    ind = np.argsort(t, kind='stable')
    t, Y, R = t[ind], Y[ind], R[ind]
    dt = np.concatenate([[0.0], np.diff(t)])
    return t, Y, R, dt


def variational_expectation_multi(lik, y, m, cov, num_quad_pts=20):
    """Likelihood.variational_expectation for one time step with several observations (likelihoods.py:363-383):
    y [Ns], m [Ns,1], cov [Ns,Ns] -> (E [Ns], dE [Ns,1], d2E [Ns,Ns] diagonal, NaN where y is missing)"""
    E, dE, d2E = sites.variational_expectation(lik, y, m.reshape(-1), np.diag(cov), num_quad_pts)
    return E, dE.reshape(-1, 1), np.diag(d2E)


class SpatioTemporalMarkovGP:
    """MarkovVariationalGP with a SpatioTemporalKernel (VI only)"""

    def __init__(self, kernel, likelihood, t, Y, R, num_quad_pts=20):
        self.kernel, self.likelihood, self.num_quad_pts = kernel, likelihood, num_quad_pts
        self.t, self.Y, self.R, self.dt = st_input_admin(t, Y, R)
        self.N, self.Ns = self.Y.shape
        self.M = kernel.M
        Ns = self.Ns
        self.site_mean = np.zeros((self.N, Ns, 1))                    # basemodels.py:130-133 (pseudo_lik_size = obs_dim)
        self.site_cov = 1e2 * np.tile(np.eye(Ns), (self.N, 1, 1))
        self.site_nat1, self.site_nat2 = sites.reparametrise(self.site_mean, self.site_cov)
        self.post_mean = np.zeros((self.N, self.M, 1))
        self.post_cov = np.tile(np.eye(self.M), (self.N, 1, 1))
        self.mask_y = np.isnan(self.Y)
        # basemodels.py:136-137 then :652-653: a mask on the pseudo observations only when func_dim == obs_dim
        self.mask_pseudo_y = self.mask_y if self.M == Ns else None
        self.B, self.C = kernel.spatial_conditional(self.t, self.R)

    def _mask3(self):
        return None if self.mask_pseudo_y is None else self.mask_pseudo_y[..., None]

    def compute_full_pseudo_lik(self):
        """basemodels.py:676-687"""
        nat1_full = T(self.B) @ self.site_nat1
        nat2_full = T(self.B) @ self.site_nat2 @ self.B
        pseudo_var = inv(nat2_full + 1e-12 * np.eye(self.M))
        return pseudo_var @ nat1_full, pseudo_var

    def update_posterior(self):
        """basemodels.py:689-706 (sequential filter and smoother)"""
        py, pv = self.compute_full_pseudo_lik()
        ell, (fm, fP) = kalman.kalman_filter(self.dt, self.kernel, py, pv, self._mask3())
        dts = np.concatenate([self.dt[1:], [0.0]])
        sm, sP, _ = kalman.rauch_tung_striebel_smoother(dts, self.kernel, fm, fP)
        self.filter_mean, self.filter_cov = fm, fP
        self.post_mean, self.post_cov = sm, sP
        return ell

    def conditional_posterior_to_data(self):
        """basemodels.py:743-764"""
        return self.B @ self.post_mean, self.B @ self.post_cov @ T(self.B) + self.C

    def _variational_expectation(self):
        mean_f, cov_f = self.conditional_posterior_to_data()
        E = np.zeros((self.N, self.Ns))
        dE = np.zeros((self.N, self.Ns, 1))
        d2E = np.zeros((self.N, self.Ns, self.Ns))
        for k in range(self.N):
            E[k], dE[k], d2E[k] = variational_expectation_multi(self.likelihood, self.Y[k], mean_f[k], cov_f[k],
                                                                self.num_quad_pts)
        return mean_f, E, dE, d2E

    def inference(self, lr=1.0, ensure_psd=True):
        """inference.py:65-90 with VariationalInference.update_variational_params (:170-195)"""
        self.update_posterior()
        mean_f, _, dE, d2E = self._variational_expectation()
        if ensure_psd:
            d2E = -sites.ensure_diagonal_positive_precision(-d2E)
        out = sites.damped_site_update(self.site_nat1, self.site_nat2, mean_f, dE, d2E, lr)
        self.site_nat1, self.site_nat2, self.site_mean, self.site_cov, d1, d2 = out
        self.update_posterior()
        return (mean_f, dE, d2E), (d1, d2)

    def compute_log_lik(self):
        py, pv = self.compute_full_pseudo_lik()
        ell, _ = kalman.kalman_filter(self.dt, self.kernel, py, pv, self._mask3())
        return ell

    def compute_kl(self):
        """basemodels.py:708-724"""
        py, pv = self.compute_full_pseudo_lik()
        ell, _ = kalman.kalman_filter(self.dt, self.kernel, py, pv, self._mask3())
        edp = sites.gaussian_expected_log_lik(py, self.post_mean, self.post_cov, pv, self._mask3())
        return np.sum(edp) - ell

    def energy(self):
        """inference.py:197-222"""
        _, E, _, _ = self._variational_expectation()
        return -(np.nansum(E) - self.compute_kl())


def markov_predict(model, X_test, R_test=None):
    """MarkovGaussianProcess.predict (basemodels.py:766-816) for a SpatioTemporalMarkovGP: latent mean and variance at
    test times X_test [N*] and spatial inputs R_test [N_s*, n_dims] (default: the training ones) -> ([N*, N_s*], same)"""
    from .predict import temporal_conditional
    X_test = np.asarray(X_test, dtype=np.float64).reshape(-1)
    R_test = model.R[0] if R_test is None else np.asarray(R_test, dtype=np.float64)
    k = model.kernel
    py, pv = model.compute_full_pseudo_lik()
    _, (fm, fP) = kalman.kalman_filter(model.dt, k, py, pv, model._mask3())
    dts = np.concatenate([model.dt[1:], [0.0]])
    sm, sP, gain = kalman.rauch_tung_striebel_smoother(dts, k, fm, fP, return_full=True)
    state_mean, state_cov = temporal_conditional(model.t, X_test, sm, sP, gain, k)
    H = k.measurement_model()
    B, C = k.spatial_conditional(X_test, np.tile(R_test[None], (X_test.shape[0], 1, 1)))
    W = B @ H
    mean = (W @ state_mean)[..., 0]
    var = np.diagonal(W @ state_cov @ T(W) + C, axis1=1, axis2=2)
    return mean, var


class DenseSpatioTemporalGP:
    """VariationalGP on the flattened space-time inputs with K = k_t * k_s (basemodels.py:265-357, ops.py:52-80),
    Gaussian sites one per observation; no missing data."""

    def __init__(self, kernel, likelihood, t, Y, R, num_quad_pts=20):
        self.kernel, self.likelihood, self.num_quad_pts = kernel, likelihood, num_quad_pts
        t, Y, R, _ = st_input_admin(t, Y, R)
        Nt, Ns = Y.shape
        self.X = np.concatenate([np.repeat(t, Ns)[:, None], R.reshape(Nt * Ns, -1)], axis=1)
        self.y = Y.reshape(-1)
        n = self.y.shape[0]
        self.site_mean, self.site_var = np.zeros(n), 1e2 * np.ones(n)
        self.site_nat1, self.site_nat2 = np.zeros(n), 1e-2 * np.ones(n)
        self.K = kernel.K(self.X, self.X)

    def update_posterior(self):
        Ky = self.K + np.diag(self.site_var)
        KiKy = solve(Ky, self.K).T
        self.post_mean = KiKy @ self.site_mean
        self.post_var = np.diag(self.K - KiKy @ self.K).copy()

    def inference(self, lr=1.0):
        self.update_posterior()
        _, dE, d2E = sites.variational_expectation(self.likelihood, self.y, self.post_mean, self.post_var, self.num_quad_pts)
        d2E = -np.where(-d2E < 0, 1e-2, -d2E)
        nat1_n, nat2_n = dE - d2E * self.post_mean, -d2E
        self.site_nat1 = (1 - lr) * self.site_nat1 + lr * nat1_n
        self.site_nat2 = (1 - lr) * self.site_nat2 + lr * nat2_n
        self.site_mean, self.site_var = self.site_nat1 / self.site_nat2, 1. / self.site_nat2
        self.update_posterior()

    def energy(self):
        E, _, _ = sites.variational_expectation(self.likelihood, self.y, self.post_mean, self.post_var, self.num_quad_pts)
        Ky = self.K + np.diag(self.site_var)
        L = chol(Ky)
        y = self.site_mean.reshape(-1, 1)
        log_lik = -0.5 * np.sum(y.T @ cho_solve(L, y)) - np.sum(np.log(np.diag(L))) - 0.5 * y.shape[0] * LOG2PI
        edp = (-0.5 * LOG2PI - 0.5 * np.log(self.site_var)
               - 0.5 * ((self.site_mean - self.post_mean) ** 2 + self.post_var) / self.site_var)
        return -(np.sum(E) - (np.sum(edp) - log_lik))


# ------------------------------------------------------------------------------------------- mean-field variants
def _block_diag_dense(Pb):
    """build_block_diag (ops.py:431-433): [M,n,n] blocks -> dense [M n, M n]"""
    M, n = Pb.shape[0], Pb.shape[1]
    P = np.zeros((M * n, M * n))
    for i in range(M):
        P[i * n:(i + 1) * n, i * n:(i + 1) * n] = Pb[i]
    return P


def _get_block_cov(P, M, n):
    """get_block_cov (ops.py:438-439)"""
    return np.stack([P[i * n:(i + 1) * n, i * n:(i + 1) * n] for i in range(M)])


def kalman_filter_meanfield(dt, kernel, y, noise_cov, mask=None):
    """kalman_filter_meanfield -> _sequential_kf_mf (ops.py:581-611, 429-467): the state covariance is truncated to its
    M diagonal n x n blocks after every update.  Returns ell, (means [N,M,n,1], covs [N,M,n,n])."""
    if mask is None:
        mask = np.zeros_like(y, dtype=bool)
    kt, M = kernel.temporal_kernel, kernel.M
    Pinf_t = kt.stationary_covariance()
    n = Pinf_t.shape[0]
    H = kernel.measurement_model()
    N = y.shape[0]
    m = np.zeros((M, n, 1))
    P = np.tile(Pinf_t, (M, 1, 1))
    ell = 0.0
    fms, fPs = np.zeros((N, M, n, 1)), np.zeros((N, M, n, n))
    for k in range(N):
        A = kt.state_transition(dt[k])
        Q = Pinf_t - A @ Pinf_t @ A.T
        mv = (A @ m).reshape(-1, 1)
        Pd = _block_diag_dense(A @ P @ A.T + Q)
        obs_mean = H @ mv
        HP = H @ Pd
        S = HP @ H.T + noise_cov[k]
        ell = ell + kalman.mvn_logpdf(y[k], obs_mean, S, mask[k])
        K = solve(S, HP).T
        m = (mv + K @ (y[k] - obs_mean)).reshape(M, n, 1)
        P = _get_block_cov(Pd - K @ HP, M, n)
        fms[k], fPs[k] = m, P
    return ell, (fms, fPs)


def rts_smoother_meanfield(dt, kernel, fms, fPs, return_full=False):
    """rauch_tung_striebel_smoother_meanfield -> _sequential_rts_mf (ops.py:681-706, 614-650): an independent RTS pass
    per block.  return_full=False: (H sm [N,M,1], H blockdiag(sP) H^T [N,M,M], gains [N,M,n,n] as blocks);
    return_full=True: (sm [N,M n,1], sP as blocks [N,M,n,n], gains as blocks)."""
    kt, M = kernel.temporal_kernel, kernel.M
    Pinf_t = kt.stationary_covariance()
    n = Pinf_t.shape[0]
    H = kernel.measurement_model()
    N = fms.shape[0]
    sm, sP = fms[-1], fPs[-1]
    means = np.zeros((N, M * n if return_full else M, 1))
    covs = np.zeros((N, M, n, n)) if return_full else np.zeros((N, M, M))
    gains = np.zeros((N, M, n, n))
    for k in range(N - 1, -1, -1):
        A = kt.state_transition(dt[k])
        Q = Pinf_t - A @ Pinf_t @ A.T
        pm = A @ fms[k]
        AfP = A @ fPs[k]
        pP = AfP @ A.T + Q
        C = T(solve(pP, AfP))
        sm = fms[k] + C @ (sm - pm)
        sP = fPs[k] + C @ (sP - pP) @ T(C)
        gains[k] = C
        if return_full:
            means[k], covs[k] = sm.reshape(-1, 1), sP
        else:
            means[k], covs[k] = H @ sm.reshape(-1, 1), H @ _block_diag_dense(sP) @ H.T
    return means, covs, gains


class SpatioTemporalMeanFieldMarkovGP(SpatioTemporalMarkovGP):
    """MarkovVariationalMeanFieldGP (models.py:154, basemodels.py:1155-1175): the same model with the mean-field
    filter and smoother"""

    def update_posterior(self):
        py, pv = self.compute_full_pseudo_lik()
        ell, (fm, fP) = kalman_filter_meanfield(self.dt, self.kernel, py, pv, self._mask3())
        dts = np.concatenate([self.dt[1:], [0.0]])
        sm, sP, _ = rts_smoother_meanfield(dts, self.kernel, fm, fP)
        self.filter_mean, self.filter_cov = fm, fP
        self.post_mean, self.post_cov = sm, sP
        return ell

    def compute_log_lik(self):
        py, pv = self.compute_full_pseudo_lik()
        return kalman_filter_meanfield(self.dt, self.kernel, py, pv, self._mask3())[0]

    def compute_kl(self):
        py, pv = self.compute_full_pseudo_lik()
        ell = kalman_filter_meanfield(self.dt, self.kernel, py, pv, self._mask3())[0]
        edp = sites.gaussian_expected_log_lik(py, self.post_mean, self.post_cov, pv, self._mask3())
        return np.sum(edp) - ell
