"""Sparse Markov GP: pairs filter and the inducing-state model (oracle; test infrastructure).

Restates ``kalman_filter_pairs`` (``bayesnewton/ops.py:383-426``), ``build_joint`` / ``set_z_stats``
(``utils.py:544-559``), ``sum_natural_params_by_group`` (``utils.py:218-224``) and the VI iteration of
``SparseMarkovGaussianProcess`` (``basemodels.py:928-1152`` with ``inference.py:65-90,170-222``).
Comparator, as in the reference's ``tests/test_sparsemarkov.py``: with ``Z = X`` the model reproduces the plain
Markov GP.
"""
import numpy as np
from . import kalman, sites
from .linalg import T, inv
from .predict import compute_conditional_statistics
from .ssm import discretise


def kalman_filter_pairs(dt, kernel, y, noise_cov, mask=None, parallel=False, order='tree'):
    """ops.py:383-426"""
    if mask is None:
        mask = np.zeros_like(y, dtype=bool)
    Pinf = kernel.stationary_covariance()
    d = Pinf.shape[0]
    zeros = np.zeros((d, d))
    Pinfpair = np.block([[Pinf, zeros], [zeros, Pinf]])
    minfpair = np.zeros((2 * d, 1))
    As, Qs = discretise(kernel, dt)
    N = As.shape[0]
    Apairs, Qpairs = np.zeros((N, 2 * d, 2 * d)), np.zeros((N, 2 * d, 2 * d))
    Apairs[:, :d, d:] = np.eye(d)
    Apairs[:, d:, d:] = As
    Qpairs[:, :d, :d] = 1e-32 * np.eye(d)
    Qpairs[:, d:, d:] = Qs
    H = np.eye(2 * d)
    if parallel:
        ell, means, covs = kalman.parallel_kf(Apairs, Qpairs, H, y, noise_cov, minfpair, Pinfpair, mask, order=order)
    else:
        ell, means, covs = kalman.sequential_kf(Apairs, Qpairs, H, y, noise_cov, minfpair, Pinfpair, mask)
    return ell, (means[1:, :d], covs[1:, :d, :d])


def build_joint(mean_aug, cov_aug, gain_aug):
    """vmap(build_joint) over all transitions (utils.py:544-553): [Mt,2d,1], [Mt,2d,2d]"""
    Mt = mean_aug.shape[0] - 1
    cross = gain_aug[:Mt] @ cov_aug[1:Mt + 1]
    mean_joint = np.concatenate([mean_aug[:Mt], mean_aug[1:Mt + 1]], axis=1)
    cov_joint = np.concatenate([np.concatenate([cov_aug[:Mt], cross], axis=2),
                                np.concatenate([T(cross), cov_aug[1:Mt + 1]], axis=2)], axis=1)
    return mean_joint, cov_joint


def set_z_stats(t, z_aug):
    """utils.py:556-559"""
    ind = np.searchsorted(z_aug.reshape(-1), np.asarray(t).reshape(-1)) - 1
    num_neighbours = np.array([np.sum(ind == m) for m in range(z_aug.shape[0] - 1)])
    return ind, num_neighbours


class SparseMarkovGP:
    """SparseMarkovVariationalGP, single-latent likelihood, full batch"""

    def __init__(self, kernel, likelihood, X, Y, Z, num_quad_pts=20, parallel=False):
        from .model import input_admin
        self.kernel, self.likelihood, self.num_quad_pts, self.parallel = kernel, likelihood, num_quad_pts, parallel
        self.t, Yh, _ = input_admin(X, Y)
        self.Y = Yh[:, 0]
        self.N = self.t.shape[0]
        self.d = kernel.stationary_covariance().shape[0]
        Z = np.sort(np.asarray(Z, dtype=np.float64).reshape(-1))
        self.Z = np.concatenate([[-1e10], Z, [1e10]])
        self.dz = np.diff(self.Z)
        self.Mt = self.dz.shape[0]
        d, Mt = self.d, self.Mt
        eyes = np.tile(np.eye(2 * d), (Mt, 1, 1))
        nat2 = 1e-8 * eyes
        nat2[:-1, d, d] = 1e-2                      # basemodels.py:954
        self.site_nat2 = nat2
        self.site_nat1 = np.zeros((Mt, 2 * d, 1))
        self.site_mean = np.zeros((Mt, 2 * d, 1))
        self.site_cov = inv(nat2)
        # GaussianDistribution(mean, cov) recomputes the natural parameters from (mean, cov) (basemodels.py:60-64)
        self.site_nat1, self.site_nat2 = sites.reparametrise(self.site_mean, self.site_cov)
        self.post_mean = np.zeros((Mt, 2 * d, 1))
        self.post_cov = eyes.copy()
        self.ind, self.num_neighbours = set_z_stats(self.t, self.Z)

    def _filter(self):
        return kalman_filter_pairs(self.dz, self.kernel, self.site_mean, self.site_cov, parallel=self.parallel)

    def update_posterior(self):
        """basemodels.py:980-1008"""
        _, (fm, fP) = self._filter()
        sm, sP, gain = kalman.rauch_tung_striebel_smoother(self.dz[1:], self.kernel, fm, fP, return_full=True,
                                                           parallel=self.parallel)
        Pinf = self.kernel.stationary_covariance()[None]
        minf = np.zeros((1, self.d, 1))
        mean_aug = np.concatenate([minf, sm, minf])
        cov_aug = np.concatenate([Pinf, sP, Pinf])
        gain_aug = np.concatenate([np.zeros_like(gain[:1]), gain])
        self.smoother_mean, self.smoother_cov, self.gain = sm, sP, gain
        self.post_mean, self.post_cov = build_joint(mean_aug, cov_aug, gain_aug)

    def conditional_posterior_to_data(self):
        """basemodels.py:1071-1104 -> (mean_f [N], var_f [N], W [N,1,2d])"""
        H = self.kernel.measurement_model()
        W = np.zeros((self.N, 1, 2 * self.d))
        nu = np.zeros(self.N)
        for n in range(self.N):
            P, Tm = compute_conditional_statistics(self.t[n], self.Z, self.kernel, self.ind[n])
            W[n] = H @ P
            nu[n] = (H @ Tm @ H.T)[0, 0]
        pm, pV = self.post_mean[self.ind], self.post_cov[self.ind]
        mean_f = (W @ pm)[:, 0, 0]
        var_f = (W @ pV @ T(W))[:, 0, 0] + nu
        return mean_f, var_f, W

    def inference(self, lr=1.0, ensure_psd=True):
        self.update_posterior()
        mean_f, var_f, W = self.conditional_posterior_to_data()
        _, dE, d2E = sites.variational_expectation(self.likelihood, self.Y, mean_f, var_f, self.num_quad_pts)
        d2E = d2E.reshape(-1, 1, 1)
        if ensure_psd:
            d2E = -sites.ensure_diagonal_positive_precision(-d2E)
        jac = T(W) @ dE.reshape(-1, 1, 1)            # conditional_data_to_posterior, basemodels.py:1106-1112
        hess = T(W) @ d2E @ W
        nat1_n, nat2_n = sites.newton_update(self.post_mean[self.ind], jac, hess)
        # group_natural_params (basemodels.py:1114-1138)
        new1, new2, counter = np.zeros_like(self.site_nat1), np.zeros_like(self.site_nat2), np.zeros(self.Mt)
        for n in range(self.N):
            new1[self.ind[n]] += nat1_n[n]
            new2[self.ind[n]] += nat2_n[n]
            counter[self.ind[n]] += 1.0
        frac = (1. - counter / np.maximum(self.num_neighbours, 1)).reshape(-1, 1, 1)
        nat1 = new1 + frac * self.site_nat1
        nat2 = new2 + frac * self.site_nat2 + 1e-8 * np.eye(2 * self.d)
        d1 = np.mean(np.abs(nat1 - self.site_nat1))
        d2 = np.mean(np.abs(nat2 - self.site_nat2))
        self.site_nat1 = (1 - lr) * self.site_nat1 + lr * nat1
        self.site_nat2 = (1 - lr) * self.site_nat2 + lr * nat2
        self.site_mean, self.site_cov = sites.reparametrise(self.site_nat1, self.site_nat2)
        self.update_posterior()
        return d1, d2

    def compute_kl(self):
        ell, _ = self._filter()
        edp = sites.gaussian_expected_log_lik(self.site_mean, self.post_mean, self.post_cov, self.site_cov, None)
        return np.sum(edp) - ell

    def energy(self):
        mean_f, var_f, _ = self.conditional_posterior_to_data()
        E, _, _ = sites.variational_expectation(self.likelihood, self.Y, mean_f, var_f, self.num_quad_pts)
        return -(np.nansum(E) - self.compute_kl())

    def predict(self, X_test):
        """basemodels.py:1033-1069: (mean [N*], var [N*])"""
        from .predict import temporal_conditional
        tm, tc = temporal_conditional(self.Z[1:-1], X_test, self.smoother_mean, self.smoother_cov, self.gain, self.kernel)
        H = self.kernel.measurement_model()
        return (H @ tm)[:, 0, 0], (H @ tc @ H.T)[:, 0, 0]
