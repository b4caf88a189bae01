"""One inference iteration of a temporal Markov GP, plus the dense-GP comparator (oracle; test infrastructure).

``MarkovGP`` restates the host orchestration the reference spreads over
``MarkovGaussianProcess`` (``basemodels.py:625-741``), ``BaseModel``
(``:103-262``), ``InferenceMixin.inference`` (``inference.py:65-90``) and the
VI / EP / Newton / PL ``energy`` methods (``inference.py:130-154,197-222,
286-325,373-428``).  ``DenseGP`` restates ``GaussianProcess``
(``basemodels.py:265-357``) on top of ``gaussian_conditional``
(``ops.py:52-80``): it is the comparator the reference's own tests use
(``tests/test_gp_vs_markovgp_{reg,class}.py``), and ``exact_marginal_likelihood``
is the closed form of ``tests/test_vs_exact_marg_lik.py:41-65``.
"""
import math
import numpy as np
from . import kalman, sites
from .linalg import T, chol, cho_solve, solve

LOG2PI = math.log(2 * math.pi)


def input_admin(t, y):
    """sort by time, dt = [0, diff(t)] (utils.py:234-265, temporal inputs only)"""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(t.shape[0], -1)
    ind = np.argsort(t, kind='stable')
    t, y = t[ind], y[ind]
    dt = np.concatenate([[0.0], np.diff(t)])
    return t, y, dt


class _Base:
    def __init__(self, kernel, likelihood, X, Y, method='vi', power=1.0, num_quad_pts=20):
        self.kernel, self.likelihood = kernel, likelihood
        self.method, self.power, self.num_quad_pts = method, power, num_quad_pts
        self.t, self.Y, self.dt = input_admin(X, Y)
        self.N = self.t.shape[0]
        H = kernel.measurement_model()
        self.func_dim = H.shape[0]
        D = self.func_dim
        # sites: mean 0, covariance 100 I  (basemodels.py:130-133)
        self.site_mean = np.zeros((self.N, D, 1))
        self.site_cov = 1e2 * np.tile(np.eye(D), (self.N, 1, 1))
        self.site_nat1, self.site_nat2 = sites.reparametrise(self.site_mean, self.site_cov)
        self.post_mean = np.zeros((self.N, D, 1))
        self.post_cov = np.tile(np.eye(D), (self.N, 1, 1))
        mask_y = np.isnan(self.Y)
        if D == self.Y.shape[1]:
            self.mask_pseudo_y = mask_y
        elif likelihood.multi_latent:
            self.mask_pseudo_y = None
        else:
            self.mask_pseudo_y = np.tile(mask_y, (1, D))
        self.mask_y = mask_y

    # ---- inference.py:65-90
    def inference(self, lr=1.0, ensure_psd=True):
        self.update_posterior()
        mean, jac, hess = sites.site_statistics(
            self.method, self.likelihood, self.Y, self.post_mean, self.post_cov, self.site_nat1, self.site_nat2,
            power=self.power, ensure_psd=ensure_psd, num_quad_pts=self.num_quad_pts,
            mask_pseudo_y=self.mask_pseudo_y)
        (self.site_nat1, self.site_nat2, self.site_mean, self.site_cov, d1, d2) = sites.damped_site_update(
            self.site_nat1, self.site_nat2, mean, jac, hess, lr)
        self.update_posterior()
        return (mean, jac, hess), (d1, d2)

    def _mask3(self):
        return None if self.mask_pseudo_y is None else self.mask_pseudo_y[..., None]

    def compute_kl(self):
        """basemodels.py:708-724 / :347-356"""
        ll = self.compute_log_lik()
        edp = sites.gaussian_expected_log_lik(self.site_mean, self.post_mean, self.post_cov, self.site_cov,
                                              self._mask3())
        return np.sum(edp) - ll

    def energy(self):
        lik, m, V = self.likelihood, self.post_mean, self.post_cov
        if self.method in ('vi', 'newton'):
            if lik.multi_latent:
                if self.method == 'vi':
                    val, _, _ = sites.variational_expectation_ml(lik, self.Y[:, 0], m[:, :, 0], V, self.num_quad_pts)
                else:
                    val, _, _ = sites.log_likelihood_gradients_ml(lik, self.Y[:, 0], m[:, :, 0])
            elif self.method == 'vi':
                val, _, _ = sites.variational_expectation(lik, self.Y[:, 0], m[:, 0, 0], V[:, 0, 0], self.num_quad_pts)
            else:
                val, _, _ = sites.log_likelihood_gradients(lik, self.Y[:, 0], m[:, 0, 0])
            return -(np.nansum(val) - self.compute_kl())
        # EP (inference.py:286-325) and PL, which uses the EP energy at power 1 (:373-428)
        power = self.power if self.method == 'ep' else 1.0
        cm, cV = sites.compute_cavity(m, V, self.site_nat1, self.site_nat2, power)
        if lik.multi_latent:
            lel, _, _ = sites.moment_match_ml(lik, self.Y[:, 0], cm[:, :, 0], cV, power, self.num_quad_pts)
        else:
            lel, _, _ = sites.moment_match(lik, self.Y[:, 0], cm[:, 0, 0], cV[:, 0, 0], power, self.num_quad_pts)
            lel = np.where(self.mask_y[:, 0], 0., lel)
        lel_pseudo = kalman.mvn_logpdf(self.site_mean, cm, self.site_cov / power + cV, self._mask3())
        if self.method == 'ep':  # PEP constant (basemodels.py:259); PL omits it (inference.py:406-411)
            D = self.func_dim
            Lc = chol(self.site_cov)
            logd = np.log(np.abs(np.diagonal(Lc, axis1=-2, axis2=-1)))
            dim = np.full(self.N, float(D))
            if self.mask_pseudo_y is not None:
                logd = np.where(self.mask_pseudo_y, 0., logd)
                dim = dim - np.sum(self.mask_pseudo_y, axis=1)
            lel_pseudo = lel_pseudo + 0.5 * dim * ((1 - power) * LOG2PI - np.log(power)) \
                + 0.5 * (1 - power) * 2 * np.sum(logd, axis=1)
        lZ = self.compute_log_lik()
        return -(lZ + 1. / power * (np.nansum(lel) - np.nansum(lel_pseudo)))


class MarkovGP(_Base):
    def __init__(self, *a, parallel=False, order='tree', **kw):
        super().__init__(*a, **kw)
        self.parallel, self.order = parallel, order

    def update_posterior(self):
        """basemodels.py:689-706"""
        ell, (fm, fP) = kalman.kalman_filter(self.dt, self.kernel, self.site_mean, self.site_cov,
                                             self._mask3(), parallel=self.parallel, order=self.order)
        dts = np.concatenate([self.dt[1:], [0.0]])
        sm, sP, _ = kalman.rauch_tung_striebel_smoother(dts, self.kernel, fm, fP, parallel=self.parallel,
                                                        order=self.order)
        self.filter_mean, self.filter_cov = fm, fP
        self.post_mean, self.post_cov = sm, sP

    def compute_log_lik(self):
        """basemodels.py:726-741"""
        ell, _ = kalman.kalman_filter(self.dt, self.kernel, self.site_mean, self.site_cov, self._mask3(),
                                      parallel=self.parallel, order=self.order)
        return ell


class DenseGP(_Base):
    """O(N^3) comparator on the covariance function of the same kernel"""

    def _Ky(self):
        D = self.func_dim
        K = self.kernel.K(self.t, self.t)
        Rbd = np.zeros_like(K)
        for n in range(self.N):
            Rbd[n * D:(n + 1) * D, n * D:(n + 1) * D] = self.site_cov[n]
        return K, K + Rbd

    def update_posterior(self):
        """ops.py:52-80, basemodels.py:287-300"""
        D = self.func_dim
        K, Ky = self._Ky()
        KiKy = solve(Ky, K).T
        mean = KiKy @ self.site_mean.reshape(-1, 1)
        cov = K - KiKy @ K
        self.post_mean = mean.reshape(self.N, D, 1)
        self.post_cov = np.stack([cov[n * D:(n + 1) * D, n * D:(n + 1) * D] for n in range(self.N)])

    def compute_log_lik(self):
        """basemodels.py:302-329"""
        _, Ky = self._Ky()
        y = self.site_mean.reshape(-1, 1)
        L = chol(Ky)
        return (-0.5 * np.sum(y.T @ cho_solve(L, y)) - np.sum(np.log(np.diag(L)))
                - 0.5 * y.shape[0] * LOG2PI)


def exact_marginal_likelihood(kernel, noise_var, t, y):
    """-log N(y | 0, K + s2 I)  (tests/test_vs_exact_marg_lik.py:41-65)"""
    t = np.asarray(t).reshape(-1)
    y = np.asarray(y).reshape(-1, 1)
    Ky = kernel.K(t, t) + noise_var * np.eye(t.shape[0])
    L = np.linalg.cholesky(Ky)
    a = np.linalg.solve(L.T, np.linalg.solve(L, y))
    return float(0.5 * (y.T @ a)[0, 0] + np.sum(np.log(np.diag(L))) + 0.5 * t.shape[0] * LOG2PI)
