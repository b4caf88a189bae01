"""Infinite-horizon (steady-state) Kalman filter / RTS smoother and the model that uses them (TEST INFRASTRUCTURE).

NumPy restatement of bayesnewton/ops.py:796-878 (dare, _sequential_kf_ih, _parallel_kf_ih, kalman_filter_infinite_horizon),
:955-1068 (rts_dare, _sequential_rts_ih, _parallel_rts_ih, rauch_tung_striebel_smoother_infinite_horizon) and of
InfiniteHorizonGaussianProcess (basemodels.py:1257-1300): the state covariance is replaced by the fixed point of the
Riccati recursion for the AVERAGED site precision (20 iterations per call, warm-started from the previous call's
result, :1272-1289), so the filter and the smoother become affine recursions in the mean alone.
Pinned on tests/golden/reference_infinite_horizon.npz (the reference's own code, run on oracle/jaxshim).
"""
import numpy as np

from . import kalman, sites
from .linalg import T, inv, solve
from .model import _Base
from .ssm import process_noise_covariance


def dare(A, H, Q, R, Pinit, num_iters=20):
    """ops.py:796-824"""
    X = Pinit
    for _ in range(num_iters):
        HX = H @ X
        S = HX @ H.T + R
        K = T(solve(S, HX))
        X = A @ (X - K @ HX) @ A.T + Q
    return X


def rts_dare(A, Q, Pinf, num_iters=20):
    """ops.py:955-975"""
    X = Pinf
    for _ in range(num_iters):
        X = A @ X @ A.T + Q
    return X


def kalman_filter_infinite_horizon(dt, kernel, y, noise_cov, mask=None, parallel=False, heteroscedastic=False,
                                   noise_cov_tied=None, dare_iters=20, dare_init=None):
    """ops.py:881-952.  Returns ell, (means [N,d,1], (Pdare, cov)).  The parallel form evaluates the same affine
    recursion m_k = (A - K_k H A) m_{k-1} + K_k y_k by an associative scan; here it is run in time order (the reference's
    heteroscedastic scan multiplies INVERSES of the contractions, ops.py:860, and loses all accuracy on long series)."""
    N = dt.shape[0]
    if mask is None:
        mask = np.zeros_like(y, dtype=bool)
    Pinf = kernel.stationary_covariance()
    minf = np.zeros((Pinf.shape[0], 1))
    A = kernel.state_transition(dt[1])
    Q = process_noise_covariance(A, Pinf)
    H = kernel.measurement_model()
    dare_init = Pinf if dare_init is None else dare_init
    Pdare = dare(A, H, Q, noise_cov_tied, dare_init, dare_iters)
    S = H @ Pdare @ H.T + noise_cov_tied
    K = Pdare @ T(solve(S, H))
    HA = H @ A
    if heteroscedastic:
        Ss = H @ Pdare @ H.T + noise_cov
        Ks = Pdare @ T(solve(Ss, np.tile(H, (N, 1, 1))))
        AKHAs = A - Ks @ HA
    else:
        Ss = np.tile(S, (N, 1, 1))
        Ks = np.tile(K, (N, 1, 1))
        AKHAs = np.tile(A - K @ HA, (N, 1, 1))
    Kys = Ks @ y
    cov = Pdare - K @ H @ Pdare
    m, ell = minf, 0.0
    means = np.zeros((N,) + minf.shape)
    for k in range(N):
        ell = ell + kalman.mvn_logpdf(y[k][None], (HA @ m)[None], Ss[k][None], mask[k][None])[0]
        m = AKHAs[k] @ m + Kys[k]
        means[k] = m
    return ell, (means, (Pdare, cov))


def rauch_tung_striebel_smoother_infinite_horizon(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False,
                                                  dare_iters=20, dare_init=None):
    """ops.py:1018-1068.  Returns means, covs, gains, dare_cov."""
    Pinf = kernel.stationary_covariance()
    A = kernel.state_transition(dt[0])
    H = kernel.measurement_model()
    N = dt.shape[0]
    Pdare, fcov = filter_cov
    gain = fcov @ T(solve(Pdare, A))
    Qdare = fcov - gain @ Pdare @ gain.T
    dare_init = Pinf if dare_init is None else dare_init
    dare_cov = rts_dare(gain, Qdare, dare_init, dare_iters)
    Afms = A @ filter_mean
    sm = filter_mean[-1]
    out = np.zeros_like(filter_mean)
    for k in range(N - 1, -1, -1):   # _sequential_rts_ih, ops.py:978-992: the last step is processed like every other
        sm = filter_mean[k] + gain @ (sm - Afms[k])
        out[k] = sm
    means = out if return_full else H @ out
    cov = dare_cov if return_full else H @ dare_cov @ H.T
    return means, np.tile(cov, (N, 1, 1)), np.tile(gain, (N, 1, 1)), dare_cov


class InfiniteHorizonGP(_Base):
    """InfiniteHorizonGaussianProcess + an inference mixin (basemodels.py:1257-1300), single-latent likelihoods"""

    def __init__(self, *a, dare_iters=20, **kw):
        super().__init__(*a, **kw)
        assert np.max(np.abs(np.diff(self.dt[1:]))) < 1e-6, 'time steps must be equidistant'
        self.heteroscedastic = bool(np.any(np.isnan(self.Y))) or not isinstance(self.likelihood, sites.Gaussian)
        self.dare_iters = dare_iters
        Pinf = self.kernel.stationary_covariance()
        self.dare_init_filter, self.dare_init_smoother = Pinf, Pinf

    def _filter(self):
        tied = inv(np.mean(self.site_nat2, axis=0))
        out = kalman_filter_infinite_horizon(self.dt, self.kernel, self.site_mean, self.site_cov, self._mask3(),
                                             heteroscedastic=self.heteroscedastic, noise_cov_tied=tied,
                                             dare_iters=self.dare_iters, dare_init=self.dare_init_filter)
        self.dare_init_filter = out[1][1][0]
        return out

    def update_posterior(self):
        _, (fm, fcov) = self._filter()
        dts = np.concatenate([self.dt[1:], [0.0]])
        sm, sP, _, dare_cov = rauch_tung_striebel_smoother_infinite_horizon(dts, self.kernel, fm, fcov,
                                                                            dare_iters=self.dare_iters,
                                                                            dare_init=self.dare_init_smoother)
        self.dare_init_smoother = dare_cov
        self.post_mean, self.post_cov = sm, sP

    def compute_log_lik(self):
        return self._filter()[0]
