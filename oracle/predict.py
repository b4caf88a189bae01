"""Prediction at test inputs (oracle; test infrastructure).

Restates ``compute_conditional_statistics`` (``bayesnewton/utils.py:173-215``), ``predict_from_state`` (``:99-120``),
``temporal_conditional`` (``:122-136``), the tail of ``MarkovGaussianProcess.predict`` (``basemodels.py:766-816``),
``predict_cubature`` (``cubature.py:438-465``) and the dense-GP predictive (``ops.py:52-80`` with test inputs,
``basemodels.py:331-345``) used as the comparator, as in ``tests/test_gp_vs_markovgp_reg.py``.
"""
import numpy as np
from . import kalman, sites
from .linalg import chol, cho_solve, solve


def compute_conditional_statistics(x_test, x, kernel, ind):
    """utils.py:173-215 for one test point"""
    dt_fwd = x_test - x[ind]
    dt_back = x[ind + 1] - x_test
    A_fwd = kernel.state_transition(dt_fwd)
    A_back = kernel.state_transition(dt_back)
    Pinf = kernel.stationary_covariance()
    Q_fwd = Pinf - A_fwd @ Pinf @ A_fwd.T
    Q_back = Pinf - A_back @ Pinf @ A_back.T
    A_back_Q_fwd = A_back @ Q_fwd
    Q_mp = Q_back + A_back @ A_back_Q_fwd.T
    L = chol(Q_mp + 1e-8 * np.eye(Q_mp.shape[0]))
    Q_mp_inv_A_back = cho_solve(L, A_back)
    T = Q_fwd - A_back_Q_fwd.T @ Q_mp_inv_A_back @ Q_fwd
    W = Q_fwd @ Q_mp_inv_A_back.T
    P = np.concatenate([A_fwd - W @ A_back @ A_fwd, W], axis=-1)
    return P, T


def temporal_conditional(X, X_test, mean, cov, gain, kernel):
    """utils.py:122-136 with the dummy states of basemodels.py:793-794; X [N] sorted, mean [N,d,1], cov/gain [N,d,d]"""
    X = np.asarray(X, dtype=np.float64).reshape(-1)
    X_test = np.asarray(X_test, dtype=np.float64).reshape(-1)
    X_aug = np.concatenate([[-1e10], X, [1e10]])
    Pinf = kernel.stationary_covariance()[None]
    minf = np.zeros((1, Pinf.shape[1], 1))
    mean_aug = np.concatenate([minf, mean, minf])
    cov_aug = np.concatenate([Pinf, cov, Pinf])
    gain_aug = np.concatenate([np.zeros_like(gain[:1]), gain])
    ind_test = np.searchsorted(X_aug, X_test) - 1
    d = Pinf.shape[1]
    tm, tc = np.zeros((X_test.shape[0], d, 1)), np.zeros((X_test.shape[0], d, d))
    for n, (xt, ind) in enumerate(zip(X_test, ind_test)):
        P, T = compute_conditional_statistics(xt, X_aug, kernel, ind)
        mean_joint = np.concatenate([mean_aug[ind], mean_aug[ind + 1]])
        cross = gain_aug[ind] @ cov_aug[ind + 1]
        cov_joint = np.block([[cov_aug[ind], cross], [cross.T, cov_aug[ind + 1]]])
        tm[n], tc[n] = P @ mean_joint, P @ cov_joint @ P.T + T
    return tm, tc


def markov_predict(model, X_test):
    """MarkovGaussianProcess.predict (basemodels.py:766-816) for an oracle.model.MarkovGP: (mean [N*,Df], var [N*,Df,Df])"""
    _, (fm, fP) = kalman.kalman_filter(model.dt, model.kernel, model.site_mean, model.site_cov, model._mask3())
    dts = np.concatenate([model.dt[1:], [0.0]])
    sm, sP, gain = kalman.rauch_tung_striebel_smoother(dts, model.kernel, fm, fP, return_full=True)
    tm, tc = temporal_conditional(model.t, X_test, sm, sP, gain, model.kernel)
    H = model.kernel.measurement_model()
    return (H @ tm)[..., 0], H @ tc @ H.T


def dense_predict(model, X_test):
    """GaussianProcess.predict (basemodels.py:331-345, ops.py:52-80): marginals of the latent at X_test, one latent"""
    X_test = np.asarray(X_test, dtype=np.float64).reshape(-1)
    K = model.kernel.K(model.t, model.t)
    Ks = model.kernel.K(model.t, X_test)
    Kss = model.kernel.K(X_test, X_test)
    Ky = K + np.diag(model.site_cov.reshape(-1))
    A = solve(Ky, Ks)
    mean = A.T @ model.site_mean.reshape(-1, 1)
    cov = Kss - A.T @ Ks
    return mean.reshape(-1), np.diag(cov).copy()


def likelihood_predict(lik, mean_f, var_f, num_quad_pts=20):
    """Likelihood.predict for scalar latents: Gaussian closed form (likelihoods.py:802-803), else predict_cubature
    (cubature.py:438-465)"""
    m, v = np.asarray(mean_f, dtype=np.float64).reshape(-1), np.asarray(var_f, dtype=np.float64).reshape(-1)
    if isinstance(lik, sites.Gaussian):
        return m.copy(), v + lik.variance
    x, w = sites.gauss_hermite(1, num_quad_pts)
    f = np.sqrt(v)[:, None] * x[0][None, :] + m[:, None]
    ce, cc = lik.conditional_moments(f)
    ey = np.sum(w[None] * ce, axis=-1)
    ey2 = np.sum(w[None] * (cc + ce ** 2), axis=-1)
    return ey, ey2 - ey ** 2
