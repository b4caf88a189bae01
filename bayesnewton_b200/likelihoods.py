"""Likelihood objects: the per-step statistics the inference schemes vmap over time, evaluated for
all N steps by one fused CUDA kernel (libbn_b200: bn_likelihood_stats).

Mirrors bayesnewton/likelihoods.py for the likelihoods on the hot path -- Gaussian (:684-803),
Bernoulli / Probit / Logit (:806-888), HeteroscedasticNoise (:1244-1281) -- with the same method
names; the difference is that every method here is already batched over the leading time axis
(the reference calls them through jax.vmap, inference.py:114,179,255,348).
"""
import math

import torch

from . import _lib
from ._util import as_dev, ptr, stream_ptr, workspace
from .cubature import host_table


def softplus(x):
    return math.log(1.0 + math.exp(x))


def softplus_inv(x):
    return math.log(math.exp(x) - 1.0)


class Likelihood:
    lik_id = None
    num_latents = 1
    multi_latent = False

    @property
    def lik_param(self):
        return 0.0

    @property
    def lik_param2(self):
        return 0.0

    def site_args(self, method, y, mean, cov, cubature=None, power=1.0):
        """bn_site_args with the inputs filled in; returns (args, keepalive)"""
        mean, cov = as_dev(mean), as_dev(cov)
        D = self.num_latents
        N = mean.numel() // D
        a = _lib.SiteArgs()
        a.method, a.likelihood, a.lik_param, a.N, a.D = method, self.lik_id, self.lik_param, N, D
        a.lik_param2 = float(self.lik_param2)
        keep = [mean, cov]
        closed = ((self.lik_id == _lib.BN_LIK_GAUSSIAN and method in (_lib.BN_METHOD_VI, _lib.BN_METHOD_EP))
                  or (self.lik_id == _lib.BN_LIK_POISSON_EXP and method == _lib.BN_METHOD_VI))
        if method != _lib.BN_METHOD_NEWTON and not closed:
            cx, cw, Q = host_table(cubature, D)  # host arrays: the rule is a kernel parameter
            a.Q, a.cub_x, a.cub_w = Q, cx.ctypes.data, cw.ctypes.data
            keep += [cx, cw]
        if y is not None:
            y = as_dev(y).reshape(-1)
            if y.numel() != N:
                raise ValueError('y must hold one observation per time step (N = %d)' % N)
            keep.append(y)
        a.y, a.post_mean, a.post_cov = ptr(y), mean.data_ptr(), cov.data_ptr()
        a.power, a.lr, a.ensure_psd = float(power), 1.0, 0
        return a, keep

    def _stats(self, method, y, mean, cov, cubature=None, power=1.0):
        a, keep = self.site_args(method, y, mean, cov, cubature, power)
        N, D = a.N, a.D
        dev = keep[0].device
        val = torch.empty((N,), dtype=torch.float64, device=dev)
        d1 = torch.empty((N, D, 1), dtype=torch.float64, device=dev)
        d2 = torch.empty((N, D, D), dtype=torch.float64, device=dev)
        ws, nb = workspace(N, D, D)
        _lib.check(_lib.lib().bn_likelihood_stats(a, ptr(val), ptr(d1), ptr(d2), ptr(ws), nb, stream_ptr()))
        return val, d1, d2

    # ---- the reference's method names, batched over N -------------------------------------------------
    def variational_expectation(self, y, m, v, cubature=None):
        """E_q[log p(y|f)], dE/dm, d2E/dm2  (likelihoods.py:363-383 / :613-664)"""
        return self._stats(_lib.BN_METHOD_VI, y, m, v, cubature)

    def moment_match(self, y, m, v, power=1.0, cubature=None):
        """log Z, dlZ/dm, d2lZ/dm2 at the cavity (m, v)  (likelihoods.py:401-412 / :597-611)"""
        return self._stats(_lib.BN_METHOD_EP, y, m, v, cubature, power)

    def log_likelihood_gradients(self, y, f):
        """log p(y|f), Jacobian, Hessian at f  (likelihoods.py:336-355 / :669-675)"""
        f = as_dev(f)
        D = self.num_latents
        dummy_cov = torch.ones((f.numel() // D, D, D), dtype=torch.float64, device=f.device)
        return self._stats(_lib.BN_METHOD_NEWTON, y, f, dummy_cov)

    def log_density(self, y, m, v, cubature=None):
        """log E_q[p(y|f)] = the EP log-partition at power 1 (likelihoods.py:385-388 / Gaussian :784-795)"""
        return self._stats(_lib.BN_METHOD_EP, y, m, v, cubature, 1.0)[0]

    def predict(self, mean_f, var_f, cubature=None):
        """(E[y], Var[y]) at every test point from the latent marginals (likelihoods.py:493-506, 802-803;
        cubature.py:438-465); single-latent likelihoods"""
        if self.multi_latent:
            raise NotImplementedError('predict for multi-latent likelihoods')
        m, v = as_dev(mean_f).reshape(-1), as_dev(var_f).reshape(-1)
        N = m.shape[0]
        my, vy = torch.empty_like(m), torch.empty_like(m)
        cx = cw = None
        Q = 0
        if self.lik_id != _lib.BN_LIK_GAUSSIAN:
            hx, hw, Q = host_table(cubature, 1)
            cx, cw = as_dev(hx.reshape(-1)), as_dev(hw)
        _lib.check(_lib.lib().bn_likelihood_predict2(self.lik_id, float(self.lik_param), float(self.lik_param2), N, ptr(m),
                                                     ptr(v), Q, ptr(cx), ptr(cw), ptr(my), ptr(vy), stream_ptr()))
        return my, vy

    def statistical_linear_regression(self, m, v, cubature=None):
        """mu = E_q[E[y|f]], omega, dmu/dm  (cubature.py:374-435); single-latent likelihoods"""
        if self.multi_latent:
            raise NotImplementedError('statistical linear regression is implemented for single-latent likelihoods')
        mu, dmu, omega = self._stats(_lib.BN_METHOD_PL, None, m, v, cubature)
        return mu.reshape(-1, 1, 1), omega, dmu


class Gaussian(Likelihood):
    """p(y|f) = N(y | f, variance)"""
    lik_id = _lib.BN_LIK_GAUSSIAN

    def __init__(self, variance=0.1, fix_variance=False):
        self.transformed_variance = softplus_inv(float(variance))
        self.fix_variance = fix_variance

    @property
    def variance(self):
        return softplus(self.transformed_variance)

    @property
    def lik_param(self):
        return self.variance

    def chain_to_transformed(self, grad):
        """gradient w.r.t. the variance -> w.r.t. the stored softplus-transformed variable (likelihoods.py:700-710)"""
        return grad / (1.0 + math.exp(-self.transformed_variance))


class Bernoulli(Likelihood):
    """p(y|f) = P^y (1-P)^(1-y), P = link(f); probit carries the reference's 1e-3 jitter (likelihoods.py:828-829)"""

    def __init__(self, link='probit'):
        if link == 'probit':
            self.lik_id = _lib.BN_LIK_BERNOULLI_PROBIT
        elif link == 'logit':
            self.lik_id = _lib.BN_LIK_BERNOULLI_LOGIT
        else:
            raise NotImplementedError('link function not implemented')
        self.link = link


class Poisson(Likelihood):
    """p(y|f) = Poisson(y | binsize * exp(f)) for count data (likelihoods.py:891-1008); the variational expectation is
    the reference's closed form, EP / PL go through the 1-D cubature rule"""
    lik_id = _lib.BN_LIK_POISSON_EXP

    def __init__(self, binsize=1, link='exp'):
        if link != 'exp':
            raise NotImplementedError('the logistic link of the Poisson likelihood is not compiled into libbn_b200')
        self.binsize, self.link = float(binsize), link

    @property
    def lik_param(self):
        return self.binsize


class StudentsT(Likelihood):
    """p(y|f) = St(y | f, scale, df) (likelihoods.py:1011-1044); every scheme goes through the 1-D cubature rule
    (Newton: the closed derivatives of the log-density)"""
    lik_id = _lib.BN_LIK_STUDENTS_T

    def __init__(self, scale=1.0, df=3.0, fix_scale=False):
        self.transformed_scale = softplus_inv(float(scale))
        self.df, self.fix_scale = float(df), fix_scale

    @property
    def scale(self):
        return softplus(self.transformed_scale)

    lik_param = scale

    @property
    def lik_param2(self):
        return self.df


class Gamma(Likelihood):
    """p(y|f) = Gamma(y | shape, scale = exp(f)) (likelihoods.py:1100-1138)"""
    lik_id = _lib.BN_LIK_GAMMA_EXP

    def __init__(self, link='exp', shape=1.0):
        if link != 'exp':
            raise NotImplementedError('the logistic link of the Gamma likelihood is not compiled into libbn_b200')
        self.link = link
        self.transformed_shape = softplus_inv(float(shape))

    @property
    def shape(self):
        return softplus(self.transformed_shape)

    lik_param = shape


class NegativeBinomial(Likelihood):
    """p(y|f) = NB(y | mean = scale exp(f), dispersion alpha) (likelihoods.py:1141-1189)"""
    lik_id = _lib.BN_LIK_NEGBIN_EXP

    def __init__(self, alpha=1.0, link='exp', scale=1.0):
        if link != 'exp':
            raise NotImplementedError('the logistic link of the negative-binomial likelihood is not compiled into libbn_b200')
        self.link = link
        self.transformed_alpha = softplus_inv(float(alpha))
        self.scale = float(scale)

    @property
    def alpha(self):
        return softplus(self.transformed_alpha)

    lik_param = alpha

    @property
    def lik_param2(self):
        return self.scale


class Beta(Likelihood):
    """p(y|f) = Beta(y | scale m, scale (1 - m)), m = probit link with the reference's 1e-3 jitter
    (likelihoods.py:1047-1097)"""
    lik_id = _lib.BN_LIK_BETA_PROBIT

    def __init__(self, link='probit', scale=1.0, fix_scale=False):
        if link != 'probit':
            raise NotImplementedError('the logit link of the Beta likelihood is not compiled into libbn_b200')
        self.link = link
        self.transformed_scale = softplus_inv(float(scale))
        self.fix_scale = fix_scale

    @property
    def scale(self):
        return softplus(self.transformed_scale)

    lik_param = scale


class Probit(Bernoulli):
    def __init__(self):
        super().__init__('probit')


class Logit(Bernoulli):
    def __init__(self):
        super().__init__('logit')


Erf, Logistic = Probit, Logit


class MultiLatentLikelihood(Likelihood):
    multi_latent = True


class HeteroscedasticNoise(MultiLatentLikelihood):
    """p(y|f1,f2) = N(y | f1, link(f2)^2)  (likelihoods.py:1244-1281)"""
    num_latents = 2

    def __init__(self, link='softplus'):
        if link == 'softplus':
            self.lik_id = _lib.BN_LIK_HETEROSCEDASTIC_SOFTPLUS
        elif link == 'exp':
            self.lik_id = _lib.BN_LIK_HETEROSCEDASTIC_EXP
        else:
            raise NotImplementedError('link function not implemented')
        self.link = link
