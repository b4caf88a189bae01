"""Cubature site statistics and the Newton-step site update (oracle; test infrastructure).

Restates, batched over the time axis N (the reference ``vmap``s the same
per-step functions):
  * Gauss-Hermite tables                     cubature.py:56-84
  * VI  expected log-lik + score derivatives cubature.py:198-246, likelihoods.py:363-383
  * EP  moment matching                      cubature.py:310-371, likelihoods.py:401-412
  * PL  statistical linear regression        cubature.py:374-435
  * multi-latent (pathwise/autodiff) forms   likelihoods.py:561-664
  * Gaussian closed forms                    likelihoods.py:727-782, utils.py:431-466
  * Bernoulli / HeteroscedasticNoise         likelihoods.py:806-860, 1244-1281
  * newton_update, ensure_psd                inference.py:21-39, utils.py:89-96
  * the VI / EP / Newton / PL update bodies  inference.py:99-128,170-195,238-284,339-371
  * compute_cavity, reparametrise            utils.py:534-541, basemodels.py:85-100
Estimator fidelity: single-latent likelihoods use the Gaussian
score identity for derivatives; multi-latent ones use what JAX autodiff of the
cubature sum gives (pathwise), written out analytically here.
"""
import itertools
import math
import numpy as np
from numpy.polynomial.hermite import hermgauss
from .linalg import T, chol, cho_solve, solve, inv

try:  # scipy is in the image; math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

LOG2PI = math.log(2 * math.pi)
INV2PI = (2 * math.pi) ** -1


def gauss_hermite(dim=1, num_quad_pts=20):
    """sigma points x[dim, Q**dim] and weights w[Q**dim]; first coordinate slowest (cubature.py:56-84)"""
    gh_x, gh_w = hermgauss(num_quad_pts)
    x = np.array(list(itertools.product(*(gh_x,) * dim)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * dim))), 1)
    return np.sqrt(2) * x.T, w.T * np.pi ** (-0.5 * dim)


def softplus(x):
    return np.log(1. + np.exp(x))


def sigmoid(x):
    return np.exp(x) / (np.exp(x) + 1.)


# ----------------------------------------------------------------------------- likelihoods

class Gaussian:
    multi_latent = False
    name = 'gaussian'

    def __init__(self, variance=0.1):
        self.variance = float(variance)

    def log_lik(self, y, f):
        return -0.5 * np.log(2 * np.pi * self.variance) - 0.5 * (y - f) ** 2 / self.variance

    def log_lik_derivs(self, y, f):
        return self.log_lik(y, f), (y - f) / self.variance, np.full(np.broadcast(y, f).shape, -1. / self.variance)

    def conditional_moments(self, f):
        return f, np.full_like(f, self.variance)

    def dconditional_mean(self, f):
        return np.ones_like(f)


class Bernoulli:
    multi_latent = False
    name = 'bernoulli'

    def __init__(self, link='probit'):
        assert link in ('probit', 'logit')
        self.link = link

    def p(self, f):
        if self.link == 'logit':
            return 1 / (1 + np.exp(-f))
        jitter = 1e-3
        return 0.5 * (1.0 + _erf(f / np.sqrt(2.0))) * (1 - 2 * jitter) + jitter

    def dp(self, f):
        if self.link == 'logit':
            return np.exp(f) / (1 + np.exp(f)) ** 2
        return (1 - 2e-3) * np.exp(-0.5 * f * f) / np.sqrt(2 * np.pi)

    def d2p(self, f):
        if self.link == 'logit':
            s = 1 / (1 + np.exp(-f))
            return s * (1 - s) * (1 - 2 * s)
        return -f * self.dp(f)

    def log_lik(self, y, f):
        p = self.p(f)
        return np.log(np.where(np.equal(y, 1), p, 1 - p))

    def log_lik_derivs(self, y, f):
        p, dp, d2p = self.p(f), self.dp(f), self.d2p(f)
        one = np.equal(y, 1)
        q = np.where(one, p, 1 - p)
        s = np.where(one, 1.0, -1.0)
        return np.log(q), s * dp / q, s * d2p / q - (dp / q) ** 2

    def conditional_moments(self, f):
        p = self.p(f)
        return p, p - p ** 2

    def dconditional_mean(self, f):
        return self.dp(f)


class Poisson:
    """p(y|f) = Poisson(y | binsize exp(f)); likelihoods.py:891-1008 (exp link)"""
    multi_latent = False
    name = 'poisson'

    def __init__(self, binsize=1.0):
        self.binsize = float(binsize)

    def log_lik(self, y, f):
        from scipy.special import gammaln
        mu = np.exp(f) * self.binsize
        return y * np.log(mu) - mu - gammaln(y + 1.0)

    def log_lik_derivs(self, y, f):
        mu = np.exp(f) * self.binsize
        return self.log_lik(y, f), y - mu, -mu * np.ones(np.broadcast(y, f).shape)

    def conditional_moments(self, f):
        mu = np.exp(f) * self.binsize
        return mu, mu

    def dconditional_mean(self, f):
        return np.exp(f) * self.binsize


class StudentsT:
    """p(y|f) = St(y | f, scale, df); likelihoods.py:1011-1044"""
    multi_latent = False
    name = 'studentst'

    def __init__(self, scale=1.0, df=3.0):
        self.scale, self.df = float(scale), float(df)

    def log_lik(self, y, f):
        from scipy.special import gammaln
        c = gammaln((self.df + 1.0) * 0.5) - gammaln(self.df * 0.5) - 0.5 * (np.log(np.square(self.scale)) + np.log(self.df) + np.log(np.pi))
        return c - 0.5 * (self.df + 1.0) * np.log(1.0 + (1.0 / self.df) * np.square((y - f) / self.scale))

    def log_lik_derivs(self, y, f):
        r, a = y - f, self.df * self.scale ** 2
        return self.log_lik(y, f), (self.df + 1.0) * r / (a + r * r), (self.df + 1.0) * (r * r - a) / (a + r * r) ** 2

    def conditional_moments(self, f):
        return f, (self.scale ** 2) * (self.df / (self.df - 2.0)) * np.ones_like(f)

    def dconditional_mean(self, f):
        return np.ones_like(f)


class Gamma:
    """p(y|f) = Gamma(y | shape, scale = exp(f)); likelihoods.py:1100-1138"""
    multi_latent = False
    name = 'gamma'

    def __init__(self, shape=1.0):
        self.shape = float(shape)

    def log_lik(self, y, f):
        from scipy.special import gammaln
        sc = np.exp(f)
        with np.errstate(invalid='ignore', divide='ignore'):
            return -self.shape * np.log(sc) - gammaln(self.shape) + (self.shape - 1.0) * np.log(y) - y / sc

    def log_lik_derivs(self, y, f):
        t = y * np.exp(-f)
        return self.log_lik(y, f), -self.shape + t, -t

    def conditional_moments(self, f):
        sc = np.exp(f)
        return self.shape * sc, self.shape * sc ** 2

    def dconditional_mean(self, f):
        return self.shape * np.exp(f)


class NegativeBinomial:
    """p(y|f) = NB(y | mean scale exp(f), alpha); likelihoods.py:1141-1189"""
    multi_latent = False
    name = 'negbin'

    def __init__(self, alpha=1.0, scale=1.0):
        self.alpha, self.scale = float(alpha), float(scale)

    def log_lik(self, y, f):
        from scipy.special import gammaln
        m, k = np.exp(f) * self.scale, 1.0 / self.alpha
        return gammaln(k + y) - gammaln(y + 1) - gammaln(k) + y * np.log(m / (m + k)) - k * np.log(1 + m * self.alpha)

    def log_lik_derivs(self, y, f):
        m, k = np.exp(f) * self.scale, 1.0 / self.alpha
        return self.log_lik(y, f), k * (y - m) / (m + k), -k * m * (k + y) / (m + k) ** 2

    def conditional_moments(self, f):
        E = np.exp(f) * self.scale
        return E, E + E ** 2 * self.alpha

    def dconditional_mean(self, f):
        return np.exp(f) * self.scale


class Beta:
    """p(y|f) = Beta(y | scale m, scale (1 - m)), m = jittered probit link; likelihoods.py:1047-1097"""
    multi_latent = False
    name = 'beta'

    def __init__(self, scale=1.0):
        self.scale = float(scale)

    def link(self, f):
        return 0.5 * (1.0 + _erf(f / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3

    def dlink(self, f):
        return (1 - 2e-3) * np.exp(-0.5 * f * f) / np.sqrt(2 * np.pi)

    def log_lik(self, y, f):
        from scipy.special import gammaln
        al = self.link(f) * self.scale
        be = self.scale - al
        y = np.clip(y, 1e-6, 1. - 1e-6)
        return (al - 1.0) * np.log(y) + (be - 1.0) * np.log(1.0 - y) + gammaln(al + be) - gammaln(al) - gammaln(be)

    def log_lik_derivs(self, y, f):
        from scipy.special import digamma, polygamma
        al = self.link(f) * self.scale
        be = self.scale - al
        yc = np.clip(y, 1e-6, 1. - 1e-6)
        g = np.log(yc) - np.log(1.0 - yc) - digamma(al) + digamma(be)
        gp = -polygamma(1, al) - polygamma(1, be)
        da = self.scale * self.dlink(f)
        return self.log_lik(y, f), g * da, gp * da * da + g * self.scale * (-f * self.dlink(f))

    def conditional_moments(self, f):
        p = self.link(f)
        return p, (p - p ** 2) / (self.scale + 1.0)

    def dconditional_mean(self, f):
        return self.dlink(f)


class HeteroscedasticNoise:
    """p(y|f1,f2) = N(y | f1, link(f2)^2); likelihoods.py:1244-1281"""
    multi_latent = True
    name = 'heteroscedastic'

    def __init__(self, link='softplus'):
        assert link in ('softplus', 'exp')
        self.link = link

    def _g(self, f2):
        if self.link == 'exp':
            e = np.exp(f2)
            return e, e, e
        g1 = sigmoid(f2)
        return softplus(f2), g1, g1 * (1. - g1)

    def log_lik(self, y, f):
        """f[..., 2]"""
        g, _, _ = self._g(f[..., 1])
        var = g ** 2
        return -0.5 * np.log(2 * np.pi * var) - 0.5 * (y - f[..., 0]) ** 2 / var

    def log_lik_derivs(self, y, f):
        """value, gradient [...,2], Hessian [...,2,2] of log p(y|f) w.r.t. f (what jacrev gives, likelihoods.py:322-330)"""
        g, g1, g2 = self._g(f[..., 1])
        r = y - f[..., 0]
        ll = -0.5 * np.log(2 * np.pi * g ** 2) - 0.5 * r ** 2 / g ** 2
        d1 = r / g ** 2
        d2 = -g1 / g + r ** 2 * g1 / g ** 3
        h11 = -1. / g ** 2 + 0. * r
        h12 = -2. * r * g1 / g ** 3
        h22 = -(g2 * g - g1 ** 2) / g ** 2 + r ** 2 * (g2 / g ** 3 - 3. * g1 ** 2 / g ** 4)
        grad = np.stack([d1, d2], axis=-1)
        hess = np.stack([np.stack([h11, h12], -1), np.stack([h12, h22], -1)], -2)
        return ll, grad, hess


# ----------------------------------------------------------------------------- single-latent cubature

def variational_expectation(lik, y, m, v, num_quad_pts=20):
    """E_q[log p(y|f)], d/dm, d2/dm2 for scalar latents; y,m,v: [N].
    Gaussian closed form: likelihoods.py:727-753; else cubature.py:198-246 wrapped by likelihoods.py:363-383
    (NaN y -> value 0, derivatives NaN)."""
    y, m, v = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (y, m, v))
    mask = np.isnan(y)
    y = np.where(mask, m, y)
    if isinstance(lik, Gaussian):
        E = -0.5 * np.log(2 * np.pi) - 0.5 * np.log(lik.variance) - 0.5 * ((y - m) ** 2 + v) / lik.variance
        dE = (y - m) / lik.variance
        d2E = np.full_like(m, -1. / lik.variance)
    elif isinstance(lik, Poisson):  # closed form, likelihoods.py:979-1008
        from scipy.special import gammaln
        emc = lik.binsize * np.exp(m + v / 2)
        E = y * np.log(lik.binsize) + y * m - emc - gammaln(y + 1.0)
        dE = y - emc
        d2E = -emc
    else:
        x, w = gauss_hermite(1, num_quad_pts)
        sd = np.sqrt(v)  # 1x1 Cholesky
        f = sd[:, None] * x[0][None, :] + m[:, None]
        wl = w[None, :] * lik.log_lik(y[:, None], f)
        invv = 1. / v
        E = np.sum(wl, axis=-1)
        dE = np.sum(invv[:, None] * (f - m[:, None]) * wl, axis=-1)
        dEdv = np.sum((0.5 * (invv[:, None] ** 2 * (f - m[:, None]) ** 2) - 0.5 * invv[:, None]) * wl, axis=-1)
        d2E = 2 * dEdv
    E = np.where(mask, 0., E)
    dE = np.where(mask, np.nan, dE)
    d2E = np.where(mask, np.nan, d2E)
    return E, dE, d2E


def pep_constant(var, power, mask=None):
    """utils.py:431-445 for scalar sites; var [N]"""
    logdiag = np.log(np.abs(np.sqrt(var)))
    dim = np.ones_like(var)
    if mask is not None:
        logdiag = np.where(mask, 0., logdiag)
        dim = dim - mask.astype(var.dtype)
    return 0.5 * dim * ((1 - power) * LOG2PI - np.log(power)) + 0.5 * (1 - power) * 2 * logdiag


def moment_match(lik, y, cm, cv, power=1.0, num_quad_pts=20):
    """log Z and derivatives w.r.t. the cavity mean, scalar latents; [N] each.
    Gaussian: likelihoods.py:755-782 (+utils.py:448-466); else cubature.py:310-371 via likelihoods.py:401-412."""
    y, cm, cv = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (y, cm, cv))
    mask = np.isnan(y)
    y = np.where(mask, cm, y)
    if isinstance(lik, Gaussian):
        var = lik.variance / power + cv
        lZ = -0.5 * ((y - cm) ** 2 / var + LOG2PI + np.log(var))  # via chol: log_det = 2 log|sqrt(var)|
        dlZ = (y - cm) / var
        d2lZ = -1. / var
        lZ = lZ + pep_constant(np.full_like(cv, lik.variance), power)
        return lZ, dlZ, d2lZ
    x, w = gauss_hermite(1, num_quad_pts)
    sd = np.sqrt(cv)
    f = sd[:, None] * x[0][None, :] + cm[:, None]
    wp = w[None, :] * np.exp(power * lik.log_lik(y[:, None], f))
    Z = np.sum(wp, axis=-1)
    lZ = np.log(np.maximum(Z, 1e-8))
    Zinv = 1.0 / np.maximum(Z, 1e-8)
    icv = 1. / cv  # inv() of a 1x1 through Cholesky
    dZ = np.sum(icv[:, None] * (f - cm[:, None]) * wp, axis=-1)
    dlZ = Zinv * dZ
    d2Z = np.sum((icv[:, None] * (f - cm[:, None]) * (f - cm[:, None]) * icv[:, None] - icv[:, None]) * wp, axis=-1)
    d2lZ = -dlZ * dlZ + Zinv * d2Z
    return lZ, dlZ, d2lZ


def statistical_linear_regression(lik, m, v, num_quad_pts=20):
    """mu, omega, dmu/dm for scalar latents (cubature.py:374-435); [N] each"""
    m, v = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (m, v))
    x, w = gauss_hermite(1, num_quad_pts)
    f = np.sqrt(v)[:, None] * x[0][None, :] + m[:, None]
    Ey, Cy = lik.conditional_moments(f)
    mu = np.sum(w * Ey, axis=-1)
    S = np.sum(w * ((Ey - mu[:, None]) ** 2 + Cy), axis=-1)
    C = np.sum(w * (f - m[:, None]) * (Ey - mu[:, None]), axis=-1)
    omega = S - C * (C / v)
    dmu = np.sum(w * lik.dconditional_mean(f), axis=-1)
    return mu, omega, dmu


def log_likelihood_gradients(lik, y, f):
    """log p(y|f) and its first two derivatives at f (likelihoods.py:322-355); NaN y -> 0 / NaN / NaN"""
    y, f = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (y, f))
    mask = np.isnan(y)
    y = np.where(mask, f, y)
    ll, J, Hh = lik.log_lik_derivs(y, f)
    return np.where(mask, 0., ll), np.where(mask, np.nan, J), np.where(mask, np.nan, Hh)


# ----------------------------------------------------------------------------- multi-latent cubature (2 latents)

def _sigma_points_2d(m, V, num_quad_pts):
    """f[N,Q,2] = chol(sym(V)) x + m"""
    x, w = gauss_hermite(2, num_quad_pts)
    V = (V + T(V)) / 2
    L = chol(V)
    f = np.einsum('nij,jq->nqi', L, x) + m[:, None, :]
    return f, w


def variational_expectation_ml(lik, y, m, V, num_quad_pts=20):
    """likelihoods.py:613-664: E = sum w l_i, dE = sum w grad l_i, d2E = sum w hess l_i.  y[N], m[N,2], V[N,2,2]"""
    f, w = _sigma_points_2d(m, V, num_quad_pts)
    ll, g, h = lik.log_lik_derivs(y[:, None], f)
    return np.sum(w * ll, -1), np.einsum('q,nqi->ni', w, g), np.einsum('q,nqij->nij', w, h)


def moment_match_ml(lik, y, cm, cV, power=1.0, num_quad_pts=20):
    """likelihoods.py:561-611 (no 1e-8 clamp on Z in this branch)"""
    f, w = _sigma_points_2d(cm, cV, num_quad_pts)
    ll, g, h = lik.log_lik_derivs(y[:, None], f)
    wp = w * np.exp(power * ll)
    Z = np.sum(wp, -1)
    dlZ = np.einsum('nq,nqi->ni', wp, power * g) / Z[:, None]
    second = power * h + power ** 2 * g[..., :, None] * g[..., None, :]
    d2lZ = np.einsum('nq,nqij->nij', wp, second) / Z[:, None, None] - dlZ[:, :, None] * dlZ[:, None, :]
    return np.log(Z), dlZ, d2lZ


def log_likelihood_gradients_ml(lik, y, f):
    """likelihoods.py:669-675, 1278-1281: value, gradient, Hessian at the posterior mean.  f[N,2]"""
    return lik.log_lik_derivs(y, f)


# ----------------------------------------------------------------------------- Newton step on the sites

def ensure_diagonal_positive_precision(K):
    """utils.py:89-96; K[N,D,D]"""
    D = K.shape[-1]
    Kd = np.zeros_like(K)  # vmap(np.diag) PLACES the diagonal: off-diagonals are exact zeros even for NaN entries
    for i in range(D):
        Kd[:, i, i] = K[:, i, i]
    with np.errstate(invalid='ignore'):
        return np.where(Kd < 0, 1e-2, Kd)


def newton_update(mean, jacobian, hessian):
    """inference.py:21-39; mean/jacobian [N,D,1], hessian [N,D,D]"""
    hessian = np.where(np.isnan(hessian), -1e-6, hessian)
    jacobian = np.where(np.isnan(jacobian), hessian @ mean, jacobian)
    return jacobian - hessian @ mean, -hessian


def reparametrise(p1, p2):
    """basemodels.py:85-90: (nat1, nat2) <-> (mean, cov) by a batched Cholesky inverse"""
    L = chol(p2)
    eye = np.broadcast_to(np.eye(p2.shape[-1]), p2.shape)
    return cho_solve(L, p1), cho_solve(L, eye)


def compute_cavity(post_mean, post_cov, nat1, nat2, power, jitter=1e-8):
    """utils.py:534-541, batched"""
    D = post_cov.shape[-1]
    post_nat2 = inv(post_cov + jitter * np.eye(D))
    cav_cov = inv(post_nat2 - power * nat2)
    cav_mean = cav_cov @ (post_nat2 @ post_mean - power * nat1)
    return cav_mean, cav_cov


def site_statistics(method, lik, Y, post_mean, post_cov, nat1=None, nat2=None, power=1.0, ensure_psd=True,
                    num_quad_pts=20, mask_pseudo_y=None):
    """the ``update_variational_params`` bodies -> (mean, jacobian, hessian), shapes [N,D,1],[N,D,1],[N,D,D].
    method in {'vi','ep','newton','pl'} (inference.py:105-128, 170-195, 238-284, 339-371), temporal models only
    (conditional_posterior_to_data is the identity, basemodels.py:759-764)."""
    N, D = post_mean.shape[0], post_mean.shape[1]
    Yv = np.asarray(Y, dtype=np.float64).reshape(N, -1)
    if method == 'ep':
        mean, cov = compute_cavity(post_mean, post_cov, nat1, nat2, power)
    else:
        mean, cov = post_mean, post_cov

    if lik.multi_latent:
        y = Yv[:, 0]
        if method == 'vi':
            _, jac, hess = variational_expectation_ml(lik, y, mean[:, :, 0], cov, num_quad_pts)
        elif method == 'ep':
            _, jac, hess = moment_match_ml(lik, y, mean[:, :, 0], cov, power, num_quad_pts)
        elif method == 'newton':
            _, jac, hess = log_likelihood_gradients_ml(lik, y, mean[:, :, 0])
        else:
            raise NotImplementedError('PL for multi-latent likelihoods is outside the oracle')
        jac = jac[:, :, None]
    else:
        assert D == 1
        y, m, v = Yv[:, 0], mean[:, 0, 0], cov[:, 0, 0]
        if method == 'vi':
            _, j, h = variational_expectation(lik, y, m, v, num_quad_pts)
        elif method == 'ep':
            _, j, h = moment_match(lik, y, m, v, power, num_quad_pts)
        elif method == 'newton':
            _, j, h = log_likelihood_gradients(lik, y, m)
        elif method == 'pl':
            mu, omega, dmu = statistical_linear_regression(lik, m, v, num_quad_pts)
            res = y - mu
            msk = np.isnan(res)
            res = np.where(msk, 0., res)
            omega = np.where(msk, 1e6, omega)
            dmu_omega = dmu / omega
            j, h = dmu_omega * res, -dmu_omega * dmu
        else:
            raise ValueError(method)
        jac, hess = j[:, None, None], h[:, None, None]

    if method == 'ep':
        cav_prec = inv(cov)
        scale = cav_prec @ inv(hess + cav_prec) / power
        jac = scale @ jac
        hess = scale @ hess
        if mask_pseudo_y is not None:
            mk = np.asarray(mask_pseudo_y).reshape(N, D)[..., None]
            jac = np.where(mk, np.nan, jac)
            hm = np.where(mk | T(mk), 0., hess)
            hess = np.where((np.eye(D, dtype=bool)[None] & mk), np.nan, hm)
    if ensure_psd and method != 'pl':
        hess = -ensure_diagonal_positive_precision(-hess)
    return mean, jac, hess


def damped_site_update(nat1_old, nat2_old, mean, jac, hess, lr):
    """inference.py:76-86 + basemodels.py:97-100 -> new (nat1, nat2, mean, cov) and the two mean-abs diffs"""
    nat1_n, nat2_n = newton_update(mean, jac, hess)
    diff1 = np.mean(np.abs(nat1_n - nat1_old))
    diff2 = np.mean(np.abs(nat2_n - nat2_old))
    nat1 = (1 - lr) * nat1_old + lr * nat1_n
    nat2 = (1 - lr) * nat2_old + lr * nat2_n
    mean_s, cov_s = reparametrise(nat1, nat2)
    return nat1, nat2, mean_s, cov_s, diff1, diff2


def gaussian_expected_log_lik(Y, q_mu, q_covar, noise, mask=None):
    """utils.py:510-531, batched: [N,D,1],[N,D,1],[N,D,D],[N,D,D]"""
    from .kalman import mvn_logpdf
    D = noise.shape[-1]
    if mask is not None:
        maskv = np.asarray(mask).reshape(noise.shape[:-2] + (D, 1))
        eye = np.eye(D, dtype=bool)
        q_mu = np.where(maskv, Y, q_mu)
        noise = np.where(maskv | T(maskv), 0., noise)
        noise = np.where(eye & maskv, INV2PI, noise)
        q_covar = np.where(maskv | T(maskv), 0., q_covar)
        q_covar = np.where(eye & maskv, 1e-20, q_covar)
    ml = mvn_logpdf(Y, q_mu, noise)
    trace_term = -0.5 * np.trace(solve(noise, q_covar), axis1=-2, axis2=-1)
    return ml + trace_term
