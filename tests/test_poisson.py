"""Poisson likelihood with the exp link (bayesnewton/likelihoods.py:891-1008): count data,
e.g. the log-Gaussian Cox process demos.  CPU: the oracle's closed-form variational expectation against 1-D cubature of
the log-likelihood (what the base class would do) and the Newton statistics against finite differences.  GPU: every
inference scheme end to end against the oracle, prediction included."""
import numpy as np
import pytest

from _data import rel_err
from oracle import model, predict as opred, sites, ssm

TOL = 1e-9


def count_data(N, seed=0, binsize=1.0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0, 30, N))
    rate = binsize * np.exp(0.8 * np.sin(0.5 * x) + 0.3)
    return x, rng.poisson(rate).astype(np.float64)


def test_oracle_closed_form_matches_cubature():
    lik = sites.Poisson(0.7)
    rng = np.random.default_rng(1)
    y = rng.poisson(2.0, 50).astype(np.float64)
    m, v = 0.5 * rng.standard_normal(50), 0.05 + 0.3 * rng.uniform(size=50)
    E, dE, d2E = sites.variational_expectation(lik, y, m, v)
    x, w = sites.gauss_hermite(1, 40)
    f = np.sqrt(v)[:, None] * x[0][None] + m[:, None]
    Ec = np.sum(w[None] * lik.log_lik(y[:, None], f), -1)
    assert np.abs(E - Ec).max() < 1e-10
    h = 1e-5
    Ep, _, _ = sites.variational_expectation(lik, y, m + h, v)
    Em, _, _ = sites.variational_expectation(lik, y, m - h, v)
    assert np.abs((Ep - Em) / (2 * h) - dE).max() < 1e-7 and np.abs((Ep - 2 * E + Em) / h ** 2 - d2E).max() < 1e-4
    ll, d1, d2 = lik.log_lik_derivs(y, m)
    assert np.abs((lik.log_lik(y, m + h) - lik.log_lik(y, m - h)) / (2 * h) - d1).max() < 1e-7


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('method', ['vi', 'ep', 'newton', 'pl'])
@pytest.mark.parametrize('binsize', [1.0, 0.25])
def test_gpu_poisson_iteration_vs_oracle(bn, method, binsize):
    x, y = count_data(300, seed=3, binsize=binsize)
    y[[7, 120]] = np.nan
    cls = {'vi': bn.models.MarkovVariationalGP, 'ep': bn.models.MarkovExpectationPropagationGP,
           'newton': bn.models.MarkovNewtonGP, 'pl': bn.models.MarkovPosteriorLinearisationGP}[method]
    kw = dict(power=0.5) if method == 'ep' else {}
    mg = cls(kernel=bn.kernels.Matern52(1.0, 2.0), likelihood=bn.likelihoods.Poisson(binsize=binsize), X=x, Y=y, **kw)
    mo = model.MarkovGP(ssm.Matern52(1.0, 2.0), sites.Poisson(binsize), x, y, method=method, power=0.5)
    for lr in (0.5, 0.5, 0.3):
        mo.inference(lr=lr)
        mg.inference(lr=lr)
    assert rel_err(np_(mg.posterior_mean), mo.post_mean) < TOL
    assert rel_err(np_(mg.posterior_variance), mo.post_cov) < TOL
    E0, E1 = mo.energy(), float(mg.energy())
    assert abs(E1 - E0) <= TOL * abs(E0), (E0, E1)
    if method == 'vi':
        xs = np.linspace(-2, 33, 41)
        pm0, pv0 = opred.markov_predict(mo, xs)
        ey0, vy0 = opred.likelihood_predict(mo.likelihood, pm0[:, 0], pv0[:, 0, 0])
        ey1, vy1 = mg.predict_y(xs)
        assert rel_err(np_(ey1), ey0) < TOL and rel_err(np_(vy1), vy0) < TOL
