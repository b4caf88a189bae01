"""Prediction at test inputs.

CPU: the oracle's Markov prediction against the dense-GP predictive on the reference's regression grid
(tests/test_gp_vs_markovgp_reg.py:44-47 compare posteriors the same way), and the identity "predicting at a training
input returns the smoother marginal".  GPU: bn_temporal_conditional / bn_likelihood_predict through the host mirror
against the oracle, relative 1e-9."""
import numpy as np
import pytest

from _data import classification_data, regression_data, rel_err
from oracle import model, predict as opred, sites, ssm

TOL = 1e-9


def query_points(x, n, seed=3):
    rng = np.random.default_rng(seed)
    inside = rng.uniform(x.min(), x.max(), n)
    return np.concatenate([[x.min() - 7.0, x.min() - 0.01], inside, x[[0, 5, -1]], [x.max() + 0.02, x.max() + 11.0]])


@pytest.mark.parametrize('var_f', [0.5, 1.5])
@pytest.mark.parametrize('len_f', [0.75, 2.5])
@pytest.mark.parametrize('var_y', [0.1, 0.5])
@pytest.mark.parametrize('kern', ['Matern12', 'Matern32', 'Matern52', 'Matern72'])
def test_oracle_markov_predict_vs_dense_gp(var_f, len_f, var_y, kern):
    x, y = regression_data(40)
    k = getattr(ssm, kern)(var_f, len_f)
    m = model.MarkovGP(k, sites.Gaussian(var_y), x, y, method='vi')
    g = model.DenseGP(k, sites.Gaussian(var_y), x, y, method='vi')
    m.inference()
    g.inference()
    xs = query_points(m.t, 25)
    pm, pv = opred.markov_predict(m, xs)
    dm, dv = opred.dense_predict(g, xs)
    assert np.abs(pm[:, 0] - dm).max() < 1e-6 and np.abs(pv[:, 0, 0] - dv).max() < 1e-6


def test_oracle_predict_at_training_inputs_is_the_smoother():
    x, y = classification_data(30)
    m = model.MarkovGP(ssm.Matern52(1.2, 0.8), sites.Bernoulli(), x, y, method='vi')
    m.inference(lr=0.6)
    pm, pv = opred.markov_predict(m, m.t)
    assert rel_err(pm[..., None], m.post_mean) < 1e-6 and rel_err(pv, m.post_cov) < 1e-6


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy()


KERNELS = {
    'm12': lambda K: K.Matern12(0.8, 1.7), 'm32': lambda K: K.Matern32(1.1, 0.6), 'm52': lambda K: K.Matern52(1.3, 0.9),
    'm72': lambda K: K.Matern72(0.7, 1.4),
    'ind32': lambda K: K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.5, 2.0)]),
    'ind12': lambda K: K.Independent([K.Matern12(1.0, 1.0), K.Matern12(0.5, 2.0)]),
}


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(KERNELS))
@pytest.mark.parametrize('N', [1, 2, 57])
def test_gpu_temporal_conditional_vs_oracle(bn, name, N):
    from oracle import kalman
    kg, ko = KERNELS[name](bn.kernels), KERNELS[name](ssm)
    D = 2 if name.startswith('ind') else 1
    rng = np.random.default_rng(N)
    x = np.cumsum(0.1 + 0.4 * rng.uniform(size=N))
    dt = np.concatenate([[0.0], np.diff(x)])
    y = rng.standard_normal((N, D, 1))
    R = np.stack([np.diag(0.3 + rng.uniform(size=D)) for _ in range(N)])
    _, (fm, fP) = kalman.kalman_filter(dt, ko, y, R)
    sm, sP, G = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), ko, fm, fP, return_full=True)
    xs = query_points(x, 40, seed=N) if N > 5 else np.array([x[0] - 3.0, x[0], x[0] + 0.05, x[-1], x[-1] + 2.5])
    tm0, tc0 = opred.temporal_conditional(x, xs, sm, sP, G, ko)
    tm1, tc1 = bn.ops.temporal_conditional(x, xs, sm, sP, G, kg)
    assert rel_err(np_(tm1), tm0) < TOL and rel_err(np_(tc1), tc0) < TOL
    H = ko.measurement_model()
    hm, hc = bn.ops.temporal_conditional(np.concatenate([[-1e10], x, [1e10]]), xs, sm, sP, G, kg, return_full=False)
    assert rel_err(np_(hm), H @ tm0) < TOL and rel_err(np_(hc), H @ tc0 @ H.T) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('lik', ['gaussian', 'probit', 'logit'])
def test_gpu_model_predict_vs_oracle(bn, lik):
    x, y = regression_data(80) if lik == 'gaussian' else classification_data(80)
    lg = {'gaussian': lambda L: L.Gaussian(0.2), 'probit': lambda L: L.Bernoulli('probit'), 'logit': lambda L: L.Bernoulli('logit')}[lik]
    mo = model.MarkovGP(ssm.Matern52(1.5, 0.75), lg(sites), x, y, method='vi')
    mg = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.5, 0.75), likelihood=lg(bn.likelihoods), X=x, Y=y)
    mo.inference(lr=0.8)
    mg.inference(lr=0.8)
    xs = query_points(mo.t, 60)
    pm0, pv0 = opred.markov_predict(mo, xs)
    pm1, pv1 = mg.predict(xs)
    assert rel_err(np_(pm1), pm0[:, 0]) < TOL and rel_err(np_(pv1), pv0[:, 0, 0]) < TOL
    ey0, vy0 = opred.likelihood_predict(mo.likelihood, pm0[:, 0], pv0[:, 0, 0])
    ey1, vy1 = mg.predict_y(xs)
    assert rel_err(np_(ey1), ey0) < TOL and rel_err(np_(vy1), vy0) < TOL
    ys = np.where(np.arange(xs.shape[0]) % 2 == 0, 1.0, 0.0) if lik != 'gaussian' else np.sin(xs)
    ld0, _, _ = sites.moment_match(mo.likelihood, ys, pm0[:, 0], pv0[:, 0, 0], 1.0)
    nlpd1 = float(mg.negative_log_predictive_density(xs, ys))
    assert abs(nlpd1 + np.nanmean(ld0)) <= TOL * abs(np.nanmean(ld0))
