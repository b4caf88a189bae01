"""The plain-C restatement (oracle/c/markov_c.c, the timed CPU baseline) agrees with the NumPy oracle."""
import numpy as np
import pytest

from _data import bench_inputs, classification_data, rel_err
from oracle import cport, model, sites, ssm


@pytest.mark.parametrize('family,K', [(1, ssm.Matern12), (2, ssm.Matern32), (3, ssm.Matern52), (4, ssm.Matern72)])
@pytest.mark.parametrize('lik', [1, 2])
def test_cport_iteration_matches_numpy_oracle(family, K, lik):
    x, y = classification_data(120)
    if lik == 1:
        y = y + 0.1 * np.random.default_rng(0).standard_normal(y.shape[0])
    y[::23] = np.nan
    olik = sites.Gaussian(0.3) if lik == 1 else sites.Bernoulli()
    o = model.MarkovGP(K(1.5, 0.75), olik, x, y, method='vi')
    c = cport.ViModel(family, 1.5, 0.75, lik, 0.3, o.dt, o.Y[:, 0])
    for it in range(3):
        o.inference(lr=0.6)
        E0 = o.energy()
        E1 = c.iteration(lr=0.6)
        assert abs(E1 - E0) < 1e-10 * abs(E0)
        assert rel_err(c.post_mean, o.post_mean[:, 0, 0]) < 1e-10 and rel_err(c.post_var, o.post_cov[:, 0, 0]) < 1e-10
        assert rel_err(c.nat1, o.site_nat1[:, 0, 0]) < 1e-10 and rel_err(c.nat2, o.site_nat2[:, 0, 0]) < 1e-10


def test_cport_on_bench_workload():
    t, dt, y = bench_inputs(5000)
    o = model.MarkovGP(ssm.Matern52(1.0, 1.0), sites.Bernoulli(), t, y, method='vi')
    c = cport.ViModel(3, 1.0, 1.0, 2, 0.0, dt, y)
    o.inference(lr=1.0)
    assert abs(c.iteration(1.0) - o.energy()) < 1e-10 * abs(o.energy())
