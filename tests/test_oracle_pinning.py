"""Pins the oracle the way the reference pins its own Markov path: against the dense GP with the same
kernel (tests/test_gp_vs_markovgp_reg.py:44-70,121-138; tests/test_gp_vs_markovgp_class.py:37-60,109-124)
on the reference's parameter grids, and against the closed-form marginal likelihood
(tests/test_vs_exact_marg_lik.py:41-65) -- with FIXED seeds.  The reference asserts rtol 1e-4 /
2 decimals; the oracle meets 1e-9, which is what lets it serve as the 1e-9 yardstick for the kernels."""
import numpy as np
import pytest

from _data import classification_data, regression_data, rel_err
from oracle import kalman, model, sites, ssm


@pytest.mark.parametrize('var_f', [0.5, 1.5])
@pytest.mark.parametrize('len_f', [0.75, 2.5])
@pytest.mark.parametrize('var_y', [0.1, 0.5])
@pytest.mark.parametrize('N', [30, 60])
def test_markov_vs_dense_regression(var_f, len_f, var_y, N):
    x, y = regression_data(N)
    k = ssm.Matern52(var_f, len_f)
    a = model.MarkovGP(k, sites.Gaussian(var_y), x, y, method='vi')
    b = model.DenseGP(k, sites.Gaussian(var_y), x, y, method='vi')
    a.update_posterior(); b.update_posterior()
    assert rel_err(a.post_mean, b.post_mean) < 1e-9 and rel_err(a.post_cov, b.post_cov) < 1e-9
    assert abs(a.energy() - b.energy()) < 1e-8 * abs(b.energy())
    a.inference(lr=1.); b.inference(lr=1.)
    assert rel_err(a.post_mean, b.post_mean) < 1e-9 and rel_err(a.post_cov, b.post_cov) < 1e-9
    # one full-step VI iteration on a Gaussian likelihood is exact: energy = -log N(y | 0, K + s2 I)
    assert abs(a.energy() - model.exact_marginal_likelihood(k, var_y, a.t, a.Y)) < 1e-8 * abs(a.energy())


@pytest.mark.parametrize('var_f', [0.5, 1.5])
@pytest.mark.parametrize('len_f', [0.75, 2.5])
@pytest.mark.parametrize('N', [30, 60])
@pytest.mark.parametrize('method', ['vi', 'ep', 'newton', 'pl'])
def test_markov_vs_dense_classification(var_f, len_f, N, method):
    x, y = classification_data(N)
    k = ssm.Matern52(var_f, len_f)
    a = model.MarkovGP(k, sites.Bernoulli(), x, y, method=method)
    b = model.DenseGP(k, sites.Bernoulli(), x, y, method=method)
    for _ in range(2):
        a.inference(lr=0.7); b.inference(lr=0.7)
    assert rel_err(a.post_mean, b.post_mean) < 1e-9 and rel_err(a.post_cov, b.post_cov) < 1e-9
    assert abs(a.energy() - b.energy()) < 1e-9 * abs(b.energy())


@pytest.mark.parametrize('method', ['vi', 'ep', 'newton'])
def test_markov_vs_dense_heteroscedastic(method):
    rng = np.random.default_rng(3)
    N = 40
    x = np.sort(10 * rng.random(N))
    y = np.sin(x) + 0.3 * (1 + np.cos(x)) * rng.standard_normal(N)
    k = ssm.Independent([ssm.Matern32(1.0, 1.0), ssm.Matern32(1.0, 1.0)])
    lik = sites.HeteroscedasticNoise()
    a = model.MarkovGP(k, lik, x, y, method=method, power=0.5)
    b = model.DenseGP(k, lik, x, y, method=method, power=0.5)
    for _ in range(2):
        a.inference(lr=0.3); b.inference(lr=0.3)
    assert rel_err(a.post_mean, b.post_mean) < 1e-8 and rel_err(a.post_cov, b.post_cov) < 1e-8
    assert abs(a.energy() - b.energy()) < 1e-8 * abs(b.energy())


@pytest.mark.parametrize('K', [ssm.Matern12, ssm.Matern32, ssm.Matern52, ssm.Matern72])
def test_discretisation_matches_expm_and_lyapunov(K):
    """A = expm(F dt) (kernels.py:63-67 is the generic definition) and Pinf solves F P + P F^T + L Qc L^T = 0"""
    from scipy.linalg import expm
    k = K(1.3, 0.8)
    d = k.state_dim
    lam = {1: 1.0, 2: 3 ** 0.5, 3: 5 ** 0.5, 4: 7 ** 0.5}[d] / 0.8
    # companion form of (lam + s)^d
    from math import comb
    F = np.diag(np.ones(d - 1), 1)
    F[-1] = [-comb(d, i) * lam ** (d - i) for i in range(d)]
    for dt in [0.0, 0.01, 0.3, 2.0]:
        assert np.allclose(k.state_transition(dt), expm(F * dt), rtol=1e-11, atol=1e-13)
    Pinf = k.stationary_covariance()
    res = F @ Pinf + Pinf @ F.T
    res[-1, -1] = 0.0  # only the driven state has a diffusion term
    assert np.abs(res).max() < 1e-9 * np.abs(Pinf).max()


def test_scan_forms_agree_with_sequential():
    """the reference never tests parallel=True: pin it by equivalence, both evaluation orders"""
    rng = np.random.default_rng(0)
    N = 257
    k = ssm.Matern52(1.3, 0.9)
    dt = np.concatenate([[0], 0.1 + 0.2 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, 1, 1)); R = 0.3 + rng.uniform(size=(N, 1, 1))
    mask = rng.uniform(size=(N, 1, 1)) < 0.1
    e0, (m0, P0) = kalman.kalman_filter(dt, k, y, R, mask)
    dts = np.concatenate([dt[1:], [0]])
    s0 = kalman.rauch_tung_striebel_smoother(dts, k, m0, P0, return_full=True)
    for order, tol in (('fold', 1e-11), ('tree', 1e-7)):
        e1, (m1, P1) = kalman.kalman_filter(dt, k, y, R, mask, parallel=True, order=order)
        assert abs(e1 - e0) < tol * abs(e0) and rel_err(m1, m0) < tol and rel_err(P1, P0) < tol
        s1 = kalman.rauch_tung_striebel_smoother(dts, k, m0, P0, return_full=True, parallel=True, order=order)
        assert all(rel_err(a, b) < 1e-11 for a, b in zip(s1, s0))


def test_longdouble_rerun_bounds_fp64_error():
    rng = np.random.default_rng(1)
    N = 120
    dt = np.concatenate([[0], 0.1 + 0.2 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, 1, 1)); R = 0.3 + rng.uniform(size=(N, 1, 1))
    k64 = ssm.Matern52(1.0, 1.0)
    k80 = ssm.Matern52(1.0, 1.0, dtype=np.longdouble)
    e64, (m64, P64) = kalman.kalman_filter(dt, k64, y, R)
    e80, (m80, P80) = kalman.kalman_filter(dt.astype(np.longdouble), k80, y.astype(np.longdouble),
                                           R.astype(np.longdouble))
    assert m80.dtype == np.longdouble
    assert rel_err(m64, m80.astype(np.float64)) < 1e-12 and rel_err(P64, P80.astype(np.float64)) < 1e-12
    assert abs(e64 - float(e80)) < 1e-12 * abs(e64)


def test_golden_fixture_is_reproducible():
    """tests/golden/markov_small.npz was written by tests/golden/make_golden.py from this oracle"""
    import os
    path = os.path.join(os.path.dirname(__file__), 'golden', 'markov_small.npz')
    g = np.load(path)
    from golden.make_golden import build_cases
    fresh = build_cases()
    assert set(g.files) == set(fresh)
    for name in g.files:
        np.testing.assert_allclose(g[name], fresh[name], rtol=1e-12, atol=1e-14, equal_nan=True)
