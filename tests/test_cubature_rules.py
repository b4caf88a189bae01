"""Cubature rules the site kernels accept as tables (bayesnewton/cubature.py:12-162): exactness on Gaussian moments
(CPU) and use through the fused site kernel against the oracle evaluated with the same rule (GPU)."""
import numpy as np
import pytest

from bayesnewton_b200 import cubature as cub


def moment(x, w, powers):
    x = np.atleast_2d(x)
    return float(np.sum(w * np.prod(x ** np.array(powers)[:, None], axis=0)))


@pytest.mark.parametrize('rule,order', [(cub.UnscentedThirdOrder(), 3), (cub.UnscentedFifthOrder(), 5), (cub.Unscented(), 5),
                                        (cub.GaussHermite(num_cub_points=7), 13)])
@pytest.mark.parametrize('dim', [1, 2])
def test_rules_integrate_gaussian_moments_exactly(rule, order, dim):
    x, w = rule(dim)
    assert abs(np.sum(w) - 1) < 1e-14
    dfact = lambda n: 1.0 if n <= 0 else n * dfact(n - 2)  # E[z^p] = (p-1)!! for even p
    import itertools
    for powers in itertools.product(range(order + 1), repeat=dim):
        if sum(powers) > order:
            continue
        exact = 1.0
        for p in powers:
            exact *= 0.0 if p % 2 else dfact(p - 1)
        assert abs(moment(x, w, powers) - exact) < 1e-11 * max(1.0, exact), (powers, moment(x, w, powers), exact)


def test_reference_shapes():
    x, w = cub.Unscented()(2)
    assert x.shape == (2, 9) and w.shape == (9,)
    x, w = cub.UnscentedThirdOrder()(1)
    assert np.allclose(x, [0.0, 1.0, -1.0]) and np.allclose(w, [0.0, 0.5, 0.5])
    assert cub.GaussHermite(dim=1)(1)[0].shape == (1, 20)


@pytest.mark.gpu
@pytest.mark.parametrize('rule', ['u3', 'u5', 'gh9'])
def test_gpu_site_statistics_with_a_custom_rule(rule):
    """the kernels take any 1-D table: variational expectation of the probit likelihood under each rule vs numpy"""
    import bayesnewton_b200 as bn
    from oracle import sites
    r = {'u3': cub.UnscentedThirdOrder(), 'u5': cub.Unscented(), 'gh9': cub.GaussHermite(num_cub_points=9)}[rule]
    rng = np.random.default_rng(0)
    N = 257
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    m, v = rng.standard_normal(N), 0.2 + rng.uniform(size=N)
    x, w = r(1)
    x = np.atleast_2d(x)
    lik = sites.Bernoulli()
    f = np.sqrt(v)[:, None] * x[0][None, :] + m[:, None]
    wl = w[None, :] * lik.log_lik(y[:, None], f)
    E0 = wl.sum(-1)
    dE0 = ((f - m[:, None]) / v[:, None] * wl).sum(-1)
    d2E0 = 2 * ((0.5 * (f - m[:, None]) ** 2 / v[:, None] ** 2 - 0.5 / v[:, None]) * wl).sum(-1)
    E1, dE1, d2E1 = bn.likelihoods.Bernoulli().variational_expectation(y, m.reshape(N, 1, 1), v.reshape(N, 1, 1), cubature=r)
    g = lambda t: t.detach().cpu().numpy().reshape(-1)
    assert np.abs(g(E1) - E0).max() < 1e-9 * np.abs(E0).max()
    assert np.abs(g(dE1) - dE0).max() < 1e-9 * np.abs(dE0).max()
    assert np.abs(g(d2E1) - d2E0).max() < 1e-9 * max(np.abs(d2E0).max(), 1.0)
