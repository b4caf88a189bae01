"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) by the host mirror of the
reference interface, against the CPU oracle on the same seeded inputs.

Tolerance (north_star): relative 1e-9 on posterior means, covariances and energy in fp64, measured
normwise (max|a-b| / max|b|) -- posterior means cross zero.  Observed: ~1e-15.
"""
import os

import numpy as np
import pytest

import _emu
from _data import bench_inputs, classification_data, filter_problem, rel_err
from oracle import kalman, model, sites, ssm

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy()


def make_kernels(bn):
    K, O = bn.kernels, ssm
    return {
        'm12': (K.Matern12(0.8, 1.7), O.Matern12(0.8, 1.7), 1),
        'm32': (K.Matern32(1.1, 0.6), O.Matern32(1.1, 0.6), 1),
        'm52': (K.Matern52(1.3, 0.9), O.Matern52(1.3, 0.9), 1),
        'm72': (K.Matern72(0.7, 1.4), O.Matern72(0.7, 1.4), 1),
        'ind32': (K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.5, 2.0)]),
                  O.Independent([O.Matern32(1.0, 1.0), O.Matern32(0.5, 2.0)]), 2),
        'ind52': (K.Independent([K.Matern52(1.3, 0.9), K.Matern52(0.7, 2.1)]),
                  O.Independent([O.Matern52(1.3, 0.9), O.Matern52(0.7, 2.1)]), 2),
        'mixed': (K.Independent([K.Matern32(1.0, 1.0), K.Matern12(0.5, 2.0)]),
                  O.Independent([O.Matern32(1.0, 1.0), O.Matern12(0.5, 2.0)]), 2),
    }


@pytest.mark.parametrize('name', ['m12', 'm32', 'm52', 'm72', 'ind32', 'ind52', 'mixed'])
@pytest.mark.parametrize('parallel', [False, True])
@pytest.mark.parametrize('N', [1, 7, 203, 3001])
def test_filter_smoother_vs_oracle(bn, name, parallel, N):
    kg, ko, D = make_kernels(bn)[name]
    dt, y, R, mask = filter_problem(N, D=D, seed=N)
    for rp in (False, True):
        e0, (m0, P0) = kalman.kalman_filter(dt, ko, y, R, mask, return_predict=rp)
        e1, (m1, P1) = bn.ops.kalman_filter(dt, kg, y, R, mask, parallel=parallel, return_predict=rp)
        assert abs(float(e1) - e0) <= TOL * abs(e0)
        assert rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    _, (fm, fP) = kalman.kalman_filter(dt, ko, y, R, mask)
    dts = np.concatenate([dt[1:], [0.0]])
    for rf in (False, True):
        s0 = kalman.rauch_tung_striebel_smoother(dts, ko, fm, fP, return_full=rf)
        s1 = bn.ops.rauch_tung_striebel_smoother(dts, kg, fm, fP, return_full=rf, parallel=parallel)
        for a, b in zip(s1, s0):
            assert rel_err(np_(a), b) < TOL


@pytest.mark.parametrize('name', ['m12', 'm32', 'm52', 'm72', 'ind32', 'ind52'])
@pytest.mark.parametrize('N', [1, 7, 8, 203, 3001, 70001])
def test_fused_update_posterior_vs_oracle(bn, name, N):
    """bn_update_posterior (filter + smoother in one call, smoothing elements derived from the filter's chunk
    elements) against the oracle's sequential filter -> smoother"""
    kg, ko, D = make_kernels(bn)[name]
    dt, y, R, mask = filter_problem(N, D=D, seed=N + 1)
    e0, (fm, fP) = kalman.kalman_filter(dt, ko, y, R, mask)
    sm, sP, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), ko, fm, fP)
    e1, m1, P1 = bn.ops.update_posterior(dt, kg, y, R, mask, want_ell=True)
    assert abs(float(e1) - e0) <= TOL * abs(e0)
    assert rel_err(np_(m1), sm) < TOL and rel_err(np_(P1), sP) < TOL
    e2, m2, P2 = bn.ops.update_posterior(dt, kg, y, R, None, want_ell=False)
    assert e2 is None and rel_err(np_(m2), sm) < TOL and rel_err(np_(P2), sP) < TOL


@pytest.mark.parametrize('name', ['m52', 'ind32'])
@pytest.mark.parametrize('N,shards', [(9, 3), (1000, 2), (70001, 5), (70001, 8)])
def test_fused_update_in_time_shards(bn, name, N, shards):
    """the three-phase sharded form of the fused update (what each rank of a multi-GPU run executes),
    all shards in one process, against the single-call result and the oracle"""
    from bayesnewton_b200 import distributed
    kg, ko, D = make_kernels(bn)[name]
    dt, y, R, mask = filter_problem(N, D=D, seed=N + 2)
    e0, (fm, fP) = kalman.kalman_filter(dt, ko, y, R, mask)
    sm, sP, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), ko, fm, fP)
    res = distributed.update_posterior_in_shards(kg, dt, y, R, mask, shards)
    assert abs(float(res['ell']) - e0) <= TOL * abs(e0)
    assert rel_err(np_(res['post_mean']), sm) < TOL and rel_err(np_(res['post_cov']), sP) < TOL


def test_no_mask_and_skipped_outputs(bn):
    kg, ko, _ = make_kernels(bn)['m52']
    dt, y, R, _ = filter_problem(500, seed=3)
    e0, (m0, P0) = kalman.kalman_filter(dt, ko, y, R)
    e1, (m1, P1) = bn.ops.kalman_filter(dt, kg, y, R, parallel=True)
    assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0) < TOL
    e2, (m2, P2) = bn.ops.kalman_filter(dt, kg, y, R, parallel=True, want_ell=False)
    assert e2 is None and rel_err(np_(m2), m0) < TOL and rel_err(np_(P2), P0) < TOL
    e3, (m3, P3) = bn.ops.kalman_filter(dt, kg, y, R, parallel=True, want_states=False)
    assert m3 is None and abs(float(e3) - e0) <= TOL * abs(e0)


def test_array_level_entry_points(bn):
    ko = ssm.Matern52(1.3, 0.9)
    dt, y, R, mask = filter_problem(400, seed=5)
    As, Qs = ssm.discretise(ko, dt)
    H, Pinf = ko.measurement_model(), ko.stationary_covariance()
    m0 = 0.1 * np.ones((3, 1))
    for f_or, f_gpu in ((kalman.sequential_kf, bn.ops._sequential_kf), (kalman.parallel_kf, bn.ops._parallel_kf)):
        e0, m0_, P0_ = f_or(As, Qs, H, y, R, m0, Pinf, mask)
        e1, m1, P1 = f_gpu(As, Qs, H, y, R, m0, Pinf, mask)
        assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0_) < TOL and rel_err(np_(P1), P0_) < TOL
    _, fm, fP = kalman.sequential_kf(As, Qs, H, y, R, m0, Pinf, mask)
    dts = np.concatenate([dt[1:], [0.0]])
    As2, Qs2 = ssm.discretise(ko, dts)
    for f_or, f_gpu in ((kalman.sequential_rts, bn.ops._sequential_rts), (kalman.parallel_rts, bn.ops._parallel_rts)):
        for rf in (False, True):
            s0 = f_or(fm, fP, As2, Qs2, H, rf)
            s1 = f_gpu(fm, fP, As2, Qs2, H, rf)
            assert all(rel_err(np_(a), b) < TOL for a, b in zip(s1, s0))


@pytest.mark.parametrize('name', ['m12', 'm32', 'm52', 'm72', 'ind32'])
def test_discretise(bn, name):
    kg, ko, _ = make_kernels(bn)[name]
    dt = np.concatenate([[0.0], np.random.default_rng(0).uniform(0.01, 2.0, 300)])
    A0, Q0 = ssm.discretise(ko, dt)
    A1, Q1 = bn.kernels.discretise(kg, dt)
    assert rel_err(np_(A1), A0) < 1e-13 and rel_err(np_(Q1), Q0) < 1e-11
    assert np.allclose(kg.stationary_covariance(), ko.stationary_covariance(), rtol=1e-15)
    assert np.array_equal(kg.measurement_model(), ko.measurement_model())
    assert rel_err(np_(kg.state_transition(0.37)), ko.state_transition(0.37)) < 1e-13


def test_golden_fixtures(bn):
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'markov_small.npz'))
    kers = make_kernels(bn)
    for name in ('m12', 'm32', 'm52', 'm72', 'ind32'):
        kg = kers[name][0]
        for par in (False, True):
            e, (fm, fP) = bn.ops.kalman_filter(g[name + '_dt'], kg, g[name + '_y'], g[name + '_R'], g[name + '_mask'],
                                               parallel=par)
            assert abs(float(e) - g[name + '_ell']) <= TOL * abs(g[name + '_ell'])
            assert rel_err(np_(fm), g[name + '_fm']) < TOL and rel_err(np_(fP), g[name + '_fP']) < TOL
            dts = np.concatenate([g[name + '_dt'][1:], [0.0]])
            sm, sP, G = bn.ops.rauch_tung_striebel_smoother(dts, kg, g[name + '_fm'], g[name + '_fP'], parallel=par)
            assert rel_err(np_(sm), g[name + '_sm']) < TOL and rel_err(np_(sP), g[name + '_sP']) < TOL
            assert rel_err(np_(G), g[name + '_gains']) < TOL
    models = {'vi': bn.models.MarkovVariationalGP, 'ep': bn.models.MarkovExpectationPropagationGP,
              'newton': bn.models.MarkovNewtonGP, 'pl': bn.models.MarkovPosteriorLinearisationGP}
    for method, cls in models.items():
        kw = dict(power=0.5) if method == 'ep' else {}
        m = cls(kernel=bn.kernels.Matern52(1.5, 0.75), likelihood=bn.likelihoods.Bernoulli(), X=g['cls_x'],
                Y=g['cls_y'], **kw)
        m.inference(lr=0.7)
        m.inference(lr=0.7)
        assert rel_err(np_(m.posterior_mean), g['cls_%s_post_mean' % method]) < TOL
        assert rel_err(np_(m.posterior_variance), g['cls_%s_post_cov' % method]) < TOL
        assert rel_err(np_(m.pseudo_likelihood.nat1), g['cls_%s_site_nat1' % method]) < TOL
        assert rel_err(np_(m.pseudo_likelihood.nat2), g['cls_%s_site_nat2' % method]) < TOL
        E = float(m.energy())
        assert abs(E - g['cls_%s_energy' % method]) <= TOL * abs(g['cls_%s_energy' % method])


LIKS = {'probit': (lambda bn: bn.likelihoods.Bernoulli('probit'), lambda: sites.Bernoulli('probit')),
        'logit': (lambda bn: bn.likelihoods.Logit(), lambda: sites.Bernoulli('logit')),
        'gaussian': (lambda bn: bn.likelihoods.Gaussian(0.3), lambda: sites.Gaussian(0.3))}


@pytest.mark.parametrize('likname', sorted(LIKS))
def test_likelihood_statistics(bn, likname):
    """the vmapped per-step signatures of SURVEY 8b: (Y, mean, cov, [power]) -> (val, d1, d2)"""
    rng = np.random.default_rng(4)
    N = 1000
    lg, lo = LIKS[likname][0](bn), LIKS[likname][1]()
    y = (rng.uniform(size=N) < 0.5).astype(float) if likname != 'gaussian' else rng.standard_normal(N)
    y[::13] = np.nan
    m = rng.standard_normal(N); v = 0.2 + rng.uniform(size=N)
    E0 = sites.variational_expectation(lo, y, m, v)
    E1 = lg.variational_expectation(y, m[:, None, None], v[:, None, None])
    assert all(rel_err(np_(a).reshape(-1), b) < TOL for a, b in zip(E1, E0))
    Z0 = sites.moment_match(lo, y, m, v, 0.5)
    Z1 = lg.moment_match(y, m[:, None, None], v[:, None, None], power=0.5)
    assert all(rel_err(np_(a).reshape(-1), b) < TOL for a, b in zip(Z1, Z0))
    G0 = sites.log_likelihood_gradients(lo, y, m)
    G1 = lg.log_likelihood_gradients(y, m[:, None, None])
    assert all(rel_err(np_(a).reshape(-1), b) < TOL for a, b in zip(G1, G0))
    mu0, om0, dmu0 = sites.statistical_linear_regression(lo, m, v)
    mu1, om1, dmu1 = lg.statistical_linear_regression(m[:, None, None], v[:, None, None])
    assert rel_err(np_(mu1).reshape(-1), mu0) < TOL and rel_err(np_(om1).reshape(-1), om0) < TOL
    assert rel_err(np_(dmu1).reshape(-1), dmu0) < TOL


def test_heteroscedastic_statistics(bn):
    rng = np.random.default_rng(5)
    N = 500
    lg, lo = bn.likelihoods.HeteroscedasticNoise(), sites.HeteroscedasticNoise()
    y = rng.standard_normal(N)
    m = 0.5 * rng.standard_normal((N, 2))
    A = 0.3 * rng.standard_normal((N, 2, 2))
    V = A @ np.swapaxes(A, 1, 2) + 0.2 * np.eye(2)
    for f_or, got in ((sites.variational_expectation_ml(lo, y, m, V), lg.variational_expectation(y, m[..., None], V)),
                      (sites.moment_match_ml(lo, y, m, V, 0.5), lg.moment_match(y, m[..., None], V, power=0.5)),
                      (sites.log_likelihood_gradients_ml(lo, y, m), lg.log_likelihood_gradients(y, m[..., None]))):
        assert rel_err(np_(got[0]), f_or[0]) < TOL
        assert rel_err(np_(got[1])[..., 0], f_or[1]) < TOL and rel_err(np_(got[2]), f_or[2]) < TOL


@pytest.mark.parametrize('method', ['vi', 'ep', 'newton', 'pl'])
@pytest.mark.parametrize('parallel', [False, True])
def test_model_iteration_classification(bn, method, parallel):
    """the reference's own model-level test (tests/test_gp_vs_markovgp_class.py) with the oracle as comparator"""
    x, y = classification_data(200)
    y[::19] = np.nan
    cls = {'vi': bn.models.MarkovVariationalGP, 'ep': bn.models.MarkovExpectationPropagationGP,
           'newton': bn.models.MarkovLaplaceGP, 'pl': bn.models.MarkovPosteriorLinearisationGP}[method]
    kw = dict(power=0.5) if method == 'ep' else {}
    g = cls(kernel=bn.kernels.Matern52(1.5, 0.75), likelihood=bn.likelihoods.Bernoulli(), X=x, Y=y,
            parallel=parallel, **kw)
    o = model.MarkovGP(ssm.Matern52(1.5, 0.75), sites.Bernoulli(), x, y, method=method, power=0.5)
    for it in range(3):
        (mean, jac, hess), (d1, d2) = g.inference(lr=0.6, return_state=True)
        (mean0, jac0, hess0), (d10, d20) = o.inference(lr=0.6)
        assert rel_err(np_(mean), mean0) < TOL and rel_err(np_(jac), jac0) < TOL and rel_err(np_(hess), hess0) < TOL
        assert abs(float(d1) - d10) < TOL * d10 and abs(float(d2) - d20) < TOL * d20
        assert rel_err(np_(g.posterior_mean), o.post_mean) < TOL
        assert rel_err(np_(g.posterior_variance), o.post_cov) < TOL
        assert abs(float(g.energy()) - o.energy()) <= TOL * abs(o.energy())


@pytest.mark.parametrize('method', ['vi', 'ep', 'newton'])
def test_model_iteration_heteroscedastic(bn, method):
    """config C3 in miniature: Independent[Matern32 x2] + HeteroscedasticNoise (demos/heteroscedastic.py:49-56)"""
    rng = np.random.default_rng(3)
    N = 150
    x = np.sort(15 * rng.random(N))
    y = np.sin(x) + 0.3 * (1 + np.cos(x)) * rng.standard_normal(N)
    K = bn.kernels
    cls = {'vi': bn.models.MarkovVariationalGP, 'ep': bn.models.MarkovExpectationPropagationGP,
           'newton': bn.models.MarkovNewtonGP}[method]
    kw = dict(power=0.5) if method == 'ep' else {}
    g = cls(kernel=K.Independent([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)]),
            likelihood=bn.likelihoods.HeteroscedasticNoise(), X=x, Y=y, parallel=True, **kw)
    o = model.MarkovGP(ssm.Independent([ssm.Matern32(1.0, 1.0), ssm.Matern32(1.0, 1.0)]),
                       sites.HeteroscedasticNoise(), x, y, method=method, power=0.5)
    for it in range(2):
        g.inference(lr=0.3)
        o.inference(lr=0.3)
        assert rel_err(np_(g.posterior_mean), o.post_mean) < 1e-8
        assert rel_err(np_(g.posterior_variance), o.post_cov) < 1e-8
        assert abs(float(g.energy()) - o.energy()) <= 1e-8 * abs(o.energy())


def test_regression_config_c1(bn):
    """BASELINE config 1: demos/regression.py with MarkovVariationalGP, Matern52, Gaussian, N = 1000"""
    N = 1000
    x = np.linspace(-17, 147, N)
    rng = np.random.default_rng(12345)
    y = np.cos(0.04 * x + 0.33 * np.pi) * np.sin(0.2 * x) + np.sqrt(0.2) * rng.standard_normal(N)
    g = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 5.0), likelihood=bn.likelihoods.Gaussian(0.2),
                                      X=x, Y=y, parallel=False)
    o = model.MarkovGP(ssm.Matern52(1.0, 5.0), sites.Gaussian(0.2), x, y, method='vi')
    g.inference(lr=1.0)
    o.inference(lr=1.0)
    assert rel_err(np_(g.posterior_mean), o.post_mean) < TOL and rel_err(np_(g.posterior_variance), o.post_cov) < TOL
    E = float(g.energy())
    assert abs(E - o.energy()) <= TOL * abs(o.energy())
    assert abs(E - model.exact_marginal_likelihood(ssm.Matern52(1.0, 5.0), 0.2, x, y)) < 1e-7 * abs(E)


def test_multilevel_scan_against_fast_sequential(bn):
    """N = 600k gives > 65536 chunks, i.e. three scan levels; the checker is the host emulation of the
    sequential form (itself pinned to the oracle in tests/test_host_emulation.py)"""
    emu = _emu.load()
    N = 600_000
    t, dt, y = bench_inputs(N)
    rng = np.random.default_rng(9)
    R = 0.5 + rng.uniform(size=(N, 1, 1))
    yy = (y + 0.1 * rng.standard_normal(N)).reshape(N, 1, 1)
    sp = _emu.spec(3, [1.0], [1.0])
    e0, m0, P0 = _emu.kalman_filter(emu, sp, 0, dt, yy, R)
    e1, (m1, P1) = bn.ops.kalman_filter(dt, bn.kernels.Matern52(1.0, 1.0), yy, R, parallel=True)
    assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    dts = np.concatenate([dt[1:], [0.0]])
    s0 = _emu.rts_smoother(emu, sp, 0, dts, m0, P0)
    s1 = bn.ops.rauch_tung_striebel_smoother(dts, bn.kernels.Matern52(1.0, 1.0), m1, P1, parallel=True)
    assert all(rel_err(np_(a), b) < TOL for a, b in zip(s1, s0))


def test_full_size_properties_c2(bn):
    """BASELINE config 2 at full size (N = 1e7): size-independent checks.
    (1) the filter is causal: its first K outputs equal the oracle run on the first K steps;
    (2) Markov restart: the oracle started from the GPU state at step a-1 reproduces steps a..a+K;
    (3) the smoother is anti-causal: the last K outputs equal the oracle on the last K filtered states;
    (4) two time shards stitched by the carry API equal the single-call result."""
    import torch
    N, K = 10_000_000, 400
    t, dt, y = bench_inputs(N)
    kg, ko = bn.kernels.Matern52(1.0, 1.0), ssm.Matern52(1.0, 1.0)
    m = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
    m.inference(lr=1.0, return_state=False)  # non-trivial sites
    py, pv = np_(m.pseudo_likelihood.mean), np_(m.pseudo_likelihood.covariance)
    ell, (fm, fP) = bn.ops.kalman_filter(m.dt, kg, m.pseudo_likelihood.mean, m.pseudo_likelihood.covariance,
                                         parallel=True)
    fm_h, fP_h = np_(fm), np_(fP)
    assert np.isfinite(float(ell)) and np.isfinite(fm_h).all() and np.isfinite(fP_h).all()
    dt_h = np_(m.dt)
    _, (m0, P0) = kalman.kalman_filter(dt_h[:K], ko, py[:K], pv[:K])
    assert rel_err(fm_h[:K], m0) < TOL and rel_err(fP_h[:K], P0) < TOL
    for a in (N // 3, N - K):
        As, Qs = ssm.discretise(ko, dt_h[a:a + K])
        _, m1, P1 = kalman.sequential_kf(As, Qs, ko.measurement_model(), py[a:a + K], pv[a:a + K], fm_h[a - 1],
                                         fP_h[a - 1], np.zeros((K, 1, 1), bool))
        assert rel_err(fm_h[a:a + K], m1) < TOL and rel_err(fP_h[a:a + K], P1) < TOL
    sm, sP, _ = bn.ops.rauch_tung_striebel_smoother(m.dt_smoother, kg, fm, fP, parallel=True, return_full=True)
    sm_h, sP_h = np_(sm), np_(sP)
    dts_h = np_(m.dt_smoother)
    s0 = kalman.rauch_tung_striebel_smoother(dts_h[N - K:], ko, fm_h[N - K:], fP_h[N - K:], return_full=True)
    assert rel_err(sm_h[N - K:], s0[0]) < TOL and rel_err(sP_h[N - K:], s0[1]) < TOL
    a = N // 2  # window [a, a+K): append the GPU's smoothed state at a+K as a terminal "filtered" state with dt = 0
    fm_w = np.concatenate([fm_h[a:a + K], sm_h[a + K:a + K + 1]])
    fP_w = np.concatenate([fP_h[a:a + K], sP_h[a + K:a + K + 1]])
    dts_w = np.concatenate([dts_h[a:a + K], [0.0]])
    s1 = kalman.rauch_tung_striebel_smoother(dts_w, ko, fm_w, fP_w, return_full=True)
    assert rel_err(sm_h[a:a + K], s1[0][:K]) < TOL and rel_err(sP_h[a:a + K], s1[1][:K]) < TOL
    # posterior stored by the model = H-projection of the full smoothed state
    assert rel_err(np_(m.posterior_mean)[:, 0, 0], sm_h[:, 0, 0]) < TOL
    # (4) two shards through the carry API
    from bayesnewton_b200 import distributed
    res = distributed.filter_smoother_in_shards(kg, m.dt, m.pseudo_likelihood.mean, m.pseudo_likelihood.covariance,
                                                None, n_shards=2)
    assert abs(float(res['ell']) - float(ell)) <= TOL * abs(float(ell))
    assert rel_err(np_(res['post_mean']), np_(m.posterior_mean)) < TOL
    assert rel_err(np_(res['post_cov']), np_(m.posterior_variance)) < TOL
    res2 = distributed.update_posterior_in_shards(kg, m.dt, m.pseudo_likelihood.mean, m.pseudo_likelihood.covariance,
                                                  None, 3)
    assert abs(float(res2['ell']) - float(ell)) <= TOL * abs(float(ell))
    assert rel_err(np_(res2['post_mean']), np_(m.posterior_mean)) < TOL
    assert rel_err(np_(res2['post_cov']), np_(m.posterior_variance)) < TOL
    # the log-likelihood the fused update keeps for energy() is the stand-alone filter's
    assert abs(float(m.compute_log_lik()) - float(ell)) <= TOL * abs(float(ell))
    del res, res2, sm, sP, fm, fP
    torch.cuda.empty_cache()


@pytest.mark.parametrize('lik', ['gaussian', 'probit', 'logit'])
@pytest.mark.parametrize('method', ['vi', 'newton'])
def test_fused_energy_terms_match_the_separate_sums(bn, lik, method):
    """bn_energy_terms = bn_expected_density + bn_gaussian_expected_log_lik in one pass (missing data included)"""
    x, y = classification_data(700) if lik != 'gaussian' else __import__('_data').regression_data(700)
    y = y.copy()
    y[[3, 50, 51]] = np.nan
    L = bn.likelihoods
    lk = {'gaussian': L.Gaussian(0.2), 'probit': L.Bernoulli('probit'), 'logit': L.Bernoulli('logit')}[lik]
    cls = bn.models.MarkovVariationalGP if method == 'vi' else bn.models.MarkovNewtonGP
    m = cls(kernel=bn.kernels.Matern52(1.2, 0.8), likelihood=lk, X=x, Y=y)
    m.inference(lr=0.7)
    fused = m._energy_terms_fused()
    assert fused is not None
    e0, x0 = float(m.expected_density()), float(m.expected_density_pseudo())
    assert abs(float(fused[0]) - e0) <= 1e-12 * abs(e0) and abs(float(fused[1]) - x0) <= 1e-12 * abs(x0)
    assert abs(float(m.energy()) + (e0 - (x0 - float(m.compute_log_lik())))) <= 1e-12 * abs(float(m.energy()))
