"""GPU tests of the fused inference iteration on chunk-tiled state (csrc/iter_impl.cuh, C ABI bn_iter_*):
layout round trips, each fused pass against the oracle and against the library's own unfused entry points
(bn_update_posterior, bn_site_update, bn_energy_terms), the time-sharded phases, and the model-level iteration
(which takes the fused path by default) against the oracle model.  Tolerance 1e-9 normwise (north_star)."""
import os

import numpy as np
import pytest

from _data import bench_inputs, classification_data, rel_err
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


def kernels(bn):
    K, O = bn.kernels, ssm
    return {'m12': (K.Matern12(0.8, 1.7), O.Matern12(0.8, 1.7)), 'm32': (K.Matern32(1.1, 0.6), O.Matern32(1.1, 0.6)),
            'm52': (K.Matern52(1.3, 0.9), O.Matern52(1.3, 0.9)), 'm72': (K.Matern72(0.7, 1.4), O.Matern72(0.7, 1.4))}


@pytest.mark.parametrize('N', [1, 5, 32, 257, 3001, 70_001, 2_500_003])
def test_tiled_layout_round_trip(bn, N):
    import torch
    from bayesnewton_b200 import fused
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N)
    sh = fused.FusedShard(bn.kernels.Matern52(1.0, 1.0), np.abs(x), x)
    xt = sh.to_tiled(x, fill=-7.0)
    assert xt.numel() == sh.tlen
    back = np_(sh.from_tiled(xt))
    assert np.array_equal(back, x)
    # element (chunk c, step j) sits at ((c >> 5) * L + j) * 32 + (c & 31); everything else is the fill value
    L = sh.chunk_len
    k = np.arange(N)
    c, j = k // L, k % L
    idx = ((c >> 5) * L + j) * 32 + (c & 31)
    xt_h = np_(xt)
    assert np.array_equal(xt_h[idx], x)
    rest = np.ones(sh.tlen, dtype=bool)
    rest[idx] = False
    assert (xt_h[rest] == -7.0).all()


def _problem(N, lik, seed=0):
    t, dt, y = bench_inputs(N, seed)
    rng = np.random.default_rng(seed + 5)
    if lik == 'gaussian':
        y = np.sin(0.3 * t) + 0.4 * rng.standard_normal(N)
    elif lik == 'poisson':
        y = rng.poisson(np.exp(0.5 * np.sin(0.3 * t))).astype(np.float64)
    sy = 0.3 * rng.standard_normal(N)
    sR = 0.5 + rng.random(N)
    return t, dt, y, sy, sR


def _lik(bn, name):
    L = bn.likelihoods
    return {'probit': L.Bernoulli(link='probit'), 'logit': L.Bernoulli(link='logit'), 'gaussian': L.Gaussian(0.3),
            'poisson': L.Poisson()}[name]


@pytest.mark.parametrize('kname', ['m12', 'm32', 'm52', 'm72'])
@pytest.mark.parametrize('lik', ['probit', 'logit', 'gaussian', 'poisson'])
@pytest.mark.parametrize('method', ['vi', 'newton'])
@pytest.mark.parametrize('N', [1, 7, 203, 3001])
def test_fused_passes_vs_unfused_library_path(bn, kname, lik, method, N):
    """pass SITES = bn_update_posterior + bn_site_update; pass ENERGY = bn_update_posterior + bn_energy_terms"""
    import torch
    from bayesnewton_b200 import _lib, fused
    from bayesnewton_b200._util import ptr, stream_ptr, workspace
    kg, _ = kernels(bn)[kname]
    t, dt, y, sy, sR = _problem(N, lik, seed=N)
    if N > 20:
        y[::17] = np.nan  # missing observations: masked sites
    lk = _lik(bn, lik)
    meth = {'vi': _lib.BN_METHOD_VI, 'newton': _lib.BN_METHOD_NEWTON}[method]
    dev = torch.device('cuda')
    dt_d, y_d = torch.as_tensor(dt, device=dev), torch.as_tensor(y, device=dev)
    sy_d, sR_d = torch.as_tensor(sy, device=dev).reshape(N, 1, 1), torch.as_tensor(sR, device=dev).reshape(N, 1, 1)
    nan = torch.isnan(y_d)
    mask = nan.to(torch.uint8).contiguous() if bool(nan.any()) else None
    for lr in (1.0, 0.4):
        # unfused
        ell0, pm0, pc0 = bn.ops.update_posterior(dt_d, kg, sy_d, sR_d, mask=mask, want_ell=True)[:3]
        a, keep = lk.site_args(meth, y_d, pm0, pc0, None, 1.0)
        n2 = (1.0 / sR_d).clone()
        n1 = (sy_d * n2).clone()
        om, oc = torch.empty_like(sy_d), torch.empty_like(sR_d)
        diffs = torch.zeros(2, dtype=torch.float64, device=dev)
        a.nat1, a.nat2, a.site_mean, a.site_cov, a.diffs = n1.data_ptr(), n2.data_ptr(), om.data_ptr(), oc.data_ptr(), diffs.data_ptr()
        a.lr, a.ensure_psd = lr, 1
        ws, nb = workspace(N, 4, 1)
        _lib.check(_lib.lib().bn_site_update(a, ptr(ws), nb, stream_ptr()))
        # fused
        sh = fused.FusedShard(kg, dt_d, y_d, mask)
        sh.load_sites(sy_d, sR_d)
        ell1, d = sh.run(fused.SITES, lk, meth, None, lr, 1.0, True, want_ell=True)
        m1, c1 = sh.sites()
        assert abs(float(ell1) - float(ell0)) <= 1e-12 * abs(float(ell0)) + 1e-13
        assert rel_err(np_(m1), np_(om)) < 1e-11 and rel_err(np_(c1), np_(oc)) < 1e-11
        dd = np_(diffs) * N
        assert np.allclose(np_(d), dd, rtol=1e-10, atol=1e-300)
    # energy pass on the original sites
    a, keep = lk.site_args(meth, y_d, pm0, pc0, None, 1.0)
    a.site_mean, a.site_cov = sy_d.data_ptr(), sR_d.data_ptr()
    parts = torch.zeros(2, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().bn_energy_terms(a, ptr(mask), parts.data_ptr(), ptr(ws), nb, stream_ptr()))
    sh.load_sites(sy_d, sR_d)
    ell2, s = sh.run(fused.ENERGY, lk, meth, None, 1.0, 1.0, True)
    pm1, pc1 = sh.posterior()
    assert rel_err(np_(pm1), np_(pm0)) < 1e-12 and rel_err(np_(pc1), np_(pc0)) < 1e-12
    assert np.allclose(np_(s), np_(parts), rtol=1e-11, atol=1e-12)
    assert abs(float(ell2) - float(ell0)) <= 1e-12 * abs(float(ell0)) + 1e-13
    # plain pass
    sh.run(fused.PLAIN)
    pm2, pc2 = sh.posterior()
    assert rel_err(np_(pm2), np_(pm0)) < 1e-12 and rel_err(np_(pc2), np_(pc0)) < 1e-12
    # the same marginals written straight to [N, 1, 1] arrays in time order by the sweep itself
    pl, cl = torch.full((N, 1, 1), -3.0, dtype=torch.float64, device=dev), torch.full((N, 1, 1), -3.0, dtype=torch.float64, device=dev)
    sh.run(fused.PLAIN, post=(pl, cl))
    assert np.array_equal(np_(pl), np_(pm2)) and np.array_equal(np_(cl), np_(pc2))
    ell3, s3 = sh.run(fused.ENERGY, lk, meth, None, 1.0, 1.0, True, post=(pl, cl))
    assert np.array_equal(np_(pl), np_(pm1)) and np.array_equal(np_(cl), np_(pc1)) and np.array_equal(np_(s3), np_(s))


@pytest.mark.parametrize('kname', ['m32', 'm52'])
@pytest.mark.parametrize('N,world', [(203, 3), (3001, 2), (70_001, 8)])
def test_fused_shard_phases_stitch(bn, kname, N, world):
    """the three phases on `world` shards inside one process, carries handed over by hand, equal the single pass"""
    import torch
    from bayesnewton_b200 import _lib, fused
    kg, _ = kernels(bn)[kname]
    t, dt, y, sy, sR = _problem(N, 'probit', seed=3)
    lk = _lik(bn, 'probit')
    one = fused.FusedShard(kg, dt, y)
    one.load_sites(sy, sR)
    ell0, d0 = one.run(fused.SITES, lk, _lib.BN_METHOD_VI, None, 0.7)
    m0, c0 = one.sites()
    b = [N * r // world for r in range(world + 1)]
    shards = [fused.FusedShard(kg, dt[b[r]:b[r + 1]], y[b[r]:b[r + 1]], rank=r, world=world) for r in range(world)]
    for r, s in enumerate(shards):
        s.load_sites(sy[b[r]:b[r + 1]], sR[b[r]:b[r + 1]])
    kf = torch.stack([s.reduce(want_ell=True) for s in shards])
    filt = [s.filter(kf) for s in shards]
    rts = torch.stack([f[1] for f in filt])
    sums = [s.smooth(fused.SITES, rts, lk, _lib.BN_METHOD_VI, None, 0.7) for s in shards]
    ell1 = sum(float(f[0]) for f in filt)
    m1 = np.concatenate([np_(s.sites()[0]) for s in shards])
    c1 = np.concatenate([np_(s.sites()[1]) for s in shards])
    assert abs(ell1 - float(ell0)) <= 1e-11 * abs(float(ell0))
    assert rel_err(m1, np_(m0)) < 1e-10 and rel_err(c1, np_(c0)) < 1e-10
    assert np.allclose(sum(np_(x) for x in sums), np_(d0), rtol=1e-9)


@pytest.mark.parametrize('kname', ['m12', 'm32', 'm52', 'm72'])
@pytest.mark.parametrize('lik,method', [('probit', 'vi'), ('probit', 'newton'), ('logit', 'vi'), ('gaussian', 'vi'),
                                        ('poisson', 'vi'), ('poisson', 'newton')])
def test_model_iteration_fused_vs_oracle(bn, kname, lik, method):
    """model.inference() / model.energy() (the fused path by default) against the oracle model, 3 damped iterations,
    and against the same model with the fused path switched off"""
    kg, ko = kernels(bn)[kname]
    N = 400
    x, y = classification_data(N, seed=7)
    rng = np.random.default_rng(1)
    if lik == 'gaussian':
        y = np.sin(0.3 * x) + 0.4 * rng.standard_normal(N)
    elif lik == 'poisson':
        y = rng.poisson(np.exp(0.5 * np.sin(0.3 * x))).astype(np.float64)
    y = y.copy()
    y[::23] = np.nan
    cls = {'vi': bn.models.MarkovVariationalGP, 'newton': bn.models.MarkovLaplaceGP}[method]
    olik = {'probit': sites.Bernoulli(), 'logit': sites.Bernoulli(link='logit'), 'gaussian': sites.Gaussian(0.3),
            'poisson': sites.Poisson()}[lik]
    g = cls(kernel=kg, likelihood=_lik(bn, lik), X=x, Y=y, parallel=True)
    assert g._fused_ok()
    o = model.MarkovGP(ko, olik, x, y, method=method)
    os.environ['BN_B200_FUSED'] = '0'
    try:
        u = cls(kernel=kg, likelihood=_lik(bn, lik), X=x, Y=y, parallel=True)
        assert not u._fused_ok()
        for it in range(3):
            u.inference(lr=0.6)
        Eu = float(u.energy())
    finally:
        del os.environ['BN_B200_FUSED']
    for it in range(3):
        _, (d1, d2) = g.inference(lr=0.6)
        _, (d10, d20) = o.inference(lr=0.6)
        assert abs(float(d1) - d10) < TOL * d10 and abs(float(d2) - d20) < TOL * d20
        assert rel_err(np_(g.posterior_mean), o.post_mean) < TOL
        assert rel_err(np_(g.posterior_variance), o.post_cov) < TOL
        E = float(g.energy())
        assert abs(E - o.energy()) <= TOL * abs(o.energy())
    assert rel_err(np_(g.pseudo_likelihood.mean), o.site_mean) < 1e-8  # sites of missing steps carry 1e-6 precisions
    assert rel_err(np_(g.pseudo_likelihood.nat2), np_(u.pseudo_likelihood.nat2)) < TOL
    assert abs(E - Eu) <= 1e-11 * abs(Eu)
    # a hyper-parameter change between iterations is picked up (spec rebuilt on every pass; caches keyed on it)
    from bayesnewton_b200.kernels import softplus_inv
    for mdl in (g, o):
        if mdl is g:
            g.kernel.transformed_lengthscale = softplus_inv(g.kernel.lengthscale * 1.3)
        else:
            o.kernel = type(ko)(ko.variance, ko.lengthscale * 1.3)
    g.inference(lr=0.5)
    o.inference(lr=0.5)
    assert rel_err(np_(g.posterior_mean), o.post_mean) < TOL
    assert abs(float(g.energy()) - o.energy()) <= TOL * abs(o.energy())


@pytest.mark.parametrize('kname', ['m32', 'm52'])
@pytest.mark.parametrize('lik', ['probit', 'logit', 'gaussian', 'poisson'])
def test_model_iteration_fused_ep_vs_oracle(bn, kname, lik, monkeypatch):
    """the EP epilogues of the fused passes (opt-in, BN_B200_FUSED_EP=1): power-EP iterations and energy against the
    oracle model and against the stage-level path the EP models take by default"""
    kg, ko = kernels(bn)[kname]
    N = 400
    x, y = classification_data(N, seed=9)
    rng = np.random.default_rng(4)
    if lik == 'gaussian':
        y = np.sin(0.3 * x) + 0.4 * rng.standard_normal(N)
    elif lik == 'poisson':
        y = rng.poisson(np.exp(0.5 * np.sin(0.3 * x))).astype(np.float64)
    y = y.copy()
    y[::23] = np.nan
    olik = {'probit': sites.Bernoulli(), 'logit': sites.Bernoulli(link='logit'), 'gaussian': sites.Gaussian(0.3),
            'poisson': sites.Poisson()}[lik]
    mk = lambda: bn.models.MarkovExpectationPropagationGP(kernel=kg, likelihood=_lik(bn, lik), X=x, Y=y, power=0.5, parallel=True)
    u = mk()
    assert not u._fused_ok()          # default: stage-level kernels
    monkeypatch.setenv('BN_B200_FUSED_EP', '1')
    g = mk()
    assert g._fused_ok()
    o = model.MarkovGP(ko, olik, x, y, method='ep', power=0.5)
    for it in range(3):
        g.inference(lr=0.6)
        o.inference(lr=0.6)
        assert rel_err(np_(g.posterior_mean), o.post_mean) < TOL and rel_err(np_(g.posterior_variance), o.post_cov) < TOL
        E = float(g.energy())
        assert abs(E - o.energy()) <= TOL * abs(o.energy())
    monkeypatch.delenv('BN_B200_FUSED_EP')
    for it in range(3):
        u.inference(lr=0.6)
    assert abs(E - float(u.energy())) <= 1e-10 * abs(E)
    assert rel_err(np_(g.pseudo_likelihood.nat2), np_(u.pseudo_likelihood.nat2)) < TOL


def test_fused_full_size_c2(bn):
    """N = 1e7 (BASELINE config 2): the fused iteration equals the unfused library path step for step"""
    N = 10_000_000
    t, dt, y = bench_inputs(N)
    mk = lambda: bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(),
                                               X=t, Y=y, parallel=True)
    g = mk()
    g.inference(lr=1.0)
    Eg = float(g.energy())
    pm, pv = np_(g.posterior_mean), np_(g.posterior_variance)
    sm = np_(g.pseudo_likelihood.mean)
    del g
    os.environ['BN_B200_FUSED'] = '0'
    try:
        u = mk()
        u.inference(lr=1.0)
        Eu = float(u.energy())
    finally:
        del os.environ['BN_B200_FUSED']
    assert rel_err(pm, np_(u.posterior_mean)) < 1e-11 and rel_err(pv, np_(u.posterior_variance)) < 1e-11
    assert rel_err(sm, np_(u.pseudo_likelihood.mean)) < 1e-11
    assert abs(Eg - Eu) <= 1e-11 * abs(Eu)
