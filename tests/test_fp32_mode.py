"""fp32 mode of the fused iteration (bn_iter_*_f32: the kernels of csrc/iter_impl.cuh compiled with a float scalar
type, real.cuh).  Parity bar of the mode: 1e-4 relative (normwise) against fp64 results -- the fp64 ORACLE on C1 / C2
shaped problems at sizes it finishes in seconds, and the library's own fp64 pass at 10^6 and 10^7 steps."""
import numpy as np
import pytest

from _data import bench_inputs, rel_err
from oracle import model, sites, ssm

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy().astype(np.float64)


def kernels(bn):
    K, O = bn.kernels, ssm
    return {'m12': (K.Matern12(0.8, 1.7), O.Matern12(0.8, 1.7)), 'm32': (K.Matern32(1.1, 0.6), O.Matern32(1.1, 0.6)),
            'm52': (K.Matern52(1.3, 0.9), O.Matern52(1.3, 0.9)), 'm72': (K.Matern72(0.7, 1.4), O.Matern72(0.7, 1.4))}


def liks(bn):
    L = bn.likelihoods
    return {'gaussian': (L.Gaussian(0.3), sites.Gaussian(0.3)), 'probit': (L.Bernoulli('probit'), sites.Bernoulli('probit')),
            'logit': (L.Bernoulli('logit'), sites.Bernoulli('logit')), 'poisson': (L.Poisson(), sites.Poisson())}


def observations(lik, t, y, seed):
    rng = np.random.default_rng(seed)
    if lik == 'gaussian':
        return np.sin(0.3 * t) + 0.4 * rng.standard_normal(t.shape[0])
    if lik == 'poisson':
        return rng.poisson(np.exp(0.5 * np.sin(0.3 * t))).astype(np.float64)
    return y


def iterate(bn, kern, lik, method, dt, y, iters, dtype, lr=0.7):
    """`iters` x (site pass, energy pass) through FusedShard in the given precision"""
    import torch
    from bayesnewton_b200 import _lib, fused
    meth = {'vi': _lib.BN_METHOD_VI, 'newton': _lib.BN_METHOD_NEWTON}[method]
    dev = torch.device('cuda')
    sh = fused.FusedShard(kern, torch.as_tensor(dt, device=dev), torch.as_tensor(y, device=dev), dtype=dtype)
    N = dt.shape[0]
    sh.load_sites(torch.zeros(N, device=dev), torch.full((N,), 100.0, device=dev))  # the reference's initial sites
    out = []
    for _ in range(iters):
        _, d = sh.run(fused.SITES, lik, meth, None, lr, 1.0, True, want_ell=False)
        ell, s = sh.run(fused.ENERGY, lik, meth, None, lr, 1.0, True)
        out.append((float(ell), np_(s), np_(d)))
    pm, pc = sh.posterior()
    sm, sc = sh.sites()
    return out, np_(pm).reshape(-1), np_(pc).reshape(-1), np_(sm).reshape(-1), np_(sc).reshape(-1)


@pytest.mark.parametrize('kname', ['m12', 'm32', 'm52', 'm72'])
@pytest.mark.parametrize('lik,method', [('gaussian', 'vi'), ('probit', 'vi'), ('probit', 'newton'), ('logit', 'vi'),
                                        ('poisson', 'vi')])
def test_fp32_iterations_vs_fp64_oracle(bn, kname, lik, method):
    """C1 (Gaussian) / C2 (probit) shaped problems, three iterations, against the fp64 oracle model"""
    import torch
    N = 3001
    kg, ko = kernels(bn)[kname]
    lg, lo = liks(bn)[lik]
    t, dt, y = bench_inputs(N, 3)
    y = observations(lik, t, y, 11)
    o = model.MarkovGP(ko, lo, t, y, method=method, parallel=False)
    energies = []
    for _ in range(3):
        o.inference(lr=0.7)
        energies.append(o.energy())
    out, pm, pc, sm, sc = iterate(bn, kg, lg, method, dt, y, 3, torch.float32)
    assert rel_err(pm, o.post_mean.reshape(-1)) < TOL32 and rel_err(pc, o.post_cov.reshape(-1)) < TOL32
    assert rel_err(sm, o.site_mean.reshape(-1)) < 10 * TOL32 and rel_err(1.0 / sc, 1.0 / o.site_cov.reshape(-1)) < TOL32
    ell = o.compute_log_lik()
    assert abs(out[-1][0] - ell) <= TOL32 * abs(ell)


@pytest.mark.parametrize('N', [1, 7, 300, 70_001])
def test_fp32_ragged_sizes_and_missing_data(bn, N):
    import torch
    kg, _ = kernels(bn)['m52']
    lg, _ = liks(bn)['probit']
    t, dt, y = bench_inputs(N, 5)
    if N > 20:
        y[::13] = np.nan
    o64 = iterate(bn, kg, lg, 'vi', dt, y, 2, torch.float64)
    o32 = iterate(bn, kg, lg, 'vi', dt, y, 2, torch.float32)
    for a, b in zip(o32[1:], o64[1:]):
        assert np.isfinite(a).all() and rel_err(a, b) < 5 * TOL32
    assert abs(o32[0][-1][0] - o64[0][-1][0]) <= TOL32 * abs(o64[0][-1][0])


@pytest.mark.parametrize('N', [1_000_000, 10_000_000])
def test_fp32_large_series_vs_fp64_pass(bn, N):
    """C2 at full size: two iterations in fp32 against the same two iterations in fp64 (the fp64 pass itself is held to
    1e-9 against the oracle elsewhere); the sums over 10^7 steps are accumulated in fp64 in both builds"""
    import torch
    kg, _ = kernels(bn)['m52']
    lg, _ = liks(bn)['probit']
    t, dt, y = bench_inputs(N, 0)
    o64 = iterate(bn, kg, lg, 'vi', dt, y, 2, torch.float64, lr=1.0)
    o32 = iterate(bn, kg, lg, 'vi', dt, y, 2, torch.float32, lr=1.0)
    assert rel_err(o32[1], o64[1]) < TOL32 and rel_err(o32[2], o64[2]) < TOL32      # posterior mean / variance
    assert rel_err(1.0 / o32[4], 1.0 / o64[4]) < TOL32                                 # site precisions
    for (e32, s32, _), (e64, s64, _) in zip(o32[0], o64[0]):
        assert abs(e32 - e64) <= TOL32 * abs(e64)
        assert np.allclose(s32, s64, rtol=TOL32)


@pytest.mark.parametrize('lik,method', [('probit', 'vi'), ('gaussian', 'vi'), ('probit', 'newton')])
def test_fp32_model_level_precision_switch(bn, lik, method):
    """model.set_precision('float32'): the reference-shaped model API on the fp32 build; iterations, energy, prediction
    (which runs outside the fused iteration, in fp64, from the fp32 sites) against the same model in fp64"""
    N = 20_011
    t, dt, y = bench_inputs(N, 7)
    y = observations(lik, t, y, 3)
    y[::19] = np.nan
    cls = bn.models.MarkovVariationalGP if method == 'vi' else bn.models.MarkovLaplaceGP
    mk = lambda: cls(kernel=bn.kernels.Matern52(1.3, 0.9), likelihood=liks(bn)[lik][0], X=t, Y=y, parallel=True)
    m64, m32 = mk(), mk()
    m32.set_precision('float32')
    for _ in range(3):
        m64.inference(lr=0.7)
        m32.inference(lr=0.7)
        E64, E32 = float(m64.energy()), float(m32.energy())
        assert abs(E32 - E64) <= TOL32 * abs(E64)
    import torch
    assert m32.posterior_mean.dtype == torch.float32
    assert rel_err(np_(m32.posterior_mean), np_(m64.posterior_mean)) < TOL32
    assert rel_err(np_(m32.posterior_variance), np_(m64.posterior_variance)) < TOL32
    assert rel_err(np_(m32.pseudo_likelihood.nat2), np_(m64.pseudo_likelihood.nat2)) < TOL32
    xt = np.linspace(t[0] - 1.0, t[-1] + 1.0, 41)
    p64, v64 = m64.predict(X=xt)
    p32, v32 = m32.predict(X=xt)
    assert rel_err(np_(p32), np_(p64)) < TOL32 and rel_err(np_(v32), np_(v64)) < TOL32
    m32.set_precision('float64')  # and back: the sites carry over
    m32.inference(lr=0.7)
    m64.inference(lr=0.7)
    assert rel_err(np_(m32.posterior_mean), np_(m64.posterior_mean)) < TOL32
