"""Runs the product's __host__ __device__ kernel bodies on the CPU (tests/hostemu) against the oracle:
arithmetic, chunk indexing, first/last-step rules, carry fold across emulated time shards, and the
fused site update -- everything except the warp-shuffle scan kernel and the launch plumbing, which
only the -m gpu tests can exercise.  Tolerance: 1e-9 normwise (north_star), observed ~1e-15."""
import numpy as np
import pytest

import _emu
from _data import filter_problem, rel_err
from oracle import kalman, sites, ssm

TOL = 1e-9

KERNELS = {
    'm12': (1, lambda: ssm.Matern12(0.8, 1.7), [0.8], [1.7]),
    'm32': (2, lambda: ssm.Matern32(1.1, 0.6), [1.1], [0.6]),
    'm52': (3, lambda: ssm.Matern52(1.3, 0.9), [1.3], [0.9]),
    'm72': (4, lambda: ssm.Matern72(0.7, 1.4), [0.7], [1.4]),
    'ind32': (2, lambda: ssm.Independent([ssm.Matern32(1.0, 1.0), ssm.Matern32(0.5, 2.0)]), [1.0, 0.5], [1.0, 2.0]),
    'ind52': (3, lambda: ssm.Independent([ssm.Matern52(1.3, 0.9), ssm.Matern52(0.7, 2.1)]), [1.3, 0.7], [0.9, 2.1]),
}


@pytest.mark.parametrize('name', sorted(KERNELS))
@pytest.mark.parametrize('form,L,world', [(0, 8, 1), (1, 8, 1), (1, 5, 3), (1, 64, 2), (1, 4, 8)])
def test_filter_and_smoother(emu, name, form, L, world):
    fam, mk, vs, ls = KERNELS[name]
    k = mk()
    D = len(vs)
    N = 203
    dt, y, R, mask = filter_problem(N, D=D, seed=11)
    sp = _emu.spec(fam, vs, ls)
    for rp in (False, True):
        e0, (m0, P0) = kalman.kalman_filter(dt, k, y, R, mask, return_predict=rp)
        e1, m1, P1 = _emu.kalman_filter(emu, sp, form, dt, y, R, mask, L=L, world=world, return_predict=rp)
        assert abs(e1 - e0) < TOL * abs(e0)
        assert rel_err(m1, m0) < TOL and rel_err(P1, P0) < TOL
    _, (fm, fP) = kalman.kalman_filter(dt, k, y, R, mask)
    dts = np.concatenate([dt[1:], [0.0]])
    for rf in (False, True):
        s0 = kalman.rauch_tung_striebel_smoother(dts, k, fm, fP, return_full=rf)
        s1 = _emu.rts_smoother(emu, sp, form, dts, fm, fP, L=L, world=world, return_full=rf)
        for a, b in zip(s1, s0):
            assert rel_err(a, b) < TOL


@pytest.mark.parametrize('name', sorted(KERNELS))
@pytest.mark.parametrize('N', [1, 7, 8, 9, 203, 1000])
@pytest.mark.parametrize('L,world', [(8, 1), (16, 1), (8, 3), (24, 2), (128, 4)])
def test_fused_update_posterior(emu, name, N, L, world):
    """csrc/up_impl.cuh: the fused filter + smoother whose smoothing elements are derived per CHUNK from the
    filter's chunk elements (fast_core.cuh chunk_smoothing_element), incl. the sharded carry path"""
    if world > N:
        pytest.skip('fewer steps than shards')
    fam, mk, vs, ls = KERNELS[name]
    k = mk()
    dt, y, R, mask = filter_problem(N, D=len(vs), seed=11)
    e0, (fm, fP) = kalman.kalman_filter(dt, k, y, R, mask)
    sm, sP, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), k, fm, fP)
    e1, pm, pc = _emu.update_posterior(emu, _emu.spec(fam, vs, ls), dt, y, R, mask, L=L, world=world)
    assert abs(e1 - e0) <= TOL * abs(e0) and rel_err(pm, sm) < TOL and rel_err(pc, sP) < TOL


@pytest.mark.parametrize('lo,hi,rscale', [(1e-3, 5e-3, 1.0), (1e-3, 5e-3, 100.0), (2.0, 5.0, 1.0), (0.1, 0.3, 100.0)])
def test_fused_update_conditioning(emu, lo, hi, rscale):
    """small dt / lengthscale (ill-conditioned Q), long gaps, and the vague initial sites (variance 100)"""
    k = ssm.Matern52(1.3, 1.0)
    dt, y, R, mask = filter_problem(4000, seed=3, dt_lo=lo, dt_hi=hi)
    R = R * rscale
    e0, (fm, fP) = kalman.kalman_filter(dt, k, y, R, mask)
    sm, sP, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), k, fm, fP)
    for L, world in ((8, 1), (128, 1), (128, 4)):
        e1, pm, pc = _emu.update_posterior(emu, _emu.spec(3, [1.3], [1.0]), dt, y, R, mask, L=L, world=world)
        assert abs(e1 - e0) <= TOL * abs(e0) and rel_err(pm, sm) < TOL and rel_err(pc, sP) < TOL


def test_ragged_and_tiny_inputs(emu):
    sp = _emu.spec(3, [1.0], [1.0])
    k = ssm.Matern52(1.0, 1.0)
    for N in (1, 2, 3, 9):
        dt, y, R, mask = filter_problem(N, seed=N)
        e0, (m0, P0) = kalman.kalman_filter(dt, k, y, R, mask)
        for form, L in ((0, 4), (1, 4), (1, 1)):
            e1, m1, P1 = _emu.kalman_filter(emu, sp, form, dt, y, R, mask, L=L)
            assert abs(e1 - e0) <= TOL * abs(e0) + 1e-15 and rel_err(m1, m0) < TOL and rel_err(P1, P0) < TOL
            dts = np.concatenate([dt[1:], [0.0]])
            s0 = kalman.rauch_tung_striebel_smoother(dts, k, m0, P0)
            s1 = _emu.rts_smoother(emu, sp, form, dts, m0, P0, L=L)
            assert all(rel_err(a, b) < TOL for a, b in zip(s1, s0))


def test_duplicate_time_stamps_stay_finite_in_scan_form(emu):
    """dt = 0 mid-sequence gives Q = 0; the reference's scan inverts C and returns NaN, the
    sequential form is fine.  The blocked scan solves with (I + C J) instead and must match the sequential result."""
    sp = _emu.spec(3, [1.0], [1.0])
    k = ssm.Matern52(1.0, 1.0)
    dt, y, R, mask = filter_problem(64, seed=5, missing=0.0)
    dt[[7, 8, 30]] = 0.0
    e0, (m0, P0) = kalman.kalman_filter(dt, k, y, R, mask)
    e1, m1, P1 = _emu.kalman_filter(emu, sp, 1, dt, y, R, mask, L=4)
    assert np.isfinite(m1).all() and rel_err(m1, m0) < TOL and rel_err(P1, P0) < TOL and abs(e1 - e0) < TOL * abs(e0)


SITE_CASES = [('probit', lambda: sites.Bernoulli('probit'), 0.0), ('logit', lambda: sites.Bernoulli('logit'), 0.0),
              ('gaussian', lambda: sites.Gaussian(0.3), 0.3), ('poisson', lambda: sites.Poisson(1.5), 1.5),
              ('studentst', lambda: sites.StudentsT(0.7, 4.0), (0.7, 4.0)), ('gamma', lambda: sites.Gamma(1.3), 1.3),
              ('negbin', lambda: sites.NegativeBinomial(0.6, 1.5), (0.6, 1.5)), ('beta', lambda: sites.Beta(3.0), 3.0)]


def site_observations(likname, rng, N):
    return {'gaussian': lambda: rng.standard_normal(N), 'studentst': lambda: 1.5 * rng.standard_t(4.0, size=N),
            'poisson': lambda: rng.poisson(1.5, size=N).astype(float), 'gamma': lambda: rng.gamma(1.3, 1.3, size=N) + 1e-3,
            'negbin': lambda: rng.negative_binomial(2, 0.4, size=N).astype(float),
            'beta': lambda: np.clip(rng.beta(1.5, 2.0, size=N), 1e-3, 1 - 1e-3)}.get(
                likname, lambda: (rng.uniform(size=N) < 0.5).astype(float))()


@pytest.mark.parametrize('likname,mk,lp', SITE_CASES)
@pytest.mark.parametrize('method', ['vi', 'ep', 'newton', 'pl'])
@pytest.mark.parametrize('lr,power', [(1.0, 1.0), (0.4, 0.5)])
def test_site_update_single_latent(emu, likname, mk, lp, method, lr, power):
    rng = np.random.default_rng(1)
    N = 300
    lik = mk()
    y = site_observations(likname, rng, N)
    y[::17] = np.nan
    lp, lp2 = lp if isinstance(lp, tuple) else (lp, 0.0)
    pm = rng.standard_normal((N, 1, 1)); pc = 0.2 + rng.uniform(size=(N, 1, 1))
    n2 = 0.01 + 0.25 * rng.uniform(size=(N, 1, 1)) / pc; n1 = 0.3 * rng.standard_normal((N, 1, 1))
    mean, jac, hess = sites.site_statistics(method, lik, y[:, None], pm, pc, n1, n2, power=power,
                                            mask_pseudo_y=np.isnan(y)[:, None])
    o = sites.damped_site_update(n1, n2, mean, jac, hess, lr)
    e = _emu.site_update(emu, method, likname, lp, y, pm, pc, n1, n2, lr=lr, power=power,
                         cub=sites.gauss_hermite(1, 20), lik_param2=lp2)
    for got, ref in [(e['mean'], mean), (e['jac'], jac), (e['hess'], hess), (e['nat1'], o[0]), (e['nat2'], o[1]),
                     (e['site_mean'], o[2]), (e['site_cov'], o[3])]:
        assert rel_err(got, ref) < TOL
    assert abs(e['diffs'][0] - o[4]) < TOL * o[4] and abs(e['diffs'][1] - o[5]) < TOL * o[5]


@pytest.mark.parametrize('method', ['vi', 'ep', 'newton'])
@pytest.mark.parametrize('lr,power,psd', [(1.0, 1.0, True), (0.3, 0.5, True), (0.3, 0.5, False)])
def test_site_update_heteroscedastic(emu, method, lr, power, psd):
    rng = np.random.default_rng(2)
    N = 200
    lik = sites.HeteroscedasticNoise('softplus')
    y = rng.standard_normal(N)
    pm = rng.standard_normal((N, 2, 1)) * 0.5
    A = rng.standard_normal((N, 2, 2)) * 0.3
    pc = A @ np.swapaxes(A, 1, 2) + 0.2 * np.eye(2)
    n2 = np.zeros((N, 2, 2)); n2[:, 0, 0] = 0.01 + 0.3 * rng.uniform(size=N); n2[:, 1, 1] = 0.01 + 0.3 * rng.uniform(size=N)
    n1 = 0.3 * rng.standard_normal((N, 2, 1))
    mean, jac, hess = sites.site_statistics(method, lik, y[:, None], pm, pc, n1, n2, power=power, ensure_psd=psd)
    o = sites.damped_site_update(n1, n2, mean, jac, hess, lr)
    e = _emu.site_update(emu, method, 'het_softplus', 0, y, pm, pc, n1, n2, lr=lr, power=power,
                         cub=sites.gauss_hermite(2, 20), ensure_psd=psd)
    for got, ref in [(e['mean'], mean), (e['jac'], jac), (e['hess'], hess), (e['nat1'], o[0]), (e['nat2'], o[1]),
                     (e['site_mean'], o[2]), (e['site_cov'], o[3])]:
        assert rel_err(got, ref) < 1e-8  # the EP scale factor inverts ill-conditioned 2x2s: 1e-13 typical


# ---------------------------------------------------------------------------- fused iteration on tiled state (iter_impl.cuh)
ITER_KERNELS = {'m12': (1, lambda: ssm.Matern12(0.8, 1.7), 0.8, 1.7), 'm32': (2, lambda: ssm.Matern32(1.1, 0.6), 1.1, 0.6),
                'm52': (3, lambda: ssm.Matern52(1.3, 0.9), 1.3, 0.9), 'm72': (4, lambda: ssm.Matern72(0.7, 1.4), 0.7, 1.4)}
ITER_LIKS = {'probit': (lambda: sites.Bernoulli('probit'), 0.0), 'logit': (lambda: sites.Bernoulli('logit'), 0.0),
             'gaussian': (lambda: sites.Gaussian(0.3), 0.3), 'poisson': (lambda: sites.Poisson(1.0), 1.0)}


def _iter_problem(N, likname, seed):
    rng = np.random.default_rng(seed)
    dt = np.concatenate([[0.0], 0.1 + 0.2 * rng.random(N - 1)]) if N > 1 else np.zeros(1)
    y = site_observations(likname, rng, N)
    if N > 20:
        y[::17] = np.nan
    sy = 0.3 * rng.standard_normal(N)
    sR = 0.5 + rng.random(N)
    return dt, y, sy, sR


def _oracle_pass(k, lik, method, dt, y, sy, sR, lr, power=1.0):
    """update_posterior, then the site update of the scheme and the two energy sums, with the oracle"""
    N = dt.shape[0]
    mask = np.isnan(y).reshape(N, 1, 1)
    ell, (fm, fP) = kalman.kalman_filter(dt, k, sy.reshape(N, 1, 1), sR.reshape(N, 1, 1), mask)
    pm, pc, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), k, fm, fP)
    n2 = 1.0 / sR.reshape(N, 1, 1)
    n1 = sy.reshape(N, 1, 1) * n2
    mean, jac, hess = sites.site_statistics(method, lik, y[:, None], pm, pc, n1, n2, power=power, mask_pseudo_y=np.isnan(y)[:, None])
    new = sites.damped_site_update(n1, n2, mean, jac, hess, lr)
    return ell, pm.reshape(-1), pc.reshape(-1), new[2].reshape(-1), new[3].reshape(-1), (new[4] * N, new[5] * N)


@pytest.mark.parametrize('kname', sorted(ITER_KERNELS))
@pytest.mark.parametrize('likname,method', [('probit', 'vi'), ('probit', 'newton'), ('logit', 'vi'), ('gaussian', 'vi'), ('poisson', 'vi'),
                                            ('probit', 'ep'), ('gaussian', 'ep'), ('poisson', 'ep')])
@pytest.mark.parametrize('N,L,world,spec', [(1, 8, 1, False), (7, 8, 1, False), (203, 8, 3, False), (1500, 8, 1, False),
                                            (2500, 256, 2, True), (4000, 8, 1, True)])
def test_fused_iteration_passes(emu, kname, likname, method, N, L, world, spec):
    """csrc/iter_impl.cuh chunk bodies on the host: the SITES pass (filter, smoother, site update in the epilogue), the
    PLAIN / ENERGY passes (marginals, log-likelihood), time shards, the speculative phase 1, and the three-part prefix of
    the scan (the emulation groups 3 warps, so 4000 steps in 8-step chunks reach all three parts)"""
    fam, mk, var, ls = ITER_KERNELS[kname]
    lik, lp = ITER_LIKS[likname][0](), ITER_LIKS[likname][1]
    dt, y, sy, sR = _iter_problem(N, likname, seed=N + L)
    sp = _emu.spec(fam, [var], [ls])
    cub = sites.gauss_hermite(1, 20)
    mask = np.isnan(y).astype(np.uint8) if np.isnan(y).any() else None
    power = 0.5 if method == 'ep' else 1.0
    ell0, pm0, pc0, nsy, nsR, diffs = _oracle_pass(mk(), lik, method, dt, y, sy, sR, lr=0.6, power=power)
    out = _emu.iter_pass(emu, sp, dt, y, sy, sR, 0, method, likname, lp, cub, mask, L=L, world=world, spec=spec)     # PLAIN
    assert rel_err(out['post_mean'], pm0) < TOL and rel_err(out['post_cov'], pc0) < TOL
    assert abs(out['ell'] - ell0) <= TOL * abs(ell0)
    out = _emu.iter_pass(emu, sp, dt, y, sy, sR, 1, method, likname, lp, cub, mask, lr=0.6, power=power, L=L, world=world,
                         spec=spec)  # SITES
    assert rel_err(out['site_mean'], nsy) < 10 * TOL and rel_err(1.0 / out['site_cov'], 1.0 / nsR) < TOL
    assert np.allclose(out['sums'], diffs, rtol=1e-8)
    # ENERGY: the likelihood term of the scheme's energy, summed in the smoother epilogue, against the stand-alone kernel
    N1 = dt.shape[0]
    n2 = 1.0 / sR.reshape(N1, 1, 1)
    n1 = sy.reshape(N1, 1, 1) * n2
    vals, tot = _emu.expected_density(emu, method, likname, lp, y, pm0.reshape(N1, 1, 1), pc0.reshape(N1, 1, 1), n1, n2, power=power, cub=cub)
    oe = _emu.iter_pass(emu, sp, dt, y, sy, sR, 2, method, likname, lp, cub, mask, power=power, L=L, world=world, spec=spec)
    assert abs(oe['sums'][0] - tot) <= 1e-8 * max(1.0, abs(tot))
    assert rel_err(oe['post_mean'], pm0) < TOL and abs(oe['ell'] - ell0) <= TOL * abs(ell0)
    if spec and L >= 256 and kname in ('m32', 'm52'):  # (the other two forget their start more slowly than 256 steps)
        assert out['jstar_mean'] < 0.75 * L   # the chunks did switch to the plain filter well before their end
