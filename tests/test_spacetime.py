"""Dense spatio-temporal path.

CPU part: pins the spatio-temporal oracle the way the reference pins its own implementation -- the Markov model
against the dense GP on the product kernel, on the parameter grid of tests/test_gp_vs_markovgp_spacetime.py:62-65
(seeded data; energies to 1e-6 relative instead of the reference's 2 decimals) -- and against the frozen vectors
in tests/golden/spacetime_small.npz.
GPU part: the CUDA kernels through the C ABI against the oracle.  Tolerance: relative 1e-9 (north_star), normwise.
"""
import os

import numpy as np
import pytest

from _data import rel_err
from oracle import kalman, sites, ssm
from oracle import spacetime as ost

TOL = 1e-9
# Missing observations enter the sites as precision 1e-6 (newton_update, inference.py:30), which makes
# nat2_full = B^T nat2 B ill-conditioned (measured 8.6e8 on the 4 x 4 case below): a 1-ulp perturbation of the
# inputs already moves the posterior mean by 2e-9, so two correct fp64 implementations agree to ~cond * eps only.
TOL_MISSING = 1e-7
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'spacetime_small.npz')


def st_data(Nt, Ns, seed=0, spatial_dims=1, missing=0.0):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 1, Nt)) * (1.0 + 0.3 * Nt)
    if spatial_dims == 1:
        r = np.linspace(0, 1, Ns)[:, None]
    else:
        g = int(round(Ns ** 0.5))
        assert g * g == Ns
        a = np.linspace(-1, 1, g)
        r = np.array([[u, v] for u in a for v in a])
    R = np.tile(r[None], (Nt, 1, 1))
    Y = np.sin(3 * t)[:, None] + np.sin(4 * r[:, 0])[None, :] + 0.1 * rng.standard_normal((Nt, Ns))
    if missing > 0:
        Y[rng.uniform(size=Y.shape) < missing] = np.nan
    return t, Y, R


def oracle_kernel(fam, var_f, len_t, len_s, z, spatial_dims=1):
    kt = getattr(ssm, fam)(var_f, len_t)
    ks = ssm.Matern32(1.0, len_s) if spatial_dims == 1 else ost.Separable([ssm.Matern32(1.0, len_s)] * spatial_dims)
    return ost.SpatioTemporalKernel(kt, ks, z=z)


# ------------------------------------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize('var_f', [0.5, 1.5])
@pytest.mark.parametrize('len_f', [0.75, 2.5])
@pytest.mark.parametrize('var_y', [0.1, 0.5])
@pytest.mark.parametrize('N', [8, 16])
def test_oracle_markov_vs_dense_gp(var_f, len_f, var_y, N):
    t, Y, R = st_data(N, N, seed=N)
    mk = lambda: ost.SpatioTemporalKernel(ssm.Matern52(var_f, len_f), ssm.Matern52(1.0, len_f), z=R[0])
    lik = sites.Gaussian(var_y)
    m = ost.SpatioTemporalMarkovGP(mk(), lik, t, Y, R)
    g = ost.DenseSpatioTemporalGP(mk(), lik, t, Y, R)
    m.update_posterior()
    g.update_posterior()
    assert abs(m.energy() - g.energy()) <= 1e-6 * abs(g.energy())
    m.inference()
    g.inference()
    assert abs(m.energy() - g.energy()) <= 1e-6 * abs(g.energy())
    mf, cf = m.conditional_posterior_to_data()
    assert np.abs(mf.reshape(-1) - g.post_mean).max() < 1e-6
    assert np.abs(np.diagonal(cf, axis1=1, axis2=2).reshape(-1) - g.post_var).max() < 1e-6


def golden_case():
    t, Y, R = st_data(12, 9, seed=3, spatial_dims=2, missing=0.1)
    k = oracle_kernel('Matern32', 1.2, 0.8, 0.9, R[0], spatial_dims=2)
    m = ost.SpatioTemporalMarkovGP(k, sites.Gaussian(0.3), t, Y, R)
    m.inference(lr=0.7)
    return m, dict(t=t, Y=Y, R=R, post_mean=m.post_mean, post_cov=m.post_cov, energy=np.array(m.energy()),
                   site_nat1=m.site_nat1, site_nat2=m.site_nat2)


def test_oracle_matches_golden():
    _, out = golden_case()
    ref = np.load(GOLDEN)
    for k in ('post_mean', 'post_cov', 'energy', 'site_nat1', 'site_nat2'):
        assert rel_err(out[k], ref[k]) < 1e-10, k


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy()


def gpu_kernel(bn, fam, var_f, len_t, len_s, z, spatial_dims=1):
    K = bn.kernels
    kt = getattr(K, fam)(var_f, len_t)
    ks = K.Matern32(1.0, len_s) if spatial_dims == 1 else bn.spacetime.Separable([K.Matern32(1.0, len_s)] * spatial_dims)
    return bn.spacetime.SpatioTemporalKernel(kt, ks, z=z)


def spd_batch(N, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, n, n))
    return A @ A.transpose(0, 2, 1) / n + 0.5 * np.eye(n)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 5, 32, 33, 100, 256])
def test_gpu_spd_inverse_batched(bn, n):
    N = 7
    P = spd_batch(N, n, n)
    rhs = np.random.default_rng(n + 1).standard_normal((N, n, 1))
    inv, sol, ld = bn.spacetime.inv_vmap(P, rhs=rhs, want_logdet=True)
    ref = np.linalg.inv(P)
    assert rel_err(np_(inv), ref) < TOL
    assert rel_err(np_(sol), ref @ rhs) < TOL
    assert rel_err(np_(ld), np.linalg.slogdet(P)[1]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('fam', ['Matern12', 'Matern32', 'Matern52'])
@pytest.mark.parametrize('M', [3, 16, 37])
@pytest.mark.parametrize('masked', [False, True])
def test_gpu_dense_filter_smoother_vs_oracle(bn, fam, M, masked):
    N = 11
    rng = np.random.default_rng(M)
    z = np.linspace(0, 1, M)[:, None]
    ko = oracle_kernel(fam, 1.3, 0.7, 0.5, z)
    kg = gpu_kernel(bn, fam, 1.3, 0.7, 0.5, z)
    dt = np.concatenate([[0.0], 0.1 + 0.4 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, M, 1))
    Rn = spd_batch(N, M, M + 1)
    mask = (rng.uniform(size=(N, M, 1)) < 0.2) if masked else None
    for rp in (False, True):
        e0, (m0, P0) = kalman.kalman_filter(dt, ko, y, Rn, mask, return_predict=rp)
        e1, (m1, P1) = bn.ops.kalman_filter(dt, kg, y, Rn, mask, return_predict=rp)
        assert abs(float(e1) - e0) <= TOL * abs(e0)
        assert rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    _, (fm, fP) = kalman.kalman_filter(dt, ko, y, Rn, mask)
    dts = np.concatenate([dt[1:], [0.0]])
    for full in (False, True):
        s0, S0, G0 = kalman.rauch_tung_striebel_smoother(dts, ko, fm, fP, return_full=full)
        s1, S1, G1 = bn.ops.rauch_tung_striebel_smoother(dts, kg, fm, fP, return_full=full)
        assert rel_err(np_(s1), s0) < TOL and rel_err(np_(S1), S0) < TOL and rel_err(np_(G1), G0) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('Ns,M', [(9, 9), (16, 9), (25, 25)])
def test_gpu_projection_steps_vs_oracle(bn, Ns, M):
    """compute_full_pseudo_lik, conditional_posterior_to_data and the Gaussian KL term"""
    import torch
    Nt = 6
    t, Y, R = st_data(Nt, Ns, seed=Ns, spatial_dims=2)
    z = R[0] if M == Ns else st_data(1, M, spatial_dims=2)[2][0]
    ko = oracle_kernel('Matern32', 0.9, 1.1, 0.8, z, spatial_dims=2)
    kg = gpu_kernel(bn, 'Matern32', 0.9, 1.1, 0.8, z, spatial_dims=2)
    rng = np.random.default_rng(5)
    mo = ost.SpatioTemporalMarkovGP(ko, sites.Gaussian(0.2), t, Y, R)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.2), X=t, Y=Y, R=R)
    nat1 = rng.standard_normal((Nt, Ns))
    nat2 = 0.5 + rng.uniform(size=(Nt, Ns))
    mo.site_nat1 = nat1[..., None]
    mo.site_nat2 = np.stack([np.diag(v) for v in nat2])
    mg.pseudo_likelihood.nat1_.copy_(torch.as_tensor(nat1))
    mg.pseudo_likelihood.nat2_.copy_(torch.as_tensor(nat2))
    mg.pseudo_likelihood.version += 1
    py0, pv0 = mo.compute_full_pseudo_lik()
    py1, pv1 = mg.compute_full_pseudo_lik()
    assert rel_err(np_(py1), py0) < TOL and rel_err(np_(pv1), pv0) < TOL
    pm = rng.standard_normal((Nt, M, 1))
    pV = spd_batch(Nt, M, 11)
    mo.post_mean, mo.post_cov = pm, pV
    mf0, cf0 = mo.conditional_posterior_to_data()
    mf1, vf1 = mg.conditional_posterior_to_data(post_mean=pm, post_cov=pV)
    assert rel_err(np_(mf1), mf0) < TOL
    assert rel_err(np_(vf1)[..., 0], np.diagonal(cf0, axis1=1, axis2=2)) < TOL
    mg.posterior_mean, mg.posterior_variance = torch.as_tensor(pm).cuda(), torch.as_tensor(pV).cuda()
    e0 = np.sum(sites.gaussian_expected_log_lik(py0, pm, pV, pv0, None))
    e1 = float(mg.expected_density_pseudo())
    assert abs(e1 - e0) <= TOL * abs(e0)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['full', 'sparse', 'missing', 'matern52_time'])
def test_gpu_spatiotemporal_vi_iteration_vs_oracle(bn, case):
    """MarkovVariationalGP + SpatioTemporalKernel: inference(lr) and energy() end to end"""
    Nt, Ns = 14, 16
    t, Y, R = st_data(Nt, Ns, seed=7, spatial_dims=2, missing=0.15 if case == 'missing' else 0.0)
    z = st_data(1, 9, spatial_dims=2)[2][0] if case == 'sparse' else R[0]
    fam = 'Matern52' if case == 'matern52_time' else 'Matern32'
    ko = oracle_kernel(fam, 1.1, 0.9, 1.2, z, spatial_dims=2)
    kg = gpu_kernel(bn, fam, 1.1, 0.9, 1.2, z, spatial_dims=2)
    mo = ost.SpatioTemporalMarkovGP(ko, sites.Gaussian(0.3), t, Y, R)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.3), X=t, Y=Y, R=R)
    tol = TOL_MISSING if case == 'missing' else TOL
    for lr in (1.0, 0.5):
        mo.inference(lr=lr)
        mg.inference(lr=lr)
        E0, E1 = mo.energy(), float(mg.energy())
        assert rel_err(np_(mg.posterior_mean), mo.post_mean) < tol
        assert rel_err(np_(mg.posterior_variance), mo.post_cov) < tol
        assert abs(E1 - E0) <= tol * abs(E0), (E0, E1)
    assert rel_err(np_(mg.pseudo_likelihood.nat2), mo.site_nat2) < tol
    assert rel_err(np_(mg.pseudo_likelihood.nat1), mo.site_nat1) < tol


@pytest.mark.gpu
def test_gpu_golden_spacetime(bn):
    ref = np.load(GOLDEN)
    t, Y, R = ref['t'], ref['Y'], ref['R']
    kg = gpu_kernel(bn, 'Matern32', 1.2, 0.8, 0.9, R[0], spatial_dims=2)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.3), X=t, Y=Y, R=R)
    mg.inference(lr=0.7)
    # the golden case has missing observations
    assert rel_err(np_(mg.posterior_mean), ref['post_mean']) < TOL_MISSING
    assert rel_err(np_(mg.posterior_variance), ref['post_cov']) < TOL_MISSING
    assert abs(float(mg.energy()) - float(ref['energy'])) <= TOL_MISSING * abs(float(ref['energy']))


def c4_data(Nt, G=16, missing=0.05):
    """16 x 16 grid on [-3, 3]^2, unit time steps, 5 % missing observations"""
    t = np.arange(Nt, dtype=np.float64)
    a = np.linspace(-3, 3, G)
    r = np.array([[u, v] for u in a for v in a])
    R = np.tile(r[None], (Nt, 1, 1))
    Y = (np.sin(t / 10)[:, None] + np.sin(r[:, 0])[None] + np.cos(r[:, 1])[None]
         + 0.1 * np.random.default_rng(1).standard_normal((Nt, G * G)))
    if missing:
        Y[np.random.default_rng(2).uniform(size=Y.shape) < missing] = np.nan
    return t, Y, R


class _LUInverseOracle(ost.SpatioTemporalMarkovGP):
    """the same model with the M x M inverse of compute_full_pseudo_lik taken by LU instead of Cholesky: the spread
    between two correct fp64 algorithms = the rounding floor of the problem"""

    def compute_full_pseudo_lik(self):
        nat1_full = np.swapaxes(self.B, 1, 2) @ self.site_nat1
        nat2_full = np.swapaxes(self.B, 1, 2) @ self.site_nat2 @ self.B
        pv = np.linalg.inv(nat2_full + 1e-12 * np.eye(self.M))
        pv = 0.5 * (pv + np.swapaxes(pv, 1, 2))
        return pv @ nat1_full, pv


@pytest.mark.gpu
@pytest.mark.parametrize('missing', [0.0, 0.05])
def test_gpu_c4_shape_iteration(bn, missing):
    """the C4 configuration (16 x 16 grid, Matern-3/2 in time and space, d = 512) on a short horizon against the
    oracle.  Without missing data the bar is 1e-9.  With 5 % missing observations nat2_full has condition number
    5e9 (sites of missing points have precision 1e-6) and two correct fp64 implementations differ by ~5e-8 in the
    posterior mean, so the bar is 20x the measured Cholesky-vs-LU spread of the oracle itself."""
    Nt = 5
    t, Y, R = c4_data(Nt, missing=missing)
    z = R[0]
    ko = oracle_kernel('Matern32', 1.0, 5.0, 1.0, z, spatial_dims=2)
    kg = gpu_kernel(bn, 'Matern32', 1.0, 5.0, 1.0, z, spatial_dims=2)
    mo = ost.SpatioTemporalMarkovGP(ko, sites.Gaussian(1.0), t, Y, R)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(1.0), X=t, Y=Y, R=R)
    mo.inference(lr=1.0)
    mg.inference(lr=1.0)
    E0, E1 = mo.energy(), float(mg.energy())
    tm = tv = te = TOL
    if missing:
        ma = _LUInverseOracle(ko, sites.Gaussian(1.0), t, Y, R)
        ma.inference(lr=1.0)
        tm = max(TOL, 20 * rel_err(ma.post_mean, mo.post_mean))
        tv = max(TOL, 20 * rel_err(ma.post_cov, mo.post_cov))
        te = max(TOL, 20 * abs(ma.energy() - E0) / abs(E0))
    assert rel_err(np_(mg.posterior_mean), mo.post_mean) < tm
    assert rel_err(np_(mg.posterior_variance), mo.post_cov) < tv
    assert abs(E1 - E0) <= te * abs(E0), (E0, E1)


# ------------------------------------------------------------------------------------------------ mean-field (8f row 2)
def test_oracle_meanfield_is_exact_for_diagonal_noise():
    """no reference test covers the mean-field filter; pinned by the case where it is exact: with a diagonal
    noise covariance the blocks never couple, so it must equal M independent filters and the full filter"""
    rng = np.random.default_rng(0)
    M, N = 5, 12
    k = oracle_kernel('Matern32', 1.3, 0.7, 0.5, np.linspace(0, 1, M)[:, None])
    dt = np.concatenate([[0.0], 0.1 + 0.4 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, M, 1))
    Rn = np.stack([np.diag(0.3 + rng.uniform(size=M)) for _ in range(N)])
    ell, (fm, fP) = ost.kalman_filter_meanfield(dt, k, y, Rn)
    e_ind = 0.0
    for i in range(M):
        e_i, (m_i, P_i) = kalman.kalman_filter(dt, k.temporal_kernel, y[:, i:i + 1], Rn[:, i:i + 1, i:i + 1])
        e_ind += e_i
        assert rel_err(fm[:, i], m_i) < 1e-12 and rel_err(fP[:, i], P_i) < 1e-12
    assert abs(ell - e_ind) < 1e-10 * abs(ell)
    dts = np.concatenate([dt[1:], [0.0]])
    sm, sP, _ = ost.rts_smoother_meanfield(dts, k, fm, fP)
    _, (fmf, fPf) = kalman.kalman_filter(dt, k, y, Rn)
    smf, sPf, _ = kalman.rauch_tung_striebel_smoother(dts, k, fmf, fPf)
    assert rel_err(sm, smf) < 1e-12 and rel_err(sP, sPf) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('fam', ['Matern12', 'Matern32', 'Matern52'])
@pytest.mark.parametrize('M', [3, 16, 37])
@pytest.mark.parametrize('masked', [False, True])
def test_gpu_meanfield_filter_smoother_vs_oracle(bn, fam, M, masked):
    N = 11
    rng = np.random.default_rng(M)
    z = np.linspace(0, 1, M)[:, None]
    ko = oracle_kernel(fam, 1.3, 0.7, 0.5, z)
    kg = gpu_kernel(bn, fam, 1.3, 0.7, 0.5, z)
    dt = np.concatenate([[0.0], 0.1 + 0.4 * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, M, 1))
    Rn = spd_batch(N, M, M + 1)
    mask = (rng.uniform(size=(N, M, 1)) < 0.2) if masked else None
    e0, (m0, P0) = ost.kalman_filter_meanfield(dt, ko, y, Rn, mask)
    e1, (m1, P1) = bn.spacetime.st_kalman_filter_meanfield(dt, kg, y, Rn, mask)
    assert abs(float(e1) - e0) <= TOL * abs(e0)
    assert rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    dts = np.concatenate([dt[1:], [0.0]])
    for full in (False, True):
        s0, S0, G0 = ost.rts_smoother_meanfield(dts, ko, m0, P0, return_full=full)
        s1, S1, G1 = bn.spacetime.st_rts_smoother_meanfield(dts, kg, m0, P0, return_full=full)
        assert rel_err(np_(s1), s0) < TOL and rel_err(np_(S1), S0) < TOL and rel_err(np_(G1), G0) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['full', 'sparse'])
def test_gpu_meanfield_vi_iteration_vs_oracle(bn, case):
    """MarkovVariationalMeanFieldGP end to end"""
    Nt, Ns = 14, 16
    t, Y, R = st_data(Nt, Ns, seed=7, spatial_dims=2)
    z = st_data(1, 9, spatial_dims=2)[2][0] if case == 'sparse' else R[0]
    ko = oracle_kernel('Matern32', 1.1, 0.9, 1.2, z, spatial_dims=2)
    kg = gpu_kernel(bn, 'Matern32', 1.1, 0.9, 1.2, z, spatial_dims=2)
    mo = ost.SpatioTemporalMeanFieldMarkovGP(ko, sites.Gaussian(0.3), t, Y, R)
    mg = bn.models.MarkovVariationalMeanFieldGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.3), X=t, Y=Y, R=R)
    for lr in (1.0, 0.5):
        mo.inference(lr=lr)
        mg.inference(lr=lr)
        E0, E1 = mo.energy(), float(mg.energy())
        assert rel_err(np_(mg.posterior_mean), mo.post_mean) < TOL
        assert rel_err(np_(mg.posterior_variance), mo.post_cov) < TOL
        assert abs(E1 - E0) <= TOL * abs(E0), (E0, E1)


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.gpu
@pytest.mark.parametrize('N', [1, 2])
def test_gpu_dense_and_meanfield_on_one_and_two_steps(bn, N):
    M = 5
    rng = np.random.default_rng(N)
    z = np.linspace(0, 1, M)[:, None]
    ko, kg = oracle_kernel('Matern32', 1.3, 0.7, 0.5, z), gpu_kernel(bn, 'Matern32', 1.3, 0.7, 0.5, z)
    dt = np.concatenate([[0.0], 0.3 * np.ones(N - 1)])
    y = rng.standard_normal((N, M, 1))
    Rn = spd_batch(N, M, 3)
    dts = np.concatenate([dt[1:], [0.0]])
    e0, (m0, P0) = kalman.kalman_filter(dt, ko, y, Rn)
    e1, (m1, P1) = bn.ops.kalman_filter(dt, kg, y, Rn)
    assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    s0, S0, G0 = kalman.rauch_tung_striebel_smoother(dts, ko, m0, P0)
    s1, S1, G1 = bn.ops.rauch_tung_striebel_smoother(dts, kg, m0, P0)
    assert rel_err(np_(s1), s0) < TOL and rel_err(np_(S1), S0) < TOL and rel_err(np_(G1), G0) < TOL
    e0, (m0, P0) = ost.kalman_filter_meanfield(dt, ko, y, Rn)
    e1, (m1, P1) = bn.spacetime.st_kalman_filter_meanfield(dt, kg, y, Rn)
    assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0) < TOL and rel_err(np_(P1), P0) < TOL
    s0, S0, _ = ost.rts_smoother_meanfield(dts, ko, m0, P0)
    s1, S1, _ = bn.spacetime.st_rts_smoother_meanfield(dts, kg, m0, P0)
    assert rel_err(np_(s1), s0) < TOL and rel_err(np_(S1), S0) < TOL


@pytest.mark.gpu
def test_gpu_spatiotemporal_bernoulli_vi_vs_oracle(bn):
    """a non-conjugate likelihood on the spatio-temporal path: every (time, space) observation is a cubature site"""
    Nt, Ns = 10, 9
    t, Y, R = st_data(Nt, Ns, seed=11, spatial_dims=2)
    Y = (Y > np.median(Y)).astype(np.float64)
    ko = oracle_kernel('Matern32', 1.1, 0.9, 1.2, R[0], spatial_dims=2)
    kg = gpu_kernel(bn, 'Matern32', 1.1, 0.9, 1.2, R[0], spatial_dims=2)
    mo = ost.SpatioTemporalMarkovGP(ko, sites.Bernoulli(), t, Y, R)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Bernoulli(), X=t, Y=Y, R=R)
    for lr in (0.5, 0.5):
        mo.inference(lr=lr)
        mg.inference(lr=lr)
    E0, E1 = mo.energy(), float(mg.energy())
    assert rel_err(np_(mg.posterior_mean), mo.post_mean) < 1e-8
    assert rel_err(np_(mg.posterior_variance), mo.post_cov) < 1e-8
    assert abs(E1 - E0) <= 1e-8 * abs(E0), (E0, E1)


@pytest.mark.gpu
def test_gpu_spatiotemporal_rejects_what_it_does_not_cover(bn):
    t, Y, R = st_data(4, 4, seed=1, spatial_dims=2)
    kg = gpu_kernel(bn, 'Matern32', 1.0, 1.0, 1.0, R[0], spatial_dims=2)
    R2 = R.copy()
    R2[1] += 0.1  # spatial inputs that move in time
    with pytest.raises(NotImplementedError):
        bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.1), X=t, Y=Y, R=R2)
    with pytest.raises(NotImplementedError):
        bn.spacetime.SpatioTemporalKernel(bn.kernels.Matern32(), bn.kernels.Matern32(), z=R[0], sparse=False)
    L = bn._lib.lib()
    spec = bn.kernels.Matern72(1.0, 1.0).spec()
    assert L.bn_st_workspace_bytes(spec, 4, 4, 4) == 0 and b'family' in L.bn_last_error()


# ------------------------------------------------------------------------------------------------ prediction
def test_oracle_st_predict_at_training_inputs_is_the_posterior():
    t, Y, R = st_data(9, 9, seed=2, spatial_dims=2)
    k = oracle_kernel('Matern32', 1.1, 0.9, 1.2, R[0], spatial_dims=2)
    m = ost.SpatioTemporalMarkovGP(k, sites.Gaussian(0.3), t, Y, R)
    m.inference()
    mean, var = ost.markov_predict(m, m.t)
    mf, cf = m.conditional_posterior_to_data()
    assert rel_err(mean, mf[..., 0]) < 1e-6 and rel_err(var, np.diagonal(cf, axis1=1, axis2=2)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('fam', ['Matern12', 'Matern32', 'Matern52'])
@pytest.mark.parametrize('new_space', [False, True])
def test_gpu_spatiotemporal_predict_vs_oracle(bn, fam, new_space):
    Nt, Ns = 10, 9
    t, Y, R = st_data(Nt, Ns, seed=5, spatial_dims=2)
    ko = oracle_kernel(fam, 1.1, 0.9, 1.2, R[0], spatial_dims=2)
    kg = gpu_kernel(bn, fam, 1.1, 0.9, 1.2, R[0], spatial_dims=2)
    mo = ost.SpatioTemporalMarkovGP(ko, sites.Gaussian(0.3), t, Y, R)
    mg = bn.models.MarkovVariationalGP(kernel=kg, likelihood=bn.likelihoods.Gaussian(0.3), X=t, Y=Y, R=R)
    mo.inference(lr=0.8)
    mg.inference(lr=0.8)
    xs = np.concatenate([[t[0] - 2.0], 0.5 * (t[:-1] + t[1:])[::2], t[[0, 4, -1]], [t[-1] + 0.7, t[-1] + 30.0]])
    Rs = np.random.default_rng(0).uniform(-1, 1, (5, 2)) if new_space else None
    m0, v0 = ost.markov_predict(mo, xs, Rs)
    m1, v1 = mg.predict(xs, Rs)
    assert rel_err(np_(m1), m0) < 1e-8 and rel_err(np_(v1), v0) < 1e-8
