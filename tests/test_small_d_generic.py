"""Warp-cooperative filter / smoother for state dimensions up to 16 (csrc/gd_impl.cuh, gd.cu; bn_kf_arrays / bn_rts_arrays
beyond the register-resident instantiations): Independent stacks Matern-7/2 x 2 (d = 8), Matern-5/2 x 3 (d = 9), a d = 16 stack,
both forms, masks, return_predict / return_full / gains.  CPU: the chunk bodies on the host (tests/hostemu) against the
oracle; GPU: through the C ABI against the oracle.  Tolerance 1e-9 normwise."""
import numpy as np
import pytest

import _emu
from _data import rel_err
from oracle import kalman, ssm

TOL = 1e-9
STACKS = {'m72x2_d8': [ssm.Matern72(0.9, 1.3), ssm.Matern72(1.2, 0.7)],
          'm52x3_d9': [ssm.Matern52(1.0, 0.8), ssm.Matern52(0.6, 1.9), ssm.Matern52(1.4, 0.5)],
          'm72x4_d16': [ssm.Matern72(0.9, 1.3), ssm.Matern72(1.2, 0.7), ssm.Matern72(0.5, 2.0), ssm.Matern72(1.1, 1.1)],
          'm32x5_d10': [ssm.Matern32(1.0 + 0.1 * i, 0.5 + 0.3 * i) for i in range(5)]}


def problem(name, N, seed=0, missing=True):
    """As, Qs, H, ys, Rs (full D x D, correlated), m0, P0, masks for an Independent stack with one latent per component"""
    k = ssm.Independent(STACKS[name])
    rng = np.random.default_rng(seed)
    dt = np.concatenate([[0.0], 0.05 + 0.3 * rng.random(N - 1)])
    As = np.stack([k.state_transition(h) for h in dt])
    Pinf = k.stationary_covariance()
    Qs = Pinf[None] - As @ Pinf[None] @ np.transpose(As, (0, 2, 1))
    H = k.measurement_model()
    D = H.shape[0]
    ys = rng.standard_normal((N, D, 1))
    B = 0.3 * rng.standard_normal((N, D, D))
    Rs = B @ np.transpose(B, (0, 2, 1)) + (0.2 + rng.random((N, 1, 1))) * np.eye(D)[None]
    masks = np.zeros((N, D, 1), dtype=bool)
    if missing and N > 5:
        masks[::7, 0] = True
        masks[3::11] = True
    m0 = np.zeros((H.shape[1], 1))
    return As, Qs, H, ys, Rs, m0, Pinf, masks


def smoother_inputs(As, Qs):
    """the transitions OUT of each step (basemodels.py:700): shifted by one, identity / zero at the end"""
    d = As.shape[1]
    return np.concatenate([As[1:], np.eye(d)[None]]), np.concatenate([Qs[1:], np.zeros((1, d, d))])


@pytest.fixture(scope='module')
def emu():
    return _emu.load()


@pytest.mark.parametrize('name', sorted(STACKS))
@pytest.mark.parametrize('form,N,L', [(0, 1, 16), (0, 57, 16), (1, 1, 16), (1, 57, 16), (1, 203, 8), (1, 64, 64), (1, 1200, 8), (1, 8300, 8)])
def test_emu_generic_filter_and_smoother(emu, name, form, N, L):
    if N > 5000 and name != 'm72x2_d8':
        pytest.skip('the three-level scan (more than 1024 chunks) is exercised on one stack')
    As, Qs, H, ys, Rs, m0, P0, masks = problem(name, N, seed=N)
    e0, m0s, P0s = kalman.sequential_kf(As, Qs, H, ys, Rs, m0, P0, masks)
    e1, m1, P1 = _emu.gd_kf(emu, form, As, Qs, H, ys, Rs, m0, P0, masks, L=L)
    assert abs(e1 - e0) <= TOL * abs(e0) and rel_err(m1, m0s) < TOL and rel_err(P1, P0s) < TOL
    _, mp0, Pp0 = kalman.sequential_kf(As, Qs, H, ys, Rs, m0, P0, masks, return_predict=True)
    _, mp1, Pp1 = _emu.gd_kf(emu, form, As, Qs, H, ys, Rs, m0, P0, masks, L=L, return_predict=True, want_ell=False)
    assert rel_err(mp1, mp0) < TOL and rel_err(Pp1, Pp0) < TOL
    A2, Q2 = smoother_inputs(As, Qs)
    for full in (False, True):
        s0, S0, G0 = kalman.sequential_rts(m0s, P0s, A2, Q2, H, full)
        s1, S1, G1 = _emu.gd_rts(emu, form, m0s, P0s, A2, Q2, H, L=L, return_full=full)
        assert rel_err(s1, s0) < TOL and rel_err(S1, S0) < TOL and rel_err(G1, G0) < TOL


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def np_(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(STACKS))
@pytest.mark.parametrize('parallel,N', [(False, 1), (False, 300), (True, 1), (True, 300), (True, 20_011)])
def test_gpu_generic_filter_and_smoother(bn, name, parallel, N):
    import torch
    from bayesnewton_b200 import ops
    As, Qs, H, ys, Rs, m0, P0, masks = problem(name, N, seed=N + 1)
    e0, m0s, P0s = kalman.sequential_kf(As, Qs, H, ys, Rs, m0, P0, masks)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    kf = ops._parallel_kf if parallel else ops._sequential_kf
    e1, m1, P1 = kf(t(As), t(Qs), t(H), t(ys), t(Rs), t(m0), t(P0), t(masks.astype(np.uint8)))
    assert abs(float(e1) - e0) <= TOL * abs(e0) and rel_err(np_(m1), m0s) < TOL and rel_err(np_(P1), P0s) < TOL
    A2, Q2 = smoother_inputs(As, Qs)
    rts = ops._parallel_rts if parallel else ops._sequential_rts
    for full in (False, True):
        s0, S0, G0 = kalman.sequential_rts(m0s, P0s, A2, Q2, H, full)
        s1, S1, G1 = rts(m1, P1, t(A2), t(Q2), t(H), full)
        assert rel_err(np_(s1), s0) < TOL and rel_err(np_(S1), S0) < TOL and rel_err(np_(G1), G0) < TOL


@pytest.mark.gpu
def test_gpu_generic_rejects_dimensions_above_16(bn):
    import torch
    from bayesnewton_b200 import ops, _lib
    d, N = 17, 4
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    As = np.tile(np.eye(d)[None], (N, 1, 1))
    with pytest.raises(_lib.BnError):
        ops._sequential_kf(t(As), t(0.1 * As), t(np.eye(1, d)), t(np.zeros((N, 1, 1))), t(np.ones((N, 1, 1))), t(np.zeros((d, 1))),
                           t(np.eye(d)), None)
