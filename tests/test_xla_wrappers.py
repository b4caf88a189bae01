"""Legacy XLA custom-call wrappers (include/bn_b200_xla.h): exported, descriptor layouts agree between C and
ctypes (CPU), and -- on a GPU -- every wrapper invoked the way XLA's runtime does reproduces the direct C-ABI call
bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def xla_header_symbols():
    src = open(os.path.join(ROOT, 'include', 'bn_b200_xla.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bn_xla_[a-z0-9_]+)\s*\(', src)))


def test_exports_and_target_table():
    from bayesnewton_b200 import _lib, xla
    handle = C.CDLL(_lib.LIB_PATH)
    syms = xla_header_symbols()
    assert len(syms) == 8
    for s in syms:
        assert hasattr(handle, s), s
    t = xla.targets()
    assert sorted(t) == sorted(s for s in syms if s not in ('bn_xla_targets', 'bn_xla_error_count'))
    assert all(t.values())


def test_descriptor_layouts_match_and_bad_opaque_is_refused():
    """a wrong-sized opaque is refused before any buffer is touched; the message carries the C sizeof"""
    from bayesnewton_b200 import _lib, xla
    L = _lib.lib()
    for name, ctype in (('bn_xla_update_posterior', xla.MarkovDesc), ('bn_xla_kalman_filter', xla.MarkovDesc),
                        ('bn_xla_rts_smoother', xla.MarkovDesc), ('bn_xla_site_update', xla.SiteDesc),
                        ('bn_xla_expected_density', xla.SiteDesc), ('bn_xla_gaussian_expected_log_lik', xla.SiteDesc)):
        before = xla.error_count()
        xla.call(name, None, [None] * 12, b'x')
        assert xla.error_count() == before + 1
        msg = L.bn_last_error().decode()
        assert name in msg and ('expected %d' % C.sizeof(ctype)) in msg, msg


@pytest.fixture(scope='module')
def bn():
    import torch
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


@pytest.mark.gpu
@pytest.mark.parametrize('want_grad', [False, True])
def test_gpu_xla_update_posterior(bn, want_grad):
    import torch
    from _data import filter_problem
    from bayesnewton_b200 import _lib, xla
    from bayesnewton_b200._util import as_dev, as_mask, stream_ptr
    N = 5003
    dt, y, R, mask = filter_problem(N, D=1, seed=1)
    k = bn.kernels.Matern52(1.3, 0.9)
    ref = bn.ops.update_posterior(dt, k, y, R, None if want_grad else mask, want_ell=True, want_grad=want_grad)
    spec = k.spec()
    nb = _lib.lib().bn_update_posterior_workspace_bytes(spec, N)
    dev = ref[1].device
    ins = [as_dev(dt), as_dev(y), as_dev(R)] + ([] if want_grad else [as_mask(mask)])
    outs = [torch.zeros((), dtype=torch.float64, device=dev), torch.empty_like(ref[1]), torch.empty_like(ref[2])]
    if want_grad:
        outs += [torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)]
    outs.append(torch.empty(nb, dtype=torch.uint8, device=dev))
    e0 = xla.error_count()
    xla.call('bn_xla_update_posterior', stream_ptr(), [t.data_ptr() for t in ins + outs],
             xla.markov_desc(spec, N, nb, has_mask=not want_grad, want_grad=want_grad))
    torch.cuda.synchronize()
    assert xla.error_count() == e0, _lib.lib().bn_last_error()
    assert float(outs[0]) == float(ref[0]) and bool((outs[1] == ref[1]).all()) and bool((outs[2] == ref[2]).all())
    if want_grad:
        assert float(outs[3][0]) == float(ref[3][0, 0]) and float(outs[4][0]) == float(ref[3][1, 0])


@pytest.mark.gpu
def test_gpu_xla_filter_smoother_sites(bn):
    import torch
    from _data import classification_data, filter_problem
    from bayesnewton_b200 import _lib, xla
    from bayesnewton_b200._util import as_dev, as_mask, stream_ptr, workspace
    L = _lib.lib()
    N = 3001
    dt, y, R, mask = filter_problem(N, D=1, seed=2)
    k = bn.kernels.Matern32(1.1, 0.6)
    spec = k.spec()
    dev = as_dev(dt).device
    ell, (fm, fP) = bn.ops.kalman_filter(dt, k, y, R, mask, parallel=True)
    nb = L.bn_workspace_bytes(N, 2, 1)
    outs = [torch.zeros((), dtype=torch.float64, device=dev), torch.empty_like(fm), torch.empty_like(fP),
            torch.empty(nb, dtype=torch.uint8, device=dev)]
    ins = [as_dev(dt), as_dev(y), as_dev(R), as_mask(mask)]
    e0 = xla.error_count()
    xla.call('bn_xla_kalman_filter', stream_ptr(), [t.data_ptr() for t in ins + outs],
             xla.markov_desc(spec, N, nb, has_mask=True))
    torch.cuda.synchronize()
    assert float(outs[0]) == float(ell) and bool((outs[1] == fm).all()) and bool((outs[2] == fP).all())
    dts = np.concatenate([dt[1:], [0.0]])
    sm, sP, G = bn.ops.rauch_tung_striebel_smoother(dts, k, fm, fP, parallel=True)
    outs = [torch.empty_like(sm), torch.empty_like(sP), torch.empty_like(G), torch.empty(nb, dtype=torch.uint8, device=dev)]
    xla.call('bn_xla_rts_smoother', stream_ptr(), [as_dev(dts).data_ptr(), fm.data_ptr(), fP.data_ptr()] +
             [t.data_ptr() for t in outs], xla.markov_desc(spec, N, nb))
    torch.cuda.synchronize()
    assert bool((outs[0] == sm).all()) and bool((outs[1] == sP).all()) and bool((outs[2] == G).all())
    # sites: VI probit update + expected density + expected pseudo log-lik
    x, yc = classification_data(N, seed=4)
    lik = bn.likelihoods.Bernoulli()
    pm, pc = sm.contiguous(), sP.contiguous()
    nat1 = torch.zeros((N, 1, 1), dtype=torch.float64, device=dev)
    nat2 = 1e-2 * torch.ones((N, 1, 1), dtype=torch.float64, device=dev)
    a, keep = lik.site_args(_lib.BN_METHOD_VI, yc, pm, pc, None, 1.0)
    a.lr, a.ensure_psd = 0.6, 1
    n1, n2 = nat1.clone(), nat2.clone()
    a.nat1, a.nat2 = n1.data_ptr(), n2.data_ptr()
    smean, scov = torch.empty_like(nat1), torch.empty_like(nat2)
    diffs = torch.zeros(2, dtype=torch.float64, device=dev)
    a.site_mean, a.site_cov, a.diffs = smean.data_ptr(), scov.data_ptr(), diffs.data_ptr()
    ws, nbs = workspace(N, 2, 1)
    _lib.check(L.bn_site_update(a, ws.data_ptr(), nbs, stream_ptr()))
    o = [torch.empty_like(nat1), torch.empty_like(nat2), torch.empty_like(nat1), torch.empty_like(nat2),
         torch.zeros(2, dtype=torch.float64, device=dev), torch.empty(nbs, dtype=torch.uint8, device=dev)]
    yd = as_dev(yc).reshape(-1)
    xla.call('bn_xla_site_update', stream_ptr(), [yd.data_ptr(), pm.data_ptr(), pc.data_ptr(), nat1.data_ptr(),
                                                   nat2.data_ptr()] + [t.data_ptr() for t in o],
             xla.site_desc(a, nbs))
    torch.cuda.synchronize()
    for got, want in zip(o[:5], (n1, n2, smean, scov, diffs)):
        assert bool((got == want).all())
    assert float(nat2.min()) == 1e-2   # the operands are untouched (XLA buffers are immutable)
    want = torch.zeros((), dtype=torch.float64, device=dev)
    _lib.check(L.bn_expected_density(a, None, want.data_ptr(), ws.data_ptr(), nbs, stream_ptr()))
    got = torch.zeros((), dtype=torch.float64, device=dev)
    xla.call('bn_xla_expected_density', stream_ptr(), [yd.data_ptr(), pm.data_ptr(), pc.data_ptr(), n1.data_ptr(),
                                                        n2.data_ptr(), got.data_ptr(), o[5].data_ptr()],
             xla.site_desc(a, nbs))
    want2 = torch.zeros((), dtype=torch.float64, device=dev)
    _lib.check(L.bn_gaussian_expected_log_lik(N, 1, smean.data_ptr(), pm.data_ptr(), pc.data_ptr(), scov.data_ptr(), None,
                                              None, want2.data_ptr(), ws.data_ptr(), nbs, stream_ptr()))
    got2 = torch.zeros((), dtype=torch.float64, device=dev)
    xla.call('bn_xla_gaussian_expected_log_lik', stream_ptr(), [smean.data_ptr(), pm.data_ptr(), pc.data_ptr(),
                                                                 scov.data_ptr(), got2.data_ptr(), o[5].data_ptr()],
             xla.site_desc(a, nbs))
    torch.cuda.synchronize()
    assert float(got) == float(want) and float(got2) == float(want2)
    assert xla.error_count() == e0, L.bn_last_error()


def test_jax_glue_is_import_safe_without_jax():
    """the JAX side of the boundary (primitives, lowering, custom_vjp) must import cleanly where jax is absent"""
    from bayesnewton_b200 import jax_glue
    assert jax_glue.available() in (True, False)
    if not jax_glue.available():
        import pytest as _pt
        with _pt.raises(ImportError):
            jax_glue.register()
