"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'bn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
    syms = header_symbols()
    for s in ['bn_kalman_filter', 'bn_rts_smoother', 'bn_kf_arrays', 'bn_rts_arrays', 'bn_site_update',
              'bn_discretise', 'bn_kf_shard_reduce', 'bn_rts_shard_apply']:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from bayesnewton_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(handle, s), 'libbn_b200.so does not export %s' % s
    assert set(_lib.SIGNATURES) == set(header_symbols())


def test_host_only_queries():
    from bayesnewton_b200 import _lib
    L = _lib.lib()
    assert L.bn_version() >= 100
    assert L.bn_state_dim(_lib.kernel_spec(_lib.BN_MATERN52, [1.0], [1.0])) == 3
    assert L.bn_state_dim(_lib.kernel_spec(_lib.BN_MATERN32, [1.0, 2.0], [1.0, 1.0])) == 4
    assert L.bn_kf_carry_len(3) == 33 and L.bn_rts_carry_len(3) == 21
    assert L.bn_workspace_bytes(10 ** 7, 3, 1) > 0


def test_bad_arguments_are_reported_not_executed():
    from bayesnewton_b200 import _lib
    L = _lib.lib()
    spec = _lib.kernel_spec(_lib.BN_MATERN52, [1.0], [1.0])
    rc = L.bn_kalman_filter(spec, 7, 10, None, None, None, None, 0, None, None, None, None, 0, None)
    assert rc < 0 and b'form' in L.bn_last_error()
    rc = L.bn_kalman_filter(spec, 1, 10, None, None, None, None, 0, None, None, None, None, 0, None)
    assert rc < 0 and b'null' in L.bn_last_error()
    with pytest.raises(_lib.BnError):
        _lib.check(rc)


def test_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from bayesnewton_b200 import _lib, _util
    with pytest.raises(_lib.BnError):
        _util.device()


def test_new_entries_report_bad_arguments():
    """the dense spatio-temporal, sparse, prediction and exchange entries validate before they launch (no GPU needed)"""
    from bayesnewton_b200 import _lib
    L = _lib.lib()
    m32 = _lib.kernel_spec(_lib.BN_MATERN32, [1.0], [1.0])
    m72 = _lib.kernel_spec(_lib.BN_MATERN72, [1.0], [1.0])
    two = _lib.kernel_spec(_lib.BN_MATERN32, [1.0, 1.0], [1.0, 1.0])
    assert L.bn_st_workspace_bytes(m32, 256, 10_000, 256) > 0
    assert L.bn_st_workspace_bytes(m72, 256, 10, 256) == 0 and b'family' in L.bn_last_error()
    assert L.bn_st_workspace_bytes(two, 8, 10, 8) == 0 and b'one component' in L.bn_last_error()
    rc = L.bn_st_kalman_filter(m32, 8, 5, None, None, None, None, 0, None, None, None, None, 0, None)
    assert rc < 0 and b'null' in L.bn_last_error()
    rc = L.bn_st_rts_smoother(m32, 0, 5, None, None, None, 0, None, None, None, None, 0, None)
    assert rc < 0 and b'out of range' in L.bn_last_error()
    rc = L.bn_spd_inverse_batched(3, 0, None, None, 0.0, None, None, None, None, 0, None)
    assert rc < 0
    rc = L.bn_temporal_conditional(two, 4, None, 3, None, None, None, None, 0, None, None, None)
    assert rc < 0 and b'null' in L.bn_last_error()
    rc = L.bn_likelihood_predict(_lib.BN_LIK_HETEROSCEDASTIC_SOFTPLUS, 0.0, 4, None, None, 0, None, None, None, None, None)
    assert rc < 0 and b'single-latent' in L.bn_last_error()
    rc = L.bn_pairs_discretise(m72, 4, None, None, None, None)
    assert rc < 0 and b'sparse Markov' in L.bn_last_error()
    assert L.bn_sparse_workspace_bytes(100) >= 3 * 101 * 8
    assert L.bn_carry_exchange_bytes(8) == 2 * 8 * 256 * 8 + 2 * 8 * 8
    rc = L.bn_carry_exchange(None, 8, 0, None, 33, 1, None, None)
    assert rc < 0 and b'null' in L.bn_last_error()
    assert L.bn_st_profile(None, 4) < 0
