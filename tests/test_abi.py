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
