"""XLA custom-call side of the boundary (include/bn_b200_xla.h).

The reference pins jax==0.4.14 (requirements.txt:1-2): no jax.ffi, so the GPU entry is the legacy
``void(stream, void** buffers, const char* opaque, size_t opaque_len)`` custom-call.  This module packs the
opaque descriptors, exposes the target table, and -- only when a JAX install is importable, which is NOT the
case in this image -- registers the targets with xla_client.  ``call`` invokes a target exactly as XLA's
runtime would (operand pointers then result pointers, one opaque byte string); the GPU tests use it to pin
the wrappers against the direct C-ABI calls.
"""
import ctypes as C

from . import _lib

MAX_CUB = 1200


class MarkovDesc(C.Structure):
    _fields_ = [('spec', _lib.KernelSpec), ('N', C.c_int64), ('workspace_bytes', C.c_uint64), ('form', C.c_int32),
                ('has_mask', C.c_int32), ('want_grad', C.c_int32), ('return_predict', C.c_int32),
                ('return_full', C.c_int32), ('pad_', C.c_int32)]


class SiteDesc(C.Structure):
    _fields_ = [('method', C.c_int32), ('likelihood', C.c_int32), ('lik_param', C.c_double), ('N', C.c_int64),
                ('D', C.c_int32), ('Q', C.c_int32), ('lr', C.c_double), ('power', C.c_double),
                ('ensure_psd', C.c_int32), ('has_mask', C.c_int32), ('workspace_bytes', C.c_uint64),
                ('cub_w', C.c_double * 400), ('cub_x', C.c_double * MAX_CUB), ('lik_param2', C.c_double)]


def markov_desc(spec, N, workspace_bytes, form=_lib.BN_SCAN, has_mask=False, want_grad=False, return_predict=False,
                return_full=False):
    d = MarkovDesc()
    d.spec, d.N, d.workspace_bytes, d.form = spec, int(N), int(workspace_bytes), int(form)
    d.has_mask, d.want_grad = int(has_mask), int(want_grad)
    d.return_predict, d.return_full = int(return_predict), int(return_full)
    return bytes(d)


def site_desc(site_args, workspace_bytes, cub_x=None, cub_w=None, has_mask=False):
    """from a filled bn_site_args (likelihoods.Likelihood.site_args) + the host cubature rule"""
    a = site_args
    d = SiteDesc()
    d.method, d.likelihood, d.lik_param, d.N, d.D, d.Q = a.method, a.likelihood, a.lik_param, a.N, a.D, a.Q
    d.lik_param2 = a.lik_param2
    d.lr, d.power, d.ensure_psd, d.has_mask, d.workspace_bytes = a.lr, a.power, a.ensure_psd, int(has_mask), int(workspace_bytes)
    if a.Q:
        if a.Q > 400 or a.D * a.Q > MAX_CUB:
            raise ValueError('cubature rule too large for the XLA descriptor')
        w = (C.c_double * a.Q).from_address(a.cub_w) if cub_w is None else cub_w
        x = (C.c_double * (a.D * a.Q)).from_address(a.cub_x) if cub_x is None else cub_x
        for i in range(a.Q):
            d.cub_w[i] = w[i]
        for i in range(a.D * a.Q):
            d.cub_x[i] = x[i]
    return bytes(d)


def targets():
    """{name: address} of every custom-call target the library exports"""
    L = C.CDLL(_lib.LIB_PATH)
    L.bn_xla_targets.restype = C.c_int
    L.bn_xla_targets.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.c_int]
    n = L.bn_xla_targets(None, None, 0)
    names, ptrs = (C.c_char_p * n)(), (C.c_void_p * n)()
    L.bn_xla_targets(names, ptrs, n)
    return {names[i].decode(): ptrs[i] for i in range(n)}


def error_count():
    L = C.CDLL(_lib.LIB_PATH)
    L.bn_xla_error_count.restype = C.c_long
    return L.bn_xla_error_count()


def call(name, stream, buffers, opaque):
    """invoke a target the way XLA's GPU runtime does: f(stream, void** buffers, opaque, len)"""
    fn = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t)(targets()[name])
    arr = (C.c_void_p * len(buffers))(*[int(b) if b is not None else None for b in buffers])
    fn(stream, arr, opaque, len(opaque))


def register_with_jax():
    """register every target with XLA (platform CUDA).  Needs jax/jaxlib; raises ImportError otherwise."""
    from jax.lib import xla_client  # noqa: not installed in this image -- exercised only where JAX exists
    C.pythonapi.PyCapsule_New.restype = C.py_object
    C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    for name, addr in targets().items():
        capsule = C.pythonapi.PyCapsule_New(addr, b'xla._CUSTOM_CALL_TARGET', None)
        xla_client.register_custom_call_target(name.encode(), capsule, platform='CUDA')
    return sorted(targets())
