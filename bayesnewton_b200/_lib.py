"""ctypes binding of libbn_b200.so (the C ABI in include/bn_b200.h).

There is no CPU fallback and no alternative backend: if the CUDA library is missing or a call
fails, the error is raised.  torch is used for device memory, streams and (in distributed.py)
the NCCL plumbing only.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart before our library resolves it)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get('BN_B200_LIBNAME', 'libbn_b200.so'))  # tuning variants: build.py

BN_SEQUENTIAL, BN_SCAN = 0, 1
BN_MATERN12, BN_MATERN32, BN_MATERN52, BN_MATERN72 = 1, 2, 3, 4
FAMILY_DIM = {BN_MATERN12: 1, BN_MATERN32: 2, BN_MATERN52: 3, BN_MATERN72: 4}
BN_LIK_GAUSSIAN, BN_LIK_BERNOULLI_PROBIT, BN_LIK_BERNOULLI_LOGIT = 1, 2, 3
BN_LIK_HETEROSCEDASTIC_SOFTPLUS, BN_LIK_HETEROSCEDASTIC_EXP = 4, 5
BN_LIK_POISSON_EXP = 6
BN_LIK_STUDENTS_T, BN_LIK_GAMMA_EXP, BN_LIK_NEGBIN_EXP, BN_LIK_BETA_PROBIT = 7, 8, 9, 10
BN_METHOD_VI, BN_METHOD_EP, BN_METHOD_NEWTON, BN_METHOD_PL = 1, 2, 3, 4
BN_MAX_COMPONENTS = 4


class KernelSpec(C.Structure):
    _fields_ = [('family', C.c_int32), ('n_components', C.c_int32),
                ('variance', C.c_double * BN_MAX_COMPONENTS), ('lengthscale', C.c_double * BN_MAX_COMPONENTS)]


class SiteArgs(C.Structure):
    _fields_ = [('method', C.c_int32), ('likelihood', C.c_int32), ('lik_param', C.c_double), ('N', C.c_int64),
                ('D', C.c_int32), ('Q', C.c_int32), ('cub_x', C.c_void_p), ('cub_w', C.c_void_p), ('y', C.c_void_p),
                ('post_mean', C.c_void_p), ('post_cov', C.c_void_p), ('lr', C.c_double), ('power', C.c_double),
                ('ensure_psd', C.c_int32), ('pad_', C.c_int32), ('nat1', C.c_void_p), ('nat2', C.c_void_p),
                ('site_mean', C.c_void_p), ('site_cov', C.c_void_p), ('out_mean', C.c_void_p),
                ('out_jac', C.c_void_p), ('out_hess', C.c_void_p), ('diffs', C.c_void_p), ('lik_param2', C.c_double)]


class IterArgs(C.Structure):
    """bn_iter_args (include/bn_b200.h)"""
    _fields_ = [('N', C.c_int64), ('rank', C.c_int32), ('world', C.c_int32), ('dt_t', C.c_void_p), ('y_t', C.c_void_p),
                ('site_mean_t', C.c_void_p), ('site_cov_t', C.c_void_p), ('mask_t', C.c_void_p),
                ('post_mean_t', C.c_void_p), ('post_cov_t', C.c_void_p), ('method', C.c_int32), ('likelihood', C.c_int32),
                ('lik_param', C.c_double), ('Q', C.c_int32), ('ensure_psd', C.c_int32), ('cub_x_host', C.c_void_p),
                ('cub_w_host', C.c_void_p), ('lr', C.c_double), ('power', C.c_double), ('want_ell', C.c_int32),
                ('reserved_', C.c_int32), ('post_mean', C.c_void_p), ('post_cov', C.c_void_p)]


class BnError(RuntimeError):
    pass


_P, _I, _L, _Z, _D = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double
_KS, _SA, _IA = C.POINTER(KernelSpec), C.POINTER(SiteArgs), C.POINTER(IterArgs)

# every symbol include/bn_b200.h declares, with its argument types
SIGNATURES = {
    'bn_last_error': (C.c_char_p, []),
    'bn_version': (_I, []),
    'bn_timing_enable': (_I, [_I]),
    'bn_timing_report': (_I, [C.c_char_p, _Z]),
    'bn_measure_dfma_peak': (_I, [_P, _Z, C.POINTER(C.c_double)]),
    'bn_measure_dmma_peak': (_I, [_P, _Z, C.POINTER(C.c_double)]),
    'bn_state_dim': (_I, [_KS]),
    'bn_discretise': (_I, [_KS, _L, _P, _P, _P, _P]),
    'bn_workspace_bytes': (_Z, [_L, _I, _I]),
    'bn_kf_arrays': (_I, [_I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_rts_arrays': (_I, [_I, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_kalman_filter': (_I, [_KS, _I, _L, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_rts_smoother': (_I, [_KS, _I, _L, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_kf_carry_len': (_I, [_I]),
    'bn_rts_carry_len': (_I, [_I]),
    'bn_kf_shard_reduce': (_I, [_KS, _L, _I, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_kf_shard_apply': (_I, [_KS, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_rts_shard_reduce': (_I, [_KS, _L, _I, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_rts_shard_apply': (_I, [_KS, _L, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_update_posterior_workspace_bytes': (_Z, [_KS, _L]),
    'bn_update_posterior': (_I, [_KS, _L, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_update_posterior_grad': (_I, [_KS, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_up_shard_reduce': (_I, [_KS, _L, _I, _I, _P, _P, _P, _P, _I, _P, _Z, _P]),
    'bn_up_shard_filter': (_I, [_KS, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    'bn_up_shard_smooth': (_I, [_KS, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_site_update': (_I, [_SA, _P, _Z, _P]),
    'bn_likelihood_stats': (_I, [_SA, _P, _P, _P, _P, _Z, _P]),
    'bn_likelihood_param_grad': (_I, [_SA, _P, _P, _Z, _P]),
    'bn_expected_density': (_I, [_SA, _P, _P, _P, _Z, _P]),
    'bn_energy_terms': (_I, [_SA, _P, _P, _P, _Z, _P]),
    'bn_gaussian_expected_log_lik': (_I, [_L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_ep_pseudo_density': (_I, [_L, _I, _D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_temporal_conditional': (_I, [_KS, _L, _P, _L, _P, _P, _P, _P, _I, _P, _P, _P]),
    'bn_likelihood_predict': (_I, [_I, _D, _L, _P, _P, _I, _P, _P, _P, _P, _P]),
    'bn_likelihood_predict2': (_I, [_I, _D, _D, _L, _P, _P, _I, _P, _P, _P, _P, _P]),
    'bn_pairs_discretise': (_I, [_KS, _L, _P, _P, _P, _P]),
    'bn_build_joint': (_I, [_KS, _L, _P, _P, _P, _P, _P, _P]),
    'bn_sparse_workspace_bytes': (_Z, [_L]),
    'bn_sparse_site_update': (_I, [_KS, _I, _D, _L, _L, _P, _P, _P, _P, _P, _P, _I, _P, _P, _D, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_sparse_expected_density': (_I, [_KS, _I, _D, _L, _L, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_st_kalman_filter_meanfield': (_I, [_KS, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_st_rts_smoother_meanfield': (_I, [_KS, _I, _L, _P, _P, _P, _I, _P, _P, _P, _P]),
    'bn_carry_exchange_bytes': (_Z, [_I]),
    'bn_carry_exchange': (_I, [_P, _I, _I, _P, _I, C.c_uint64, _P, _P]),
    'bn_st_predict_workspace_bytes': (_Z, [_KS, _I, _L]),
    'bn_st_predict_state': (_I, [_KS, _I, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_st_workspace_bytes': (_Z, [_KS, _I, _L, _I]),
    'bn_st_kalman_filter': (_I, [_KS, _I, _L, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_st_rts_smoother': (_I, [_KS, _I, _L, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_st_profile': (_I, [C.POINTER(C.c_int64), _I]),
    'bn_spd_inverse_batched': (_I, [_L, _I, _P, _P, _D, _P, _P, _P, _P, _Z, _P]),
    'bn_st_pseudo_lik': (_I, [_L, _I, _I, _P, _P, _P, _D, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_st_posterior_to_data': (_I, [_L, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'bn_st_gaussian_expected_log_lik': (_I, [_L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'bn_iter_chunk_len': (_I, [_KS, _L]),
    'bn_iter_tiled_len': (_L, [_KS, _L]),
    'bn_iter_workspace_bytes': (_Z, [_KS, _L]),
    'bn_iter_to_tiled': (_I, [_KS, _L, _P, _P, _D, _P]),
    'bn_iter_from_tiled': (_I, [_KS, _L, _P, _P, _P]),
    'bn_iter_to_tiled_u8': (_I, [_KS, _L, _P, _P, _P]),
    'bn_iter_pass': (_I, [_KS, _IA, _I, _P, _P, _P, _Z, _P]),
    'bn_iter_shard_reduce': (_I, [_KS, _IA, _P, _P, _Z, _P]),
    'bn_iter_shard_filter': (_I, [_KS, _IA, _P, _P, _P, _P, _Z, _P]),
    'bn_iter_shard_smooth': (_I, [_KS, _IA, _I, _P, _P, _P, _Z, _P]),
    'bn_iter_workspace_bytes_f32': (_Z, [_KS, _L]),
    'bn_iter_to_tiled_f32': (_I, [_KS, _L, _P, _P, C.c_float, _P]),
    'bn_iter_from_tiled_f32': (_I, [_KS, _L, _P, _P, _P]),
    'bn_iter_pass_f32': (_I, [_KS, _IA, _I, _P, _P, _P, _Z, _P]),
    'bn_iter_shard_reduce_f32': (_I, [_KS, _IA, _P, _P, _Z, _P]),
    'bn_iter_shard_filter_f32': (_I, [_KS, _IA, _P, _P, _P, _P, _Z, _P]),
    'bn_iter_shard_smooth_f32': (_I, [_KS, _IA, _I, _P, _P, _P, _Z, _P]),
    'bn_ih_workspace_bytes': (_Z, [_I, _L]),
    'bn_ih_filter': (_I, [_I, _I, _L, _P, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'bn_ih_smoother': (_I, [_I, _I, _L, _P, _P, _P, _I, _P, _P, _Z, _P]),
}

_lib = None


def lib():
    """the loaded library; raises if it has not been built (python -m bayesnewton_b200.build)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BnError('%s is missing: build it with `python -m bayesnewton_b200.build` '
                          '(there is no CPU fallback)' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().bn_last_error().decode()
        raise BnError('libbn_b200 call failed (code %d): %s' % (rc, msg))


def kernel_spec(family, variances, lengthscales):
    s = KernelSpec()
    s.family, s.n_components = family, len(variances)
    for i, (v, l) in enumerate(zip(variances, lengthscales)):
        s.variance[i], s.lengthscale[i] = float(v), float(l)
    return s
