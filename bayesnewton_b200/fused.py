"""Resident state of the fused inference iteration (csrc/iter_impl.cuh, C ABI ``bn_iter_*``).

One iteration of a temporal Markov GP -- ``model.inference()`` (inference.py:65-90: update_posterior, site update,
update_posterior) followed by ``model.energy()`` (inference.py:197-222) -- is two fused passes over the time axis:

    pass 1  bn_iter_pass(BN_ITER_SITES)   filter, smoother, and the site update in the smoother's epilogue
    pass 2  bn_iter_pass(BN_ITER_ENERGY)  filter (+ log-likelihood), smoother, and the two energy sums in its epilogue

Between them nothing but the sites (16 B per step) touches HBM in a layout other than the consuming thread's own:
dt, Y, the sites and the posterior marginals are kept "tiled" (see include/bn_b200.h), and are converted from / to
the reference's [N, 1, 1] arrays only at the boundary (construction, and when a caller reads them).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._util import as_dev, ptr, stream_ptr
from .cubature import host_table

PLAIN, SITES, ENERGY = 0, 1, 2

SUPPORTED_LIKS = (_lib.BN_LIK_GAUSSIAN, _lib.BN_LIK_BERNOULLI_PROBIT, _lib.BN_LIK_BERNOULLI_LOGIT, _lib.BN_LIK_POISSON_EXP)
SUPPORTED_METHODS = (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON)


def fused_ep():
    """the EP epilogues of the fused passes are built and exact (same energy to 1e-15), but measured 6 % SLOWER than the
    stage-level kernels at N = 1e7 (3.79 vs 3.58 ms per iteration, profiles/r7c_ep_fused_vs_stage_level_n1e7.json: the
    exp per cubature point runs at the smoother's 16 warps per SM instead of the site kernel's 32), so EP takes the
    stage-level path unless BN_B200_FUSED_EP=1"""
    import os
    return os.environ.get('BN_B200_FUSED_EP', '0') == '1'


def supported(spec, likelihood, method):
    ok = method in SUPPORTED_METHODS or (method == _lib.BN_METHOD_EP and fused_ep())
    return spec is not None and spec.n_components == 1 and getattr(likelihood, 'lik_id', None) in SUPPORTED_LIKS and ok


class FusedShard:
    """the tiled arrays of one time shard (rank of world; a single GPU is shard 0 of 1) and the passes over them"""

    def __init__(self, kernel, dt, Y, mask=None, rank=0, world=1, dtype=torch.float64):
        """dtype = torch.float32 selects the fp32 build of the same kernels (bn_iter_*_f32: storage and arithmetic in
        fp32, sums in fp64; parity bar 1e-4 against the fp64 results)"""
        if dtype not in (torch.float64, torch.float32):
            raise ValueError('the fused iteration is built for float64 and float32')
        self.kernel = kernel
        self.dtype = dtype
        self._sfx = '_f32' if dtype == torch.float32 else ''
        spec = kernel.spec()
        L = _lib.lib()
        self.rank, self.world = int(rank), int(world)
        dt = as_dev(dt).reshape(-1)
        self.N = int(dt.shape[0])
        self.dev = dt.device
        self.chunk_len = int(L.bn_iter_chunk_len(spec, self.N))
        self.tlen = int(L.bn_iter_tiled_len(spec, self.N))
        if self.chunk_len <= 0 or self.tlen <= 0:
            raise _lib.BnError('fused iteration unavailable: ' + L.bn_last_error().decode())
        nb = int(self._fn('bn_iter_workspace_bytes')(spec, self.N))
        self.ws = torch.empty(nb, dtype=torch.uint8, device=self.dev)
        self.dt_t = self.to_tiled(dt, 0.0)
        self.set_data(Y, mask)
        self.sy_t = torch.zeros(self.tlen, dtype=self.dtype, device=self.dev)
        self.sR_t = torch.ones(self.tlen, dtype=self.dtype, device=self.dev)
        self.pm_t = torch.zeros(self.tlen, dtype=self.dtype, device=self.dev)
        self.pc_t = torch.ones(self.tlen, dtype=self.dtype, device=self.dev)
        self.d = int(L.bn_state_dim(spec))
        self.kf_len, self.rts_len = int(L.bn_kf_carry_len(self.d)), int(L.bn_rts_carry_len(self.d))

    def _fn(self, name):
        return getattr(_lib.lib(), name + self._sfx)

    # ---- layout conversion at the boundary
    def to_tiled(self, x, fill=0.0, out=None):
        x = as_dev(x).reshape(-1).to(self.dtype)
        if x.shape[0] != self.N:
            raise ValueError('expected %d steps, got %d' % (self.N, x.shape[0]))
        out = torch.empty(self.tlen, dtype=self.dtype, device=self.dev) if out is None else out
        _lib.check(self._fn('bn_iter_to_tiled')(self.kernel.spec(), self.N, ptr(x), ptr(out), float(fill), stream_ptr()))
        return out

    def from_tiled(self, x_t, shape=None, out=None):
        out = torch.empty(self.N, dtype=self.dtype, device=self.dev) if out is None else out
        _lib.check(self._fn('bn_iter_from_tiled')(self.kernel.spec(), self.N, ptr(x_t), ptr(out), stream_ptr()))
        return out if shape is None else out.reshape(shape)

    def set_dt(self, dt):
        self.to_tiled(dt, 0.0, out=self.dt_t)

    def set_data(self, Y, mask=None, scan_nan=True):
        """observations (one per step) and the mask of the missing ones (None: taken from the NaNs of Y)"""
        Y = as_dev(Y).reshape(-1)
        if getattr(self, 'y_t', None) is None:
            self.y_t = torch.empty(self.tlen, dtype=self.dtype, device=self.dev)
        self.to_tiled(Y, 0.0, out=self.y_t)
        if mask is None and scan_nan:
            nan = torch.isnan(Y)
            mask = nan.to(torch.uint8) if bool(nan.any()) else None
        if mask is None:
            self.mask_t = None
        else:
            mk = mask.to(device=self.dev, dtype=torch.uint8).reshape(-1).contiguous()
            self.mask_t = torch.zeros(self.tlen, dtype=torch.uint8, device=self.dev)
            _lib.check(_lib.lib().bn_iter_to_tiled_u8(self.kernel.spec(), self.N, ptr(mk), ptr(self.mask_t), stream_ptr()))

    def load_sites(self, mean, cov):
        self.to_tiled(mean, 0.0, out=self.sy_t)
        self.to_tiled(cov, 1.0, out=self.sR_t)

    def sites(self):
        """(pseudo_y, pseudo_var) as the reference's [N, 1, 1] arrays"""
        return self.from_tiled(self.sy_t, (self.N, 1, 1)), self.from_tiled(self.sR_t, (self.N, 1, 1))

    def posterior(self, out_mean=None, out_cov=None):
        m = self.from_tiled(self.pm_t, out=None if out_mean is None else out_mean.reshape(-1))
        v = self.from_tiled(self.pc_t, out=None if out_cov is None else out_cov.reshape(-1))
        return m.reshape(self.N, 1, 1), v.reshape(self.N, 1, 1)

    def _check_post(self, post):
        """the sweep writes the marginals straight into these: they must be contiguous [N] arrays of the shard's dtype"""
        for t in post:
            if t.dtype != self.dtype or t.numel() != self.N or not t.is_contiguous() or not t.is_cuda:
                raise ValueError('post = (mean, cov): contiguous device tensors of %d %s elements are needed' % (self.N, self.dtype))

    # ---- the passes
    def _args(self, likelihood, method, cubature, lr, power, ensure_psd):
        a = _lib.IterArgs()
        a.N, a.rank, a.world = self.N, self.rank, self.world
        a.dt_t, a.y_t = self.dt_t.data_ptr(), self.y_t.data_ptr()
        a.site_mean_t, a.site_cov_t = self.sy_t.data_ptr(), self.sR_t.data_ptr()
        a.mask_t = ptr(self.mask_t)
        a.post_mean_t, a.post_cov_t = self.pm_t.data_ptr(), self.pc_t.data_ptr()
        keep = []
        if likelihood is not None:
            a.method, a.likelihood, a.lik_param = int(method), int(likelihood.lik_id), float(likelihood.lik_param)
            closed = (method == _lib.BN_METHOD_NEWTON
                      or (method == _lib.BN_METHOD_VI and likelihood.lik_id in (_lib.BN_LIK_GAUSSIAN, _lib.BN_LIK_POISSON_EXP))
                      or (method == _lib.BN_METHOD_EP and likelihood.lik_id == _lib.BN_LIK_GAUSSIAN))
            if not closed:
                cx, cw, Q = host_table(cubature, 1)
                a.Q, a.cub_x_host, a.cub_w_host = Q, cx.ctypes.data, cw.ctypes.data
                keep += [cx, cw]
        a.lr, a.power, a.ensure_psd = float(lr), float(power), int(bool(ensure_psd))
        return a, keep

    def run(self, mode, likelihood=None, method=_lib.BN_METHOD_VI, cubature=None, lr=1.0, power=1.0, ensure_psd=True,
            want_ell=True, post=None):
        """one pass on a single GPU: returns (ell or None, sums[2] or None) as device tensors.  post = (mean, cov): the
        marginals are written straight into these [N(,1,1)] tensors in time order instead of the tiled arrays"""
        a, keep = self._args(likelihood, method, cubature, lr, power, ensure_psd)
        if post is not None:
            self._check_post(post)
            a.post_mean, a.post_cov = post[0].data_ptr(), post[1].data_ptr()
        ell = torch.empty((), dtype=self.dtype, device=self.dev) if want_ell else None
        sums = torch.empty(2, dtype=self.dtype, device=self.dev) if mode != PLAIN else None
        _lib.check(self._fn('bn_iter_pass')(self.kernel.spec(), C.byref(a), int(mode), ptr(ell), ptr(sums), ptr(self.ws),
                                           self.ws.numel(), stream_ptr()))
        return ell, sums

    def reduce(self, want_ell=True):
        """want_ell must match the filter() call of the same pass"""
        a, _ = self._args(None, 0, None, 1.0, 1.0, True)
        a.want_ell = int(bool(want_ell))
        carry = torch.empty(self.kf_len, dtype=self.dtype, device=self.dev)
        _lib.check(self._fn('bn_iter_shard_reduce')(self.kernel.spec(), C.byref(a), ptr(carry), ptr(self.ws), self.ws.numel(),
                                                   stream_ptr()))
        return carry

    def filter(self, kf_carries, want_ell=True):
        a, _ = self._args(None, 0, None, 1.0, 1.0, True)
        a.want_ell = int(bool(want_ell))
        ell = torch.empty((), dtype=self.dtype, device=self.dev) if want_ell else None
        carry = torch.empty(self.rts_len, dtype=self.dtype, device=self.dev)
        _lib.check(self._fn('bn_iter_shard_filter')(self.kernel.spec(), C.byref(a), ptr(kf_carries), ptr(ell), ptr(carry),
                                                   ptr(self.ws), self.ws.numel(), stream_ptr()))
        return ell, carry

    def smooth(self, mode, rts_carries, likelihood=None, method=_lib.BN_METHOD_VI, cubature=None, lr=1.0, power=1.0,
               ensure_psd=True, post=None):
        a, keep = self._args(likelihood, method, cubature, lr, power, ensure_psd)
        if post is not None:
            self._check_post(post)
            a.post_mean, a.post_cov = post[0].data_ptr(), post[1].data_ptr()
        sums = torch.empty(2, dtype=self.dtype, device=self.dev) if mode != PLAIN else None
        _lib.check(self._fn('bn_iter_shard_smooth')(self.kernel.spec(), C.byref(a), int(mode), ptr(rts_carries), ptr(sums),
                                                   ptr(self.ws), self.ws.numel(), stream_ptr()))
        return sums


def linear_posterior():
    """BN_B200_LINEAR_POST=0: the sweeps keep the marginals tiled and a transposition kernel converts them (A/B aid)"""
    import os
    return os.environ.get('BN_B200_LINEAR_POST', '1') != '0'


def cubature_key(cubature):
    return None if cubature is None else id(cubature)


def np_labels_to_float(Y):
    """accepts uint8 / bool labels for the Bernoulli likelihoods (the reference casts to float64, utils.py:264-265)"""
    return np.asarray(Y, dtype=np.float64)
