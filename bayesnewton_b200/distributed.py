"""Time-sharded filter / smoother: the two-level associative scan across GPUs.

The time axis is cut into contiguous shards, one per rank.  Per pass: (1) every rank reduces its
shard to ONE carry -- the filtering element (A, b, C, J, eta), 3d^2+2d doubles, or the smoothing
element (E, g, L), 2d^2+d doubles (bn_*_shard_reduce); (2) the carries are all-gathered over
NCCL/NVLink (264 B / 168 B per rank at d = 3); (3) every rank folds the carries that precede it in
scan order into its incoming state and runs its local pass (bn_*_shard_apply).  Site updates are
pointwise in time and need no communication; scalars (log-likelihood, energy terms) are summed
with one all-reduce.  Associativity: bayesnewton/ops.py:203-219 (filter), :328-335 (smoother).

`filter_smoother_in_shards` runs the same three steps for n_shards shards inside ONE process (no
collective): it is what the single-GPU tests use to validate the carry path at full size.
"""
import torch

from . import _lib
from ._util import as_dev, as_mask, ptr, stream_ptr


def _ws(N, d, D, device):
    nb = _lib.lib().bn_workspace_bytes(int(N), d, D)
    return torch.empty(int(nb), dtype=torch.uint8, device=device), int(nb)


class TimeShard:
    """the local shard of one rank: device tensors for its steps, plus private workspaces that keep
    the chunk prefixes alive between the reduce and the apply call of a pass"""

    def __init__(self, kernel, dt, dt_smoother, rank, world):
        self.kernel = kernel
        if self.spec is None:
            raise NotImplementedError('time sharding needs a kernel with an in-library discretisation')
        self.rank, self.world = rank, world
        self.dt, self.dts = as_dev(dt).reshape(-1), as_dev(dt_smoother).reshape(-1)
        self.N = self.dt.shape[0]
        self.d = _lib.lib().bn_state_dim(self.spec)
        self.D = self.spec.n_components
        dev = self.dt.device
        self.ws_f, self.nb_f = _ws(self.N, self.d, self.D, dev)
        self.ws_s, self.nb_s = _ws(self.N, self.d, self.D, dev)
        self.kf_len = _lib.lib().bn_kf_carry_len(self.d)
        self.rts_len = _lib.lib().bn_rts_carry_len(self.d)
        self._ws_up = None

    @property
    def spec(self):
        """rebuilt from the kernel object on every use: hyper-parameters may change between iterations"""
        return self.kernel.spec() if hasattr(self.kernel, 'spec') else None

    def hyper_key(self):
        s = self.spec
        return (s.family, s.n_components, tuple(s.variance), tuple(s.lengthscale))

    # ---- fused posterior update (bn_up_shard_*): reduce -> [all-gather] -> filter -> [all-gather] -> smooth
    def _up_ws(self):
        if self._ws_up is None:
            nb = _lib.lib().bn_update_posterior_workspace_bytes(self.spec, self.N)
            self._ws_up = (torch.empty(int(nb), dtype=torch.uint8, device=self.dt.device), int(nb))
        return self._ws_up

    def up_reduce(self, y, R, want_grad=False):
        ws, nb = self._up_ws()
        carry = torch.empty(self.kf_len, dtype=torch.float64, device=self.dt.device)
        _lib.check(_lib.lib().bn_up_shard_reduce(self.spec, self.N, self.rank, self.world, ptr(self.dt), ptr(y), ptr(R),
                                                 ptr(carry), int(want_grad), ptr(ws), nb, stream_ptr()))
        return carry

    def up_filter(self, kf_carries, y, R, mask=None, want_ell=True, want_grad=False):
        ws, nb = self._up_ws()
        dev = self.dt.device
        ell = torch.zeros((), dtype=torch.float64, device=dev) if want_ell else None
        carry = torch.empty(self.rts_len, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bn_up_shard_filter(self.spec, self.N, self.rank, self.world, ptr(kf_carries), ptr(self.dt),
                                                 ptr(y), ptr(R), ptr(mask), ptr(ell), ptr(carry), int(want_grad),
                                                 ptr(ws), nb, stream_ptr()))
        return ell, carry

    def up_smooth(self, rts_carries, want_grad=False):
        """want_grad: also returns this shard's share of d ell / d [variance_c..., lengthscale_c...] ([2, NC])"""
        ws, nb = self._up_ws()
        dev = self.dt.device
        sm = torch.empty((self.N, self.D, 1), dtype=torch.float64, device=dev)
        sP = torch.empty((self.N, self.D, self.D), dtype=torch.float64, device=dev)
        g = torch.zeros((2, self.D), dtype=torch.float64, device=dev) if want_grad else None
        _lib.check(_lib.lib().bn_up_shard_smooth(self.spec, self.N, self.rank, self.world, ptr(rts_carries),
                                                 ptr(self.dt), ptr(sm), ptr(sP),
                                                 g[0].data_ptr() if want_grad else None,
                                                 g[1].data_ptr() if want_grad else None, ptr(ws), nb, stream_ptr()))
        if want_grad:
            return sm, sP, g
        return sm, sP

    # ---- filter
    def kf_reduce(self, y, R):
        carry = torch.empty(self.kf_len, dtype=torch.float64, device=self.dt.device)
        _lib.check(_lib.lib().bn_kf_shard_reduce(self.spec, self.N, int(self.rank == 0), ptr(self.dt), ptr(y), ptr(R),
                                                 ptr(carry), ptr(self.ws_f), self.nb_f, stream_ptr()))
        return carry

    def kf_apply(self, carries, y, R, mask=None, want_ell=True, return_predict=False):
        dev = self.dt.device
        ell = torch.zeros((), dtype=torch.float64, device=dev) if want_ell else None
        fm = torch.empty((self.N, self.d, 1), dtype=torch.float64, device=dev)
        fP = torch.empty((self.N, self.d, self.d), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bn_kf_shard_apply(self.spec, self.N, self.rank, self.world, ptr(carries), ptr(self.dt),
                                                ptr(y), ptr(R), ptr(mask), int(return_predict), ptr(ell), ptr(fm),
                                                ptr(fP), ptr(self.ws_f), self.nb_f, stream_ptr()))
        return ell, fm, fP

    # ---- smoother
    def rts_reduce(self, fm, fP):
        carry = torch.empty(self.rts_len, dtype=torch.float64, device=self.dt.device)
        _lib.check(_lib.lib().bn_rts_shard_reduce(self.spec, self.N, int(self.rank == self.world - 1), ptr(self.dts),
                                                  ptr(fm), ptr(fP), ptr(carry), ptr(self.ws_s), self.nb_s,
                                                  stream_ptr()))
        return carry

    def rts_apply(self, carries, fm, fP, return_full=False):
        dev = self.dt.device
        od = self.d if return_full else self.D
        sm = torch.empty((self.N, od, 1), dtype=torch.float64, device=dev)
        sP = torch.empty((self.N, od, od), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bn_rts_shard_apply(self.spec, self.N, self.rank, self.world, ptr(carries), ptr(self.dts),
                                                 ptr(fm), ptr(fP), int(return_full), ptr(sm), ptr(sP), None,
                                                 ptr(self.ws_s), self.nb_s, stream_ptr()))
        return sm, sP


class PeerExchange:
    """carry exchange over NVLink peer memory (csrc/exchange.cu): one symmetric inbox per rank, peer-mapped through
    torch's symmetric-memory rendezvous; every exchange is ONE kernel on the compute stream (no NCCL on the data path)"""

    def __init__(self, world, rank):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.seq = world, rank, 0
        n = int(_lib.lib().bn_carry_exchange_bytes(world)) // 8
        self.buf = symm.empty(n, dtype=torch.float64, device=torch.device('cuda', torch.cuda.current_device()))
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD.group_name)
        self.ptrs = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=self.buf.device)
        torch.cuda.synchronize()
        dist.barrier()  # every inbox is zeroed before anyone stores into it

    def all_gather(self, carry):
        flat = carry.contiguous().reshape(-1)
        out = torch.empty((self.world, flat.numel()), dtype=torch.float64, device=flat.device)
        self.seq += 1
        _lib.check(_lib.lib().bn_carry_exchange(self.ptrs.data_ptr(), self.world, self.rank, flat.data_ptr(), flat.numel(),
                                                self.seq, out.data_ptr(), stream_ptr()))
        return out.reshape((self.world,) + tuple(carry.shape))


_peer = {'tried': False, 'px': None}


def peer_exchange(world):
    """the process-wide PeerExchange, or None when it is not available (CPU / gloo runs, BN_B200_CARRY_EXCHANGE=nccl, or
    the symmetric-memory rendezvous failing on ANY rank: the ranks agree through one all-reduce)"""
    import os
    import torch.distributed as dist
    if _peer['tried']:
        return _peer['px']
    _peer['tried'] = True
    if (world <= 1 or not torch.cuda.is_available() or dist.get_backend() != 'nccl'
            or os.environ.get('BN_B200_CARRY_EXCHANGE', 'p2p') == 'nccl'):
        return None
    px, ok = None, 1
    try:
        px = PeerExchange(world, dist.get_rank())
    except Exception as ex:  # noqa: BLE001 -- any failure means: fall back to NCCL, on all ranks
        ok = 0
        if dist.get_rank() == 0:
            print('[bayesnewton_b200] peer-memory carry exchange unavailable (%s): using NCCL all-gather' % repr(ex)[:200],
                  flush=True)
    flag = torch.tensor([ok], dtype=torch.int32, device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    _peer['px'] = px if int(flag) == 1 else None
    return _peer['px']


def _all_gather(carry, world):
    """carries[world, len]: over NVLink peer memory (one kernel, csrc/exchange.cu) when available, else an all-gather
    on the default process group (NCCL on GPUs, gloo in the CPU tests)"""
    import torch.distributed as dist
    px = peer_exchange(world) if carry.is_cuda else None
    if px is not None:
        return px.all_gather(carry)
    flat = carry.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=carry.dtype, device=carry.device)  # flat: gloo wants the concatenated form
    dist.all_gather_into_tensor(out, flat)
    return out.reshape((world,) + tuple(carry.shape))


def sharded_update_posterior(shard, y, R, mask=None, want_ell=False, want_grad=False):
    """update_posterior (basemodels.py:689-706) on a time-sharded model: the fused library path, 2 carry
    all-gathers.  Every rank passes ITS shard of the sites; returns (ell_local_or_None, post_mean, post_cov)
    and, with want_grad, the LOCAL share [2, NC] of d ell / d (variance, lengthscale) (sum over ranks = total)."""
    c = shard.up_reduce(y, R, want_grad)
    carries = _all_gather(c, shard.world) if shard.world > 1 else c.reshape(1, -1)
    ell, c = shard.up_filter(carries, y, R, mask, want_ell=want_ell, want_grad=want_grad)
    carries = _all_gather(c, shard.world) if shard.world > 1 else c.reshape(1, -1)
    return (ell,) + tuple(shard.up_smooth(carries, want_grad))


def sharded_filter_smoother(shard, y, R, mask=None, want_ell=False):
    """the same update through the stand-alone filter and smoother entry points (bn_kf_shard_*, bn_rts_shard_*)"""
    c = shard.kf_reduce(y, R)
    carries = _all_gather(c, shard.world) if shard.world > 1 else c.reshape(1, -1)
    ell, fm, fP = shard.kf_apply(carries, y, R, mask, want_ell=want_ell)
    c = shard.rts_reduce(fm, fP)
    carries = _all_gather(c, shard.world) if shard.world > 1 else c.reshape(1, -1)
    sm, sP = shard.rts_apply(carries, fm, fP)
    return ell, sm, sP


def sharded_log_lik(shard, y, R, mask=None):
    """compute_log_lik (basemodels.py:726-741): local partial; the caller all-reduces the scalar"""
    c = shard.kf_reduce(y, R)
    carries = _all_gather(c, shard.world) if shard.world > 1 else c.reshape(1, -1)
    dev = shard.dt.device
    ell = torch.zeros((), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().bn_kf_shard_apply(shard.spec, shard.N, shard.rank, shard.world, ptr(carries), ptr(shard.dt),
                                            ptr(y), ptr(R), ptr(mask), 0, ptr(ell), None, None, ptr(shard.ws_f),
                                            shard.nb_f, stream_ptr()))
    return ell


def shard_bounds(N, world):
    return [N * r // world for r in range(world + 1)]


def filter_smoother_in_shards(kernel, dt, y, R, mask, n_shards):
    """single-process run of the sharded algorithm over n_shards shards (validation of the carry path)"""
    dt, y, R = as_dev(dt).reshape(-1), as_dev(y), as_dev(R)
    mk = as_mask(mask)
    N = dt.shape[0]
    dts = torch.cat([dt[1:], torch.zeros(1, dtype=dt.dtype, device=dt.device)])
    b = shard_bounds(N, n_shards)
    shards = [TimeShard(kernel, dt[b[r]:b[r + 1]], dts[b[r]:b[r + 1]], r, n_shards) for r in range(n_shards)]
    D = shards[0].D
    ys = [y.reshape(N, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    Rs = [R.reshape(N, D, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    ms = [None if mk is None else mk.reshape(N, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    carries = torch.stack([s.kf_reduce(ys[r], Rs[r]) for r, s in enumerate(shards)])
    filt = [s.kf_apply(carries, ys[r], Rs[r], ms[r]) for r, s in enumerate(shards)]
    carries = torch.stack([s.rts_reduce(filt[r][1], filt[r][2]) for r, s in enumerate(shards)])
    smo = [s.rts_apply(carries, filt[r][1], filt[r][2]) for r, s in enumerate(shards)]
    return dict(ell=sum(f[0] for f in filt), filter_mean=torch.cat([f[1] for f in filt]),
                filter_cov=torch.cat([f[2] for f in filt]), post_mean=torch.cat([s[0] for s in smo]),
                post_cov=torch.cat([s[1] for s in smo]))


def update_posterior_in_shards(kernel, dt, y, R, mask, n_shards, want_grad=False):
    """single-process run of the FUSED sharded update (bn_up_shard_*) over n_shards shards"""
    dt, y, R = as_dev(dt).reshape(-1), as_dev(y), as_dev(R)
    mk = as_mask(mask)
    N = dt.shape[0]
    b = shard_bounds(N, n_shards)
    shards = [TimeShard(kernel, dt[b[r]:b[r + 1]], dt[b[r]:b[r + 1]], r, n_shards) for r in range(n_shards)]
    D = shards[0].D
    ys = [y.reshape(N, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    Rs = [R.reshape(N, D, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    ms = [None if mk is None else mk.reshape(N, D)[b[r]:b[r + 1]].contiguous() for r in range(n_shards)]
    carries = torch.stack([s.up_reduce(ys[r], Rs[r], want_grad) for r, s in enumerate(shards)])
    filt = [s.up_filter(carries, ys[r], Rs[r], ms[r], want_grad=want_grad) for r, s in enumerate(shards)]
    carries = torch.stack([f[1] for f in filt])
    smo = [s.up_smooth(carries, want_grad) for s in shards]
    out = dict(ell=sum(f[0] for f in filt), post_mean=torch.cat([s[0] for s in smo]),
               post_cov=torch.cat([s[1] for s in smo]))
    if want_grad:
        out['grad'] = sum(s[2] for s in smo)
    return out


class TimeShardedMarkovGP:
    """One rank's view of a time-sharded temporal Markov GP: the host mirror of
    MarkovGaussianProcess + an inference mixin (basemodels.py:625-764, inference.py:65-90) where the
    time axis is partitioned over the ranks of the default process group.

    Every rank constructs it with ITS contiguous slice of (dt, Y) -- `dt_next` is the first dt of the
    right neighbour (0 on the last rank), needed by the smoother's shifted step array
    (basemodels.py:700).  Per iteration: 5 carry all-gathers + 1 all-reduce of 3 scalars.
    """

    def __init__(self, kernel, likelihood, dt_local, Y_local, dt_next, method, rank, world, power=1.0):
        from .basemodels import GaussianDistribution
        self.kernel, self.likelihood, self.method, self.power = kernel, likelihood, method, power
        self.rank, self.world = rank, world
        self._energy_cache = None
        dt = as_dev(dt_local).reshape(-1)
        dts = torch.cat([dt[1:], torch.full((1,), float(dt_next), dtype=dt.dtype, device=dt.device)])
        self.shard = TimeShard(kernel, dt, dts, rank, world)
        self.Y = as_dev(Y_local).reshape(-1)
        self.N, D = self.shard.N, self.shard.D
        self.state_dim = self.shard.d
        dev = dt.device
        eye = torch.eye(D, dtype=torch.float64, device=dev).repeat(self.N, 1, 1)
        self.pseudo_likelihood = GaussianDistribution(
            mean=torch.zeros((self.N, D, 1), dtype=torch.float64, device=dev), covariance=1e2 * eye,
            nat1=torch.zeros((self.N, D, 1), dtype=torch.float64, device=dev), nat2=1e-2 * eye)
        self.posterior_mean = torch.zeros((self.N, D, 1), dtype=torch.float64, device=dev)
        self.posterior_variance = torch.eye(D, dtype=torch.float64, device=dev).repeat(self.N, 1, 1)
        nan = torch.isnan(self.Y)
        self.mask_pseudo_y = nan.to(torch.uint8).contiguous() if (D == 1 and bool(nan.any())) else None
        self._ws = _ws(self.N, self.state_dim, D, dev)

    def update_posterior(self, want_grad=False):
        pl = self.pseudo_likelihood
        out = sharded_update_posterior(self.shard, pl.mean, pl.covariance, self.mask_pseudo_y, want_ell=True,
                                       want_grad=want_grad)
        key = (pl.version, self.shard.hyper_key())
        self._ell_cache = (out[0], key)  # the local log-likelihood partial of exactly these sites and hyper-parameters
        self._grad_cache = (out[3], key) if want_grad else None
        self._energy_cache = None
        self.posterior_mean, self.posterior_variance = out[1], out[2]

    def _site_args(self, cubature=None):
        a, keep = self.likelihood.site_args(self.method, self.Y, self.posterior_mean, self.posterior_variance,
                                            cubature, self.power)
        a.nat1, a.nat2 = self.pseudo_likelihood.nat1_.data_ptr(), self.pseudo_likelihood.nat2_.data_ptr()
        return a, keep

    # ---- fused iteration on tiled resident state (fused.py): the same two passes as on one GPU, each in its three
    # phases with the carries of all ranks exchanged in between
    def _fused_ok(self):
        import os
        from . import fused
        return (os.environ.get('BN_B200_FUSED', '1') != '0' and self.shard.D == 1
                and self.method in (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON)  # (the sharded energy is VI / Newton)
                and fused.supported(self.shard.spec, self.likelihood, self.method))

    def _fused_state(self):
        from . import fused
        st = getattr(self, '_fused', None)
        pl = self.pseudo_likelihood
        if st is None:
            st = self._fused = fused.FusedShard(self.kernel, self.shard.dt, self.Y, self.mask_pseudo_y, self.rank, self.world)
            st.sites_version = None
        if pl.source is not st or st.sites_version != pl.version:
            st.load_sites(pl.mean, pl.covariance)
            st.sites_version = pl.version
        return st

    def load_inputs(self, dt, Y, dt_next=None):
        """replace this shard's step lengths and observations (device or pinned-host tensors; Y may hold uint8 labels)"""
        dt = as_dev(dt).reshape(-1)
        labels = torch.is_tensor(Y) and not Y.dtype.is_floating_point  # integer labels cannot hold a missing value
        self.Y = as_dev(Y).reshape(-1)
        last = self.shard.dts[-1:] if dt_next is None else torch.full((1,), float(dt_next), dtype=dt.dtype, device=dt.device)
        self.shard.dt, self.shard.dts = dt, torch.cat([dt[1:], last])
        nan = None if labels else torch.isnan(self.Y)
        self.mask_pseudo_y = nan.to(torch.uint8).contiguous() if (nan is not None and self.shard.D == 1 and bool(nan.any())) else None
        self._ell_cache = self._grad_cache = self._energy_cache = None
        st = getattr(self, '_fused', None)
        if st is not None:
            st.set_dt(dt)
            st.set_data(self.Y, self.mask_pseudo_y, scan_nan=not labels)

    def _fused_pass(self, st, mode, lr, cubature, ensure_psd, want_ell):
        c = st.reduce(want_ell=want_ell)
        carries = _all_gather(c, self.world) if self.world > 1 else c.reshape(1, -1)
        ell, c = st.filter(carries, want_ell=want_ell)
        carries = _all_gather(c, self.world) if self.world > 1 else c.reshape(1, -1)
        from . import fused
        post = (self.posterior_mean, self.posterior_variance) if (mode != fused.SITES and fused.linear_posterior()) else None
        return ell, st.smooth(mode, carries, self.likelihood, self.method, cubature, lr, self.power, ensure_psd, post=post)

    def _inference_fused(self, lr, cubature, ensure_psd):
        from . import fused
        st = self._fused_state()
        pl = self.pseudo_likelihood
        _, d = self._fused_pass(st, fused.SITES, lr, cubature, ensure_psd, False)
        pl.version += 1
        pl.source, st.sites_version = st, pl.version
        ell, sums = self._fused_pass(st, fused.ENERGY, lr, cubature, ensure_psd, True)
        if not fused.linear_posterior():
            self.posterior_mean, self.posterior_variance = st.posterior(self.posterior_mean, self.posterior_variance)
        key = (pl.version, self.shard.hyper_key())
        self._ell_cache = (ell, key)
        self._grad_cache = None
        self._energy_cache = (sums, key + (float(self.likelihood.lik_param), float(self.likelihood.lik_param2), fused.cubature_key(cubature)))
        return d  # local sums of |delta nat1|, |delta nat2| (before damping)

    def inference(self, lr=1.0, cubature=None, ensure_psd=True, want_grad=False):
        """want_grad: the closing posterior update also accumulates d log-lik / d hyper-parameters, which
        energy_and_grad() then serves without another pass"""
        if not want_grad and self._fused_ok():
            return self._inference_fused(lr, cubature, ensure_psd)
        self.update_posterior()
        a, keep = self._site_args(cubature)
        a.lr, a.ensure_psd = float(lr), int(bool(ensure_psd))
        pl = self.pseudo_likelihood
        a.site_mean, a.site_cov = pl.mean_.data_ptr(), pl.covariance_.data_ptr()
        ws, nb = self._ws
        _lib.check(_lib.lib().bn_site_update(a, ptr(ws), nb, stream_ptr()))
        pl.version += 1
        self.update_posterior(want_grad)

    def energy_and_grad(self, cubature=None):
        """(energy, d energy / d [variances; lengthscales] as a [2, NC] tensor): what
        objax.GradValues(model.energy, model.vars()) returns for the kernel hyper-parameters (README.md:56-70),
        before the softplus chain of kernels.py:80-95.  d energy = - d log-lik (the other terms hold the
        kernel hyper-parameters only through StateVars)."""
        pl = self.pseudo_likelihood
        cache = getattr(self, '_grad_cache', None)
        if cache is None or cache[1] != (pl.version, self.shard.hyper_key()):
            self.update_posterior(want_grad=True)
        g = self._grad_cache[0].clone()
        E = self.energy(cubature)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(g)
        return E, -g

    def energy(self, cubature=None):
        """VI / Newton energy (inference.py:130-154,197-222): three local sums, one all-reduce"""
        if self.method not in (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON):
            raise NotImplementedError('time-sharded energy is implemented for VI and Newton')
        pl = self.pseudo_likelihood
        dev = self.Y.device
        parts = torch.zeros(3, dtype=torch.float64, device=dev)
        from . import fused
        key = (pl.version, self.shard.hyper_key())
        ec = getattr(self, '_energy_cache', None)
        if ec is not None and ec[1] == key + (float(self.likelihood.lik_param), float(self.likelihood.lik_param2), fused.cubature_key(cubature)):
            # the closing pass of the fused inference() summed both terms in its smoother epilogue
            parts[0:2] = ec[0]
            parts[2:3] = self._ell_cache[0]
            if self.world > 1:
                parts = _all_gather(parts, self.world).sum(dim=0)
            return -(parts[0] - (parts[1] - parts[2]))
        a, keep = self._site_args(cubature)
        ws, nb = self._ws
        if a.D == 1:  # single latent: both sums in one pass over the posterior marginals
            a.site_mean, a.site_cov = pl.mean.data_ptr(), pl.covariance.data_ptr()
            _lib.check(_lib.lib().bn_energy_terms(a, ptr(self.mask_pseudo_y), parts[0:2].data_ptr(), ptr(ws), nb,
                                                  stream_ptr()))
        else:
            _lib.check(_lib.lib().bn_expected_density(a, None, parts[0:1].data_ptr(), ptr(ws), nb, stream_ptr()))
            _lib.check(_lib.lib().bn_gaussian_expected_log_lik(
                self.N, a.D, ptr(pl.mean), ptr(self.posterior_mean), ptr(self.posterior_variance), ptr(pl.covariance),
                ptr(self.mask_pseudo_y), None, parts[1:2].data_ptr(), ptr(ws), nb, stream_ptr()))
        cache = getattr(self, '_ell_cache', None)
        if cache is not None and cache[1] == key:  # same sites, same kernel: the filter pass of update_posterior
            parts[2:3] = cache[0]
        else:
            parts[2:3] = sharded_log_lik(self.shard, pl.mean, pl.covariance, self.mask_pseudo_y)
        if self.world > 1:  # sum over ranks in rank order (bit-stable): gather + fixed-order sum
            parts = _all_gather(parts, self.world).sum(dim=0)
        return -(parts[0] - (parts[1] - parts[2]))
