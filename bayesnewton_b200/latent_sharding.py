"""Latent-sharded Markov GP: independent latents partitioned over ranks (finding F10).

Reference model: ``Markov{Variational,Newton,ExpectationPropagation}GP(kernel=Independent([k_1..k_C]),
likelihood=HeteroscedasticNoise())`` (demos/heteroscedastic.py:49-56).  With the default ``ensure_psd=True``
every site precision is diagonal (utils.py:89-96, applied at inference.py:119-120,186-187,275-276) and the
``Independent`` prior is block diagonal (kernels.py:1535-1583), so the joint filter/smoother of
basemodels.py:689-706 is EXACTLY separable per latent: rank r runs the fused posterior update of ITS latents over
all N steps with no communication, and the log-likelihood is the sum of the per-latent ones.

The site update (inference.py:105-128,170-195,238-284) is pointwise in time but needs ALL latents of a step
(likelihoods.py:561-664), so it runs in a second layout -- rank q owns time slice q of every latent -- and
the two layouts are bridged by an all-to-all (the one real exchange step of this path):

    inference():  UP_l -> [L->T: posterior marginals] -> site update_t -> [T->L: new sites] -> UP_l
    energy()   :  [L->T] -> expected density_t, expected pseudo log-lik_t, sum_l log-lik -> one all-reduce

Per step and latent an exchange moves 2 doubles (mean, variance) each way.  Sites live in the time layout
(both parametrisations, as basemodels.py:52-100); the latent layout holds the pseudo observations the filter
consumes and the posterior marginals it produces.

``backend`` executes the per-rank array work (libbn_b200 through ctypes by default -- there is no CPU path in
the product; the CPU tests plug the host emulation of the same kernel bodies in here) and ``comm`` the two
collectives (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
import torch

from . import _lib
from .distributed import shard_bounds


# ------------------------------------------------------------------------------------------ collectives
class NullComm:
    world = 1

    def all_to_all(self, send, out_splits, in_splits):
        return send

    def all_reduce(self, t):
        return t

    def all_gather(self, t):
        return t.reshape((1,) + tuple(t.shape))


class TorchDistComm:
    """default process group (NCCL over NVLink on the GPU box)"""

    def __init__(self, world):
        self.world = world

    def all_to_all(self, send, out_splits, in_splits):
        import torch.distributed as dist
        out = torch.empty(int(sum(out_splits)), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(out_splits), input_split_sizes=list(in_splits))
        return out

    def all_reduce(self, t):
        import torch.distributed as dist
        dist.all_reduce(t)
        return t

    def all_gather(self, t):
        import torch.distributed as dist
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat)
        return out.reshape((self.world,) + tuple(t.shape))


# ------------------------------------------------------------------------------------------ per-rank array work
class LibBackend:
    """libbn_b200 through the C ABI; owns its workspaces (several rank objects may share a device in tests)"""

    def __init__(self):
        self._ws = {}

    def _buf(self, key, nbytes, device):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws, ws.numel()

    def zeros(self, shape):
        from ._util import device
        return torch.zeros(shape, dtype=torch.float64, device=device())

    def to_dev(self, x):
        from ._util import as_dev
        return as_dev(x)

    def update_posterior(self, spec, dt, y, R, want_grad=False):
        from ._util import ptr, stream_ptr
        L = _lib.lib()
        N, D = dt.shape[0], spec.n_components
        ell = torch.zeros((), dtype=torch.float64, device=dt.device)
        pm = torch.empty((N, D, 1), dtype=torch.float64, device=dt.device)
        pc = torch.empty((N, D, D), dtype=torch.float64, device=dt.device)
        ws, nb = self._buf('up', L.bn_update_posterior_workspace_bytes(spec, N), dt.device)
        if want_grad:
            g = torch.zeros((2, D), dtype=torch.float64, device=dt.device)
            _lib.check(L.bn_update_posterior_grad(spec, N, ptr(dt), ptr(y), ptr(R), ptr(ell), ptr(pm), ptr(pc),
                                                  g[0].data_ptr(), g[1].data_ptr(), ptr(ws), nb, stream_ptr()))
            return ell, pm, pc, g
        _lib.check(L.bn_update_posterior(spec, N, ptr(dt), ptr(y), ptr(R), None, ptr(ell), ptr(pm), ptr(pc), ptr(ws), nb,
                                         stream_ptr()))
        return ell, pm, pc, None

    def _args(self, lik, method, power, Y, pm, pc, nat1, nat2, cubature):
        a, keep = lik.site_args(method, Y, pm, pc, cubature, power)
        a.nat1, a.nat2 = nat1.data_ptr(), nat2.data_ptr()
        return a, keep

    def site_update(self, lik, method, power, Y, pm, pc, nat1, nat2, site_mean, site_cov, lr, cubature):
        from ._util import ptr, stream_ptr
        a, keep = self._args(lik, method, power, Y, pm, pc, nat1, nat2, cubature)
        a.lr, a.ensure_psd = float(lr), 1
        a.site_mean, a.site_cov = site_mean.data_ptr(), site_cov.data_ptr()
        diffs = torch.zeros(2, dtype=torch.float64, device=pm.device)
        a.diffs = diffs.data_ptr()
        ws, nb = self._buf('site', _lib.lib().bn_workspace_bytes(int(a.N), 2 * a.D, a.D), pm.device)
        _lib.check(_lib.lib().bn_site_update(a, ptr(ws), nb, stream_ptr()))
        return diffs

    def expected_density(self, lik, method, power, Y, pm, pc, nat1, nat2, cubature):
        from ._util import ptr, stream_ptr
        a, keep = self._args(lik, method, power, Y, pm, pc, nat1, nat2, cubature)
        out = torch.zeros((), dtype=torch.float64, device=pm.device)
        ws, nb = self._buf('site', _lib.lib().bn_workspace_bytes(int(a.N), 2 * a.D, a.D), pm.device)
        _lib.check(_lib.lib().bn_expected_density(a, None, ptr(out), ptr(ws), nb, stream_ptr()))
        return out

    def gaussian_expected_log_lik(self, py, pm, pc, pv):
        from ._util import ptr, stream_ptr
        N, D = py.shape[0], py.shape[1]
        out = torch.zeros((), dtype=torch.float64, device=pm.device)
        ws, nb = self._buf('site', _lib.lib().bn_workspace_bytes(N, 2 * D, D), pm.device)
        _lib.check(_lib.lib().bn_gaussian_expected_log_lik(N, D, ptr(py), ptr(pm), ptr(pc), ptr(pv), None, None, ptr(out),
                                                           ptr(ws), nb, stream_ptr()))
        return out

    def ep_pseudo_density(self, power, py, pv, pm, pc, nat1, nat2):
        from ._util import ptr, stream_ptr
        N, D = py.shape[0], py.shape[1]
        out = torch.zeros((), dtype=torch.float64, device=pm.device)
        ws, nb = self._buf('site', _lib.lib().bn_workspace_bytes(N, 2 * D, D), pm.device)
        _lib.check(_lib.lib().bn_ep_pseudo_density(N, D, float(power), 1, ptr(py), ptr(pv), ptr(pm), ptr(pc), ptr(nat1),
                                                   ptr(nat2), None, ptr(out), ptr(ws), nb, stream_ptr()))
        return out


# ------------------------------------------------------------------------------------------ the model
class LatentShardedMarkovGP:
    """one rank's view; every rank constructs it with the FULL (dt, Y) and keeps only what its two layouts need"""

    def __init__(self, kernel, likelihood, X, Y, method, rank=0, world=1, power=1.0, backend=None, comm=None):
        import numpy as np
        from .basemodels import input_admin
        ks = getattr(kernel, 'kernels', None)
        if ks is None or len({k.family for k in ks}) != 1:
            raise NotImplementedError('latent sharding needs an Independent stack of one Matern family')
        C = len(ks)
        if likelihood.num_latents != C:
            raise ValueError('the likelihood takes %d latents, the kernel stacks %d' % (likelihood.num_latents, C))
        if C % world != 0:
            raise ValueError('%d latents do not split over %d ranks' % (C, world))
        self.kernel, self.likelihood, self.method, self.power = kernel, likelihood, method, power
        self.rank, self.world, self.C = rank, world, C
        self.lpr = C // world
        self.lo, self.hi = rank * self.lpr, (rank + 1) * self.lpr
        self.be = backend if backend is not None else LibBackend()
        self.comm = comm if comm is not None else (TorchDistComm(world) if world > 1 else NullComm())
        t, Yh, dt = input_admin(X, Y)
        if np.isnan(Yh).any():
            raise NotImplementedError('missing observations are outside the latent-sharded path')
        N = self.N = t.shape[0]
        self.b = shard_bounds(N, world)
        self.t0, self.t1 = self.b[rank], self.b[rank + 1]
        nt = self.nt = self.t1 - self.t0
        be = self.be
        self.dt = be.to_dev(dt)
        self.Y_t = be.to_dev(Yh[self.t0:self.t1, 0].copy())
        # time layout: the sites, both parametrisations (basemodels.py:52-100,130-133: mean 0, cov 100 I)
        eye_t = torch.eye(C, dtype=torch.float64).repeat(nt, 1, 1)
        self.nat1_t = be.zeros((nt, C, 1))
        self.nat2_t = be.to_dev(1e-2 * eye_t)
        self.site_mean_t = be.zeros((nt, C, 1))
        self.site_cov_t = be.to_dev(1e2 * eye_t)
        self.post_mean_t = be.zeros((nt, C, 1))
        self.post_cov_t = be.to_dev(eye_t)
        # latent layout: what the filter of this rank's latents consumes / produces
        eye_l = torch.eye(self.lpr, dtype=torch.float64).repeat(N, 1, 1)
        self.pseudo_y_l = be.zeros((N, self.lpr, 1))
        self.pseudo_var_l = be.to_dev(1e2 * eye_l)
        self.post_mean_l = be.zeros((N, self.lpr, 1))
        self.post_cov_l = be.to_dev(eye_l)
        self.ell_l = None
        self._post_t_fresh = False
        self._grad_l = None

    # ---- the local spec is rebuilt per call so hyper-parameter changes are seen
    def _spec(self):
        ks = self.kernel.kernels[self.lo:self.hi]
        return _lib.kernel_spec(ks[0].family, [k.variance for k in ks], [k.lengthscale for k in ks])

    def update_posterior(self, want_grad=False):
        """this rank's latents over all N steps: no communication (F10)"""
        self.ell_l, self.post_mean_l, self.post_cov_l, g = self.be.update_posterior(
            self._spec(), self.dt, self.pseudo_y_l, self.pseudo_var_l, want_grad)
        self._grad_l = g
        self._post_t_fresh = False

    # ---- layout exchanges (w doubles per step: lpr means + lpr^2 covariance block)
    def _pack(self, mean, cov, rows=None):
        n = mean.shape[0]
        return torch.cat([mean.reshape(n, -1), cov.reshape(n, -1)], dim=1)

    def _latent_to_time(self):
        if self._post_t_fresh:
            return
        l, w = self.lpr, self.lpr + self.lpr * self.lpr
        packed = self._pack(self.post_mean_l, self.post_cov_l)                       # [N, w], time-major: dest q = rows b[q]:b[q+1]
        in_splits = [(self.b[q + 1] - self.b[q]) * w for q in range(self.world)]
        out_splits = [self.nt * w] * self.world
        recv = self.comm.all_to_all(packed.reshape(-1), out_splits, in_splits).reshape(self.world, self.nt, w)
        pm = self.post_mean_t
        pc = self.post_cov_t
        pc.zero_()
        for r in range(self.world):  # latents of source rank r
            a, b_ = r * l, (r + 1) * l
            pm[:, a:b_, 0] = recv[r, :, :l]
            pc[:, a:b_, a:b_] = recv[r, :, l:].reshape(self.nt, l, l)
        self._post_t_fresh = True

    def _time_to_latent(self):
        l, w = self.lpr, self.lpr + self.lpr * self.lpr
        chunks = []
        for q in range(self.world):  # to dest q: its latents of my time slice
            a, b_ = q * l, (q + 1) * l
            chunks.append(self._pack(self.site_mean_t[:, a:b_, :], self.site_cov_t[:, a:b_, a:b_]).reshape(-1))
        in_splits = [self.nt * w] * self.world
        out_splits = [(self.b[r + 1] - self.b[r]) * w for r in range(self.world)]
        recv = self.comm.all_to_all(torch.cat(chunks), out_splits, in_splits).reshape(self.N, w)  # sources in time order
        self.pseudo_y_l = recv[:, :l].reshape(self.N, l, 1).contiguous()
        self.pseudo_var_l = recv[:, l:].reshape(self.N, l, l).contiguous()

    # ---- inference.py:65-90
    def inference(self, lr=1.0, cubature=None, ensure_psd=True, want_grad=False):
        if not ensure_psd:
            raise NotImplementedError('ensure_psd=False gives full site precisions: the latents no longer separate (F10)')
        self.update_posterior()
        self._latent_to_time()
        diffs = self.be.site_update(self.likelihood, self.method, self.power, self.Y_t, self.post_mean_t, self.post_cov_t,
                                    self.nat1_t, self.nat2_t, self.site_mean_t, self.site_cov_t, lr, cubature)
        self._time_to_latent()
        self.update_posterior(want_grad)
        return diffs

    def energy(self, cubature=None):
        """VI / Newton (inference.py:130-154,197-222) and power-EP (:286-325) energies: local sums, one all-reduce"""
        self._latent_to_time()
        be = self.be
        parts = be.zeros((4,))
        parts[0] = be.expected_density(self.likelihood, self.method, self.power, self.Y_t, self.post_mean_t, self.post_cov_t,
                                       self.nat1_t, self.nat2_t, cubature)
        if self.method in (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON):
            parts[1] = be.gaussian_expected_log_lik(self.site_mean_t, self.post_mean_t, self.post_cov_t, self.site_cov_t)
        elif self.method == _lib.BN_METHOD_EP:
            parts[1] = be.ep_pseudo_density(self.power, self.site_mean_t, self.site_cov_t, self.post_mean_t,
                                            self.post_cov_t, self.nat1_t, self.nat2_t)
        else:
            raise NotImplementedError('latent-sharded energy: VI, Newton and EP')
        parts[2] = self.ell_l
        self.comm.all_reduce(parts)
        if self.method == _lib.BN_METHOD_EP:
            return -(parts[2] + 1.0 / self.power * (parts[0] - parts[1]))
        return -(parts[0] - (parts[1] - parts[2]))

    def energy_and_grad(self, cubature=None):
        """(energy, d energy / d [variances; lengthscales] [2, C]); each rank differentiates its own latents"""
        if self._grad_l is None:
            self.update_posterior(want_grad=True)
        E = self.energy(cubature)
        g = self.comm.all_gather(self._grad_l)              # [world, 2, lpr]
        return E, -g.permute(1, 0, 2).reshape(2, self.C)

    # ---- posterior marginals of ALL latents for this rank's time slice, [nt, C, 1], [nt, C, C]
    def posterior_time_slice(self):
        self._latent_to_time()
        return self.post_mean_t, self.post_cov_t
