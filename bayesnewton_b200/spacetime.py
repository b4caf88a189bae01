"""Spatio-temporal Markov GP: host mirror of the reference's spatio-temporal branch, executed by libbn_b200's
dense path (csrc/st.cu).

    Separable                      bayesnewton/kernels.py:1666-1683
    SpatioTemporalKernel           kernels.py:385-586 (sparse=True; conditional 'Full' / 'DTC' / 'FIC')
    SpatioTemporalMixin            the spatio-temporal branches of MarkovGaussianProcess (basemodels.py:625-764):
                                   compute_full_pseudo_lik (:676-687), update_posterior (:689-706), compute_kl
                                   (:708-724), compute_log_lik (:726-741), conditional_posterior_to_data (:743-764)
    st_kalman_filter / st_rts_smoother   kalman_filter / rauch_tung_striebel_smoother (ops.py:256-285, 357-380)
                                   for a SpatioTemporalKernel

What lives where: the per-hyper-parameter constants of the spatial conditional (K_zz, its Cholesky factor, the
projection B = K_rz K_zz^-1 L_zz and the diagonal of the conditional covariance: M x M algebra, once per
hyper-parameter setting) are formed on the host in float64; everything that scales with the number of time steps
runs on the GPU.  The reference evaluates the spatial conditional per time step (vmap over time,
kernels.py:494-501); this path requires the spatial inputs to be the same at every step (gridded data, the case
the reference's own TODO names, :492-496) and raises otherwise.  Sites of a factorising likelihood are diagonal
in data space (likelihoods.py:383), so they are stored as their diagonals ([N_t, N_s]) and exposed in the
reference's [N_t, N_s, N_s] shapes on request.
"""
import numpy as np
import torch

from . import _lib
from ._util import as_dev, as_mask, device, ptr, stream_ptr

_st_workspaces = {}


def st_workspace(spec, M, N, Ns):
    need = _lib.lib().bn_st_workspace_bytes(spec, int(M), int(N), int(Ns))
    if need == 0:
        raise _lib.BnError(_lib.lib().bn_last_error().decode())
    key = torch.cuda.current_device()
    ws = _st_workspaces.get(key)
    if ws is None or ws.numel() < need:
        _st_workspaces.pop(key, None)
        ws = torch.empty(int(need), dtype=torch.uint8, device=device())
        _st_workspaces[key] = ws
    return ws, ws.numel()


class Separable:
    """product of 1-D kernels, one per spatial dimension (kernels.py:1666-1683)"""

    def __init__(self, kernels):
        self.kernels = list(kernels)
        self.num_kernels = len(self.kernels)

    def K(self, X, X2):
        X, X2 = np.asarray(X, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        out = self.kernels[0].K(X[:, :1], X2[:, :1])
        for i in range(1, self.num_kernels):
            out = out * self.kernels[i].K(X[:, i:i + 1], X2[:, i:i + 1])
        return out

    __call__ = K


class SpatioTemporalKernel:
    """temporal SDE prior (x) spatial kernel on inducing points z (kernels.py:385-586)"""

    def __init__(self, temporal_kernel, spatial_kernel, z=None, conditional=None, sparse=True, opt_z=False,
                 spatial_dims=None):
        if not sparse:
            raise NotImplementedError('sparse=False (B = L_zz, kernels.py:502-505) is not on the dense path yet')
        if z is None:
            raise NotImplementedError('please provide the spatial inducing inputs z')
        self.temporal_kernel, self.spatial_kernel = temporal_kernel, spatial_kernel
        z = np.asarray(z, dtype=np.float64)
        self.z = z[:, None] if z.ndim < 2 else z
        self.M = self.z.shape[0]
        self.sparse = sparse
        conditional = 'Full' if conditional is None else conditional
        if conditional.lower() not in ('full', 'dtc', 'fic', 'fitc'):
            raise NotImplementedError('conditional method not recognised')
        self.conditional = conditional.lower()

    variance = property(lambda self: self.temporal_kernel.variance)
    temporal_lengthscale = property(lambda self: self.temporal_kernel.lengthscale)
    state_dim = property(lambda self: self.temporal_kernel.state_dim)

    def spec(self):
        """the temporal component, which is what the dense kernels discretise per step"""
        return self.temporal_kernel.spec()

    def _ks(self, A, B):
        k = self.spatial_kernel
        A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
        if isinstance(k, Separable):
            return k.K(A, B)
        if A.shape[1] > 1:  # isotropic kernel on several spatial dimensions: the scaled Euclidean distance over ALL columns
            d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1) / k.lengthscale ** 2   # kernels.py:97-104
            return k.K_r(np.sqrt(np.maximum(d2, 1e-36)))
        return k.K(A[:, :1], B[:, :1])

    def K(self, X, X2):
        X, X2 = np.asarray(X, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        return self.temporal_kernel.K(X[:, :1], X2[:, :1]) * self._ks(X[:, 1:], X2[:, 1:])

    def inducing_precision(self):
        """(K_zz^-1, chol(K_zz))  (kernels.py:508-515)"""
        Kzz = self._ks(self.z, self.z)
        Lzz = np.linalg.cholesky(Kzz)
        Li = np.linalg.solve(Lzz, np.eye(self.M))
        return Li.T @ Li, Lzz

    def spatial_conditional(self, X=None, R=None, predict=False):
        """f(X, R) | u ~ N(B u, C) for ONE set of spatial inputs R [N_s, n_dims]: B [N_s, M], C [N_s, N_s]
        (kernels.py:486-506; the reference returns the same pair tiled over time)"""
        Qzz, Lzz = self.inducing_precision()
        R = np.asarray(R, dtype=np.float64).reshape((-1,) + self.z.shape[1:])
        Krz = self._ks(R, self.z)
        K = Krz @ Qzz
        B = K @ Lzz
        if self.conditional == 'dtc':
            C = np.zeros((R.shape[0], R.shape[0]))
        else:
            resid = self._ks(R, R) - K @ Krz.T
            if self.conditional in ('fic', 'fitc'):
                resid = np.diag(np.diag(resid))
            C = self.temporal_kernel.variance * resid  # temporal_kernel.K(t, t) at a single time stamp
        return B, C

    def stationary_covariance(self):
        return np.kron(np.eye(self.M), self.temporal_kernel.stationary_covariance())

    def measurement_model(self):
        return np.kron(np.eye(self.M), self.temporal_kernel.measurement_model())

    def feedback_matrix(self):
        return np.kron(np.eye(self.M), self.temporal_kernel.feedback_matrix())

    def state_transition(self, dt):
        A = self.temporal_kernel.state_transition(dt)
        A = A.cpu().numpy() if torch.is_tensor(A) else np.asarray(A)
        return np.kron(np.eye(self.M), A)


def st_kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict=False, want_ell=True):
    """kalman_filter for a SpatioTemporalKernel: ell, (means [N,d,1], covs [N,d,d]).  The dense path is the
    reference's sequential recursion (parallel is accepted and ignored: the associative-scan form of a d = M n
    state costs ~10 d^3 per combine and is not built)."""
    dt = as_dev(dt).reshape(-1)
    N, M = dt.shape[0], kernel.M
    spec = kernel.spec()
    d = M * kernel.state_dim
    y, R = as_dev(y), as_dev(noise_cov)
    if y.numel() != N * M or R.numel() != N * M * M:
        raise ValueError('y must be [N,%d,1] and noise_cov [N,%d,%d] for N = %d steps' % (M, M, M, N))
    mk = as_mask(mask)
    ell = torch.zeros((), dtype=torch.float64, device=dt.device) if want_ell else None
    means = torch.empty((N, d, 1), dtype=torch.float64, device=dt.device)
    covs = torch.empty((N, d, d), dtype=torch.float64, device=dt.device)
    ws, nb = st_workspace(spec, M, N, M)
    _lib.check(_lib.lib().bn_st_kalman_filter(spec, M, N, ptr(dt), ptr(y), ptr(R), ptr(mk), int(bool(return_predict)),
                                              ptr(ell), ptr(means), ptr(covs), ptr(ws), nb, stream_ptr()))
    return ell, (means, covs)


def st_rts_smoother(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False, want_gains=True):
    """rauch_tung_striebel_smoother for a SpatioTemporalKernel: (means, covs, gains)"""
    dt = as_dev(dt).reshape(-1)
    N, M = dt.shape[0], kernel.M
    spec = kernel.spec()
    d = M * kernel.state_dim
    fm, fP = as_dev(filter_mean), as_dev(filter_cov)
    if fm.numel() != N * d or fP.numel() != N * d * d:
        raise ValueError('filter_mean must be [N,%d,1] and filter_cov [N,%d,%d] for N = %d' % (d, d, d, N))
    od = d if return_full else M
    means = torch.empty((N, od, 1), dtype=torch.float64, device=dt.device)
    covs = torch.empty((N, od, od), dtype=torch.float64, device=dt.device)
    gains = torch.empty((N, d, d), dtype=torch.float64, device=dt.device) if want_gains else None
    ws, nb = st_workspace(spec, M, N, M)
    _lib.check(_lib.lib().bn_st_rts_smoother(spec, M, N, ptr(dt), ptr(fm), ptr(fP), int(bool(return_full)), ptr(means),
                                             ptr(covs), ptr(gains), ptr(ws), nb, stream_ptr()))
    return means, covs, gains


def st_kalman_filter_meanfield(dt, kernel, y, noise_cov, mask=None, parallel=False, block_index=None, want_ell=True):
    """kalman_filter_meanfield (ops.py:581-611): ell, (means [N,M,n,1], covs [N,M,n,n]); the block structure is
    implied by the kernel, so block_index is accepted and ignored"""
    dt = as_dev(dt).reshape(-1)
    N, M, n = dt.shape[0], kernel.M, kernel.state_dim
    spec = kernel.spec()
    y, R = as_dev(y), as_dev(noise_cov)
    if y.numel() != N * M or R.numel() != N * M * M:
        raise ValueError('y must be [N,%d,1] and noise_cov [N,%d,%d] for N = %d steps' % (M, M, M, N))
    mk = as_mask(mask)
    ell = torch.zeros((), dtype=torch.float64, device=dt.device) if want_ell else None
    means = torch.empty((N, M, n, 1), dtype=torch.float64, device=dt.device)
    covs = torch.empty((N, M, n, n), dtype=torch.float64, device=dt.device)
    ws, nb = st_workspace(spec, M, N, M)
    _lib.check(_lib.lib().bn_st_kalman_filter_meanfield(spec, M, N, ptr(dt), ptr(y), ptr(R), ptr(mk), ptr(ell), ptr(means),
                                                        ptr(covs), ptr(ws), nb, stream_ptr()))
    return ell, (means, covs)


def st_rts_smoother_meanfield(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False, block_index=None,
                              want_gains=True):
    """rauch_tung_striebel_smoother_meanfield (ops.py:681-706): (means, covs, gains); covariances and gains of the full
    state are returned as blocks [N,M,n,n]"""
    dt = as_dev(dt).reshape(-1)
    N, M, n = dt.shape[0], kernel.M, kernel.state_dim
    spec = kernel.spec()
    fm, fP = as_dev(filter_mean), as_dev(filter_cov)
    if fm.numel() != N * M * n or fP.numel() != N * M * n * n:
        raise ValueError('filter_mean must be [N,%d,%d,1] and filter_cov [N,%d,%d,%d]' % (M, n, M, n, n))
    if return_full:
        means = torch.empty((N, M * n, 1), dtype=torch.float64, device=dt.device)
        covs = torch.empty((N, M, n, n), dtype=torch.float64, device=dt.device)
    else:
        means = torch.empty((N, M, 1), dtype=torch.float64, device=dt.device)
        covs = torch.empty((N, M, M), dtype=torch.float64, device=dt.device)
    gains = torch.empty((N, M, n, n), dtype=torch.float64, device=dt.device) if want_gains else None
    _lib.check(_lib.lib().bn_st_rts_smoother_meanfield(spec, M, N, ptr(dt), ptr(fm), ptr(fP), int(bool(return_full)),
                                                       ptr(means), ptr(covs), ptr(gains), stream_ptr()))
    return means, covs, gains


def inv_vmap(P, rhs=None, jitter=0.0, want_logdet=False):
    """utils.py:30-35: batched SPD inverse through the Cholesky factor; optionally P^-1 rhs and log det P"""
    P = as_dev(P)
    N, n = P.shape[0], P.shape[-1]
    out = torch.empty_like(P)
    r = None if rhs is None else as_dev(rhs).reshape(N, n)
    sol = None if rhs is None else torch.empty((N, n, 1), dtype=torch.float64, device=P.device)
    ld = torch.empty((N,), dtype=torch.float64, device=P.device) if want_logdet else None
    spec = _lib.kernel_spec(_lib.BN_MATERN12, [1.0], [1.0])
    ws, nb = st_workspace(spec, n, N, n)
    _lib.check(_lib.lib().bn_spd_inverse_batched(N, n, ptr(P), ptr(r), float(jitter), ptr(out), ptr(sol), ptr(ld), ptr(ws),
                                                 nb, stream_ptr()))
    res = (out,)
    if rhs is not None:
        res += (sol,)
    if want_logdet:
        res += (ld,)
    return res[0] if len(res) == 1 else res


class _DiagSites:
    """GaussianDistribution (basemodels.py:52-100) for sites that are diagonal in data space: the four parameter
    arrays are [N_t, N_s] tensors; `.mean`, `.covariance`, `.nat1`, `.nat2` give the reference's shapes."""

    def __init__(self, N, Ns):
        dev = device()
        self.version = 0
        self.mean_ = torch.zeros((N, Ns), dtype=torch.float64, device=dev)      # mean 0, cov 100 I (basemodels.py:130-133)
        self.covariance_ = torch.full((N, Ns), 1e2, dtype=torch.float64, device=dev)
        self.nat1_ = torch.zeros((N, Ns), dtype=torch.float64, device=dev)
        self.nat2_ = torch.full((N, Ns), 1e-2, dtype=torch.float64, device=dev)

    mean = property(lambda self: self.mean_.unsqueeze(-1))
    nat1 = property(lambda self: self.nat1_.unsqueeze(-1))
    covariance = property(lambda self: torch.diag_embed(self.covariance_))
    nat2 = property(lambda self: torch.diag_embed(self.nat2_))

    def __call__(self):
        return self.mean, self.covariance


class SpatioTemporalMixin:
    """overrides of MarkovGaussianProcess for spatio-temporal inputs; mixed in by MarkovGaussianProcess.__new__"""
    _site_state_dim = 1  # the site pass works on scalar (time, space) observations
    _energy_terms_fused = None  # the KL term lives in the inducing space (M x M blocks): separate kernels

    def __init__(self, kernel, likelihood, X, Y, R=None, parallel=None):
        if getattr(likelihood, 'multi_latent', False):
            raise NotImplementedError('multi-latent likelihoods on the spatio-temporal path')
        if self.method not in (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON):
            # EP / PL form their cavity in the M-dimensional inducing space (B^T Lambda B, basemodels.py:230-245 with
            # compute_full_pseudo_nat) and project it with B (.) B^T + C; the scalar site kernels work in data space
            raise NotImplementedError('the spatio-temporal path supports VI and Newton (the EP / PL cavity in the '
                                      'inducing space is not built)')
        t = np.asarray(X, dtype=np.float64)
        if R is None:  # X = [t, r...] columns (utils.py:251-254) is the flattened-inputs form: not supported here
            raise NotImplementedError('pass the spatial inputs as R [N_t, N_s, n_dims]')
        t = t.reshape(t.shape[0], -1)[:, 0]
        Yh = np.asarray(Y, dtype=np.float64).reshape(t.shape[0], -1)
        Rh = np.asarray(R, dtype=np.float64).reshape(t.shape[0], Yh.shape[1], -1)
        ind = np.argsort(t, kind='stable')
        t, Yh, Rh = t[ind], Yh[ind], Rh[ind]
        if not np.all(np.abs(Rh - Rh[:1]) < 1e-10):
            raise NotImplementedError('the dense path needs the same spatial inputs at every time step (gridded data)')
        dt = np.concatenate([[0.0], np.diff(t)])
        self.kernel, self.likelihood = kernel, likelihood
        self.parallel = False  # the dense path is the sequential recursion
        self.spatio_temporal = True
        self.X, self.Y_host, self.R = t, Yh, Rh
        self.num_data, self.obs_dim = Yh.shape
        self.func_dim = kernel.M
        self.state_dim = kernel.M * kernel.state_dim
        self.Y = as_dev(Yh)
        self.dt = as_dev(dt)
        self.dt_smoother = as_dev(np.concatenate([dt[1:], [0.0]]))
        N, M = self.num_data, kernel.M
        self.pseudo_likelihood = _DiagSites(N, self.obs_dim)
        self.posterior_mean = torch.zeros((N, M, 1), dtype=torch.float64, device=device())
        self.posterior_variance = torch.eye(M, dtype=torch.float64, device=device()).repeat(N, 1, 1)
        mask_y = np.isnan(Yh)
        self.mask_y = mask_y if mask_y.any() else None
        # basemodels.py:136-137 and :652-653: the pseudo observations carry the data mask only when M == N_s
        self.mask_pseudo_y = as_mask(mask_y) if (M == self.obs_dim and mask_y.any()) else None
        self._proj_key = None

    # ---- per-hyper-parameter constants of the spatial conditional -----------------------------------------
    def _hyper_key(self):
        k = self.kernel
        sk = k.spatial_kernel.kernels if isinstance(k.spatial_kernel, Separable) else [k.spatial_kernel]
        return (k.temporal_kernel.variance, k.temporal_kernel.lengthscale) + tuple((s.variance, s.lengthscale) for s in sk)

    def _projection(self):
        key = self._hyper_key()
        if self._proj_key != key:
            B, C = self.kernel.spatial_conditional(self.X, self.R[0])
            self._B = as_dev(B)
            self._Bt = as_dev(np.ascontiguousarray(B.T))
            self._cdiag = as_dev(np.ascontiguousarray(np.diag(C))) if C.shape[0] == B.shape[0] else None
            self._proj_key = key
            self._full_cache = None
        return self._B, self._Bt, self._cdiag

    # ---- basemodels.py:676-687
    def compute_full_pseudo_lik(self):
        """(pseudo_y [N,M,1], pseudo_var [N,M,M]); recomputed only when sites or hyper-parameters changed (the
        reference notes the three evaluations per iteration as wasteful, basemodels.py:677)"""
        B, Bt, _ = self._projection()
        pl = self.pseudo_likelihood
        c = getattr(self, '_full_cache', None)
        if c is not None and c[0] == pl.version:
            return c[1], c[2]
        N, Ns, M = self.num_data, self.obs_dim, self.kernel.M
        py = torch.empty((N, M, 1), dtype=torch.float64, device=B.device)
        pv = torch.empty((N, M, M), dtype=torch.float64, device=B.device)
        ws, nb = st_workspace(self.kernel.spec(), M, N, Ns)
        _lib.check(_lib.lib().bn_st_pseudo_lik(N, Ns, M, ptr(Bt), ptr(pl.nat1_), ptr(pl.nat2_), 1e-12, ptr(py), ptr(pv),
                                               None, None, ptr(ws), nb, stream_ptr()))
        self._full_cache = (pl.version, py, pv)
        return py, pv

    @staticmethod
    def filter(*args, **kwargs):
        return st_kalman_filter(*args, **kwargs)

    @staticmethod
    def smoother(*args, **kwargs):
        return st_rts_smoother(*args, **kwargs)

    # ---- basemodels.py:689-706
    def update_posterior(self, want_grad=False):
        if want_grad:
            raise NotImplementedError('hyper-gradients on the dense spatio-temporal path')
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        ell, (fm, fP) = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y)
        self._ell_cache = (ell, self.pseudo_likelihood.version, self._hyper_key())
        sm, sP, _ = self.smoother(self.dt_smoother, self.kernel, fm, fP, want_gains=False)
        self.posterior_mean, self.posterior_variance = sm, sP

    # ---- basemodels.py:726-741
    def compute_log_lik(self, pseudo_y=None, pseudo_var=None):
        if pseudo_y is None:
            c = getattr(self, '_ell_cache', None)
            if c is not None and c[1] == self.pseudo_likelihood.version and c[2] == self._hyper_key():
                return c[0]
            pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        ell, _ = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y)
        return ell

    def expected_density_pseudo(self):
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        N, M = self.num_data, self.kernel.M
        vals = torch.empty((N,), dtype=torch.float64, device=pseudo_y.device)
        out = torch.zeros((), dtype=torch.float64, device=pseudo_y.device)
        ws, nb = st_workspace(self.kernel.spec(), M, N, self.obs_dim)
        _lib.check(_lib.lib().bn_st_gaussian_expected_log_lik(
            N, M, ptr(pseudo_y), ptr(self.posterior_mean), ptr(self.posterior_variance), ptr(pseudo_var),
            ptr(self.mask_pseudo_y), ptr(vals), ptr(out), ptr(ws), nb, stream_ptr()))
        return out

    # ---- basemodels.py:743-764 (marginals: the factorising likelihoods read diag(cov_f), likelihoods.py:371-373)
    def conditional_posterior_to_data(self, batch_ind=None, post_mean=None, post_cov=None):
        B, _, cdiag = self._projection()
        pm = self.posterior_mean if post_mean is None else as_dev(post_mean)
        pV = self.posterior_variance if post_cov is None else as_dev(post_cov)
        N, Ns, M = pm.shape[0], self.obs_dim, self.kernel.M
        mean_f = torch.empty((N, Ns, 1), dtype=torch.float64, device=B.device)
        var_f = torch.empty((N, Ns, 1), dtype=torch.float64, device=B.device)
        _lib.check(_lib.lib().bn_st_posterior_to_data(N, Ns, M, ptr(B), ptr(cdiag), ptr(pm), ptr(pV), ptr(mean_f),
                                                      ptr(var_f), stream_ptr()))
        return mean_f, var_f

    # ---- the site pass: every (time, space) observation is one scalar site (inference.py:170-195)
    def _site_args(self, cubature=None):
        mean_f, var_f = self.conditional_posterior_to_data()
        a, keep = self.likelihood.site_args(self.method, self.Y.reshape(-1), mean_f.reshape(-1), var_f.reshape(-1),
                                            cubature, self.power)
        pl = self.pseudo_likelihood
        a.nat1, a.nat2 = pl.nat1_.data_ptr(), pl.nat2_.data_ptr()
        return a, keep + [mean_f, var_f]

    def predict(self, X=None, R=None, pseudo_lik_params=None):
        """latent mean and variance at test times X [N*] and spatial inputs R [N_s*, n_dims] (the same at every test
        time; default: the training ones): filter, full-state smoother with gains, the temporal conditional on the
        Kronecker state (bn_st_predict_state) and the spatial conditional (basemodels.py:766-816).  Returns
        (mean [N*, N_s*], var [N*, N_s*]) -- the reference discards the spatial covariance as well (:813-815)."""
        times = self.X if X is None else np.asarray(X, dtype=np.float64).reshape(-1)
        Rs = self.R[0] if R is None else np.asarray(R, dtype=np.float64)
        if Rs.ndim == 3:
            if not np.all(np.abs(Rs - Rs[:1]) < 1e-10):
                raise NotImplementedError('the dense path predicts on one set of spatial inputs for all test times')
            Rs = Rs[0]
        B, C = self.kernel.spatial_conditional(times, Rs, predict=True)
        Bd = as_dev(B)
        cdiag = as_dev(np.ascontiguousarray(np.diag(C))) if C.shape[0] == B.shape[0] else None
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik() if pseudo_lik_params is None else pseudo_lik_params
        _, (fm, fP) = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y, want_ell=False)
        sm, sP, gain = self.smoother(self.dt_smoother, self.kernel, fm, fP, return_full=True)
        return self._predict_from_state(times, sm, sP, gain, Bd, cdiag)

    def _predict_from_state(self, times, sm, sP, gain, Bd, cdiag):
        M, Nq, Ns = self.kernel.M, times.shape[0], Bd.shape[0]
        spec = self.kernel.spec()
        xs = as_dev(times)
        fmean = torch.empty((Nq, M, 1), dtype=torch.float64, device=xs.device)
        fcov = torch.empty((Nq, M, M), dtype=torch.float64, device=xs.device)
        need = int(_lib.lib().bn_st_predict_workspace_bytes(spec, M, Nq))
        ws = torch.empty(need, dtype=torch.uint8, device=xs.device)
        _lib.check(_lib.lib().bn_st_predict_state(spec, M, self.num_data, ptr(as_dev(self.X)), Nq, ptr(xs), ptr(sm), ptr(sP),
                                                  ptr(gain), ptr(fmean), ptr(fcov), ptr(ws), need, stream_ptr()))
        mean = torch.empty((Nq, Ns), dtype=torch.float64, device=xs.device)
        var = torch.empty((Nq, Ns), dtype=torch.float64, device=xs.device)
        _lib.check(_lib.lib().bn_st_posterior_to_data(Nq, Ns, M, ptr(Bd), ptr(cdiag), ptr(fmean), ptr(fcov), ptr(mean), ptr(var),
                                                      stream_ptr()))
        return mean.squeeze(), var.squeeze()


class MeanFieldMixin(SpatioTemporalMixin):
    """MarkovMeanFieldGaussianProcess (basemodels.py:1155-1175): the spatio-temporal model with the mean-field filter
    and smoother (independent temporal blocks, coupled only through the M x M innovation)"""

    @staticmethod
    def filter(*args, **kwargs):
        return st_kalman_filter_meanfield(*args, **kwargs)

    @staticmethod
    def smoother(*args, **kwargs):
        return st_rts_smoother_meanfield(*args, **kwargs)
