"""Host orchestration of a temporal Markov GP (mirror of bayesnewton/basemodels.py:52-100,103-262,625-764).

State (sites in both parametrisations, posterior marginals) lives in float64 CUDA tensors; every
O(N) operation is one libbn_b200 call.  Only the temporal case is covered here (spatio-temporal
projection is a `next` row of SURVEY section 8f).
"""
import numpy as np
import torch

import os

from . import _lib, fused, ops
from ._util import as_dev, as_mask, device, ptr, stream_ptr, workspace
from .kernels import Independent


def input_admin(t, y):
    """sort by time, float64, dt = [0, diff(t)]  (utils.py:234-265, temporal inputs)"""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(t.shape[0], -1)
    ind = np.argsort(t, kind='stable')
    t, y = t[ind], y[ind]
    dt = np.concatenate([[0.0], np.diff(t)])
    return t, y, dt


class GaussianDistribution:
    """sites in (mean, covariance) and natural (nat1, nat2 = cov^-1) form (basemodels.py:52-100).

    While a fused iteration (fused.FusedShard) owns the sites they live in its tiled arrays; `source` then names it and
    the four arrays below are materialised on first access (diagonal sites: nat2 = 1 / cov, nat1 = nat2 mean)."""

    def __init__(self, mean, covariance, nat1=None, nat2=None):
        self.version = 0  # bumped whenever the parameters change (invalidates anything derived from them)
        self.source = None
        self._mean, self._cov = as_dev(mean), as_dev(covariance)
        if nat1 is None:
            self._nat1, self._nat2 = self.reparametrise(self._mean, self._cov)
        else:  # the caller knows both forms (e.g. the diagonal initial sites): no factorisation needed
            self._nat1, self._nat2 = as_dev(nat1), as_dev(nat2)

    def __call__(self):
        return self.mean, self.covariance

    def _sync(self):
        src, self.source = self.source, None
        if src is not None:
            self._mean, self._cov = (t.double() for t in src.sites())  # (an fp32 shard hands back float32)
            self._nat2 = 1.0 / self._cov
            self._nat1 = self._mean * self._nat2

    def _get(name):  # noqa: N805 -- property factory
        def getter(self):
            self._sync()
            return getattr(self, name)

        def setter(self, value):
            self._sync()
            setattr(self, name, value)
        return property(getter, setter)

    mean_, covariance_, nat1_, nat2_ = _get('_mean'), _get('_cov'), _get('_nat1'), _get('_nat2')
    mean, covariance, nat1, nat2 = mean_, covariance_, nat1_, nat2_
    del _get

    @staticmethod
    def reparametrise(param1, param2):
        # only used at construction and by update_mean_cov (never inside an inference iteration,
        # where the fused site kernel reparametrises in registers): batched Cholesky solve
        chol = torch.linalg.cholesky(param2)
        eye = torch.eye(param2.shape[-1], dtype=param2.dtype, device=param2.device).expand_as(param2)
        return torch.cholesky_solve(param1, chol), torch.cholesky_solve(eye, chol)

    def update_mean_cov(self, mean, covariance):
        self.version += 1
        self.source = None
        self._mean, self._cov = as_dev(mean), as_dev(covariance)
        self._nat1, self._nat2 = self.reparametrise(self._mean, self._cov)


class MarkovGaussianProcess:
    """f ~ GP in SDE form; inference by Kalman filtering / RTS smoothing (basemodels.py:625-764)"""
    method = None   # BN_METHOD_*, set by the inference mixin
    power = 1.0
    _st_mixin = 'SpatioTemporalMixin'

    def __new__(cls, kernel=None, *args, **kwargs):
        # spatio-temporal inputs take the dense path: the overrides of spacetime.SpatioTemporalMixin go in front
        if hasattr(kernel, 'temporal_kernel'):
            from . import spacetime
            mixin = getattr(spacetime, cls._st_mixin)
            if not issubclass(cls, mixin):
                cls = type(cls.__name__, (mixin, cls), {})
        elif cls._st_mixin != 'SpatioTemporalMixin':
            raise NotImplementedError('the mean-field model is built for spatio-temporal kernels')
        return object.__new__(cls)

    def __init__(self, kernel, likelihood, X, Y, R=None, parallel=None):
        if R is not None:
            raise NotImplementedError('spatial inputs need a SpatioTemporalKernel')
        if parallel is None:  # the reference switches the scan on when it runs on a GPU (basemodels.py:642-643)
            parallel = True
        self.kernel, self.likelihood, self.parallel = kernel, likelihood, parallel
        t, Yh, dt = input_admin(X, Y)
        self.num_data = t.shape[0]
        self.X, self.Y_host = t, Yh
        self.Y = as_dev(Yh)
        self.dt = as_dev(dt)
        self.dt_smoother = as_dev(np.concatenate([dt[1:], [0.0]]))
        H = kernel.measurement_model()
        self.func_dim = H.shape[0]
        self.obs_dim = Yh.shape[1]
        self.state_dim = kernel.stationary_covariance().shape[0]
        D = self.func_dim if isinstance(kernel, Independent) else self.obs_dim
        if D != self.func_dim or self.obs_dim != 1:
            raise NotImplementedError('one observation per step and one site per latent are supported')
        N = self.num_data
        eye = torch.eye(D, dtype=torch.float64, device=device()).repeat(N, 1, 1)
        self.pseudo_likelihood = GaussianDistribution(  # mean 0, cov 100 I (basemodels.py:130-133)
            mean=torch.zeros((N, D, 1), dtype=torch.float64, device=device()), covariance=1e2 * eye,
            nat1=torch.zeros((N, D, 1), dtype=torch.float64, device=device()), nat2=1e-2 * eye)
        self.posterior_mean = torch.zeros((N, D, 1), dtype=torch.float64, device=device())
        self.posterior_variance = torch.eye(D, dtype=torch.float64, device=device()).repeat(N, 1, 1)
        mask_y = np.isnan(Yh)
        self.mask_y = mask_y if mask_y.any() else None
        if D == self.obs_dim:
            self.mask_pseudo_y = as_mask(mask_y) if mask_y.any() else None
        else:  # multi-latent likelihood: no mask on the sites (basemodels.py:141-142)
            self.mask_pseudo_y = None

    # ---- basemodels.py:655-661
    @staticmethod
    def filter(*args, **kwargs):
        return ops.kalman_filter(*args, **kwargs)

    @staticmethod
    def smoother(*args, **kwargs):
        return ops.rauch_tung_striebel_smoother(*args, **kwargs)

    def compute_full_pseudo_lik(self):
        return self.pseudo_likelihood.mean, self.pseudo_likelihood.covariance

    # ---- fused iteration on tiled resident state (fused.py; csrc/iter_impl.cuh)
    def _fused_ok(self):
        """the whole iteration can run as two fused passes: scan form, one in-library Matern component, a single-latent
        likelihood of the site kernels, VI or Newton (EP with BN_B200_FUSED_EP=1), and none of the hooks overridden (spatio-temporal mixins)"""
        if os.environ.get('BN_B200_FUSED', '1') == '0' or not self.parallel or self.func_dim != 1:
            return False
        if type(self).update_posterior is not MarkovGaussianProcess.update_posterior:
            return False
        spec = self.kernel.spec() if hasattr(self.kernel, 'spec') else None
        return fused.supported(spec, self.likelihood, self.method)

    def _fused_state(self):
        """the FusedShard of this model, holding the current sites"""
        st = getattr(self, '_fused', None)
        pl = self.pseudo_likelihood
        if st is None:
            st = self._fused = fused.FusedShard(self.kernel, self.dt, self.Y, self.mask_pseudo_y,
                                                dtype=getattr(self, '_fused_dtype', torch.float64))
            st.sites_version = None
        if pl.source is not st or st.sites_version != pl.version:
            st.load_sites(pl.mean, pl.covariance)
            st.sites_version = pl.version
        return st

    def set_precision(self, precision):
        """'float32' runs the fused iteration in the fp32 build of its kernels (bn_iter_*_f32: storage and arithmetic in
        fp32, sums in fp64; parity bar 1e-4 against fp64); posterior_mean / posterior_variance become float32 tensors, the
        sites and everything outside the fused iteration (prediction, the stage-level entries) stay float64"""
        dtype = {'float64': torch.float64, 'float32': torch.float32}[precision]
        if dtype == torch.float32 and not self._fused_ok():
            raise NotImplementedError('the fp32 build covers the fused iteration (scan form, one Matern component, a '
                                      'single-latent likelihood, VI or Newton)')
        self.pseudo_likelihood.mean  # noqa: B018 -- materialise the sites before the resident state is rebuilt
        self._fused_dtype = dtype
        self._fused = None
        self.posterior_mean, self.posterior_variance = self.posterior_mean.to(dtype), self.posterior_variance.to(dtype)
        self._ell_cache = self._grad_cache = self._energy_cache = None

    def load_inputs(self, dt, Y):
        """replace the step lengths and observations (device or pinned-host tensors of the model's length, already in
        time order; Y may hold uint8 / bool labels) -- the streaming entry the end-to-end benchmark drives"""
        dt = as_dev(dt).reshape(-1)
        labels = torch.is_tensor(Y) and not Y.dtype.is_floating_point  # integer labels cannot hold a missing value
        Y = as_dev(Y).reshape(-1, 1)
        if dt.shape[0] != self.num_data or Y.shape[0] != self.num_data:
            raise ValueError('load_inputs needs %d steps' % self.num_data)
        self.dt, self.Y = dt, Y
        self.dt_smoother = torch.cat([dt[1:], torch.zeros(1, dtype=dt.dtype, device=dt.device)])
        nan = None if labels else torch.isnan(Y)
        self.mask_pseudo_y = nan.to(torch.uint8).reshape(-1, 1, 1).contiguous() if (nan is not None and bool(nan.any())) else None
        self._ell_cache = self._grad_cache = self._energy_cache = None
        st = getattr(self, '_fused', None)
        if st is not None:
            st.set_dt(dt)
            st.set_data(Y, self.mask_pseudo_y, scan_nan=not labels)

    def _energy_key(self, cubature):
        return (self.pseudo_likelihood.version, self._hyper_key(), float(self.likelihood.lik_param), float(self.likelihood.lik_param2),
                fused.cubature_key(cubature), self.method, float(getattr(self, 'power', 1.0)))

    def _hyper_key(self):
        spec = self.kernel.spec() if hasattr(self.kernel, 'spec') else None
        if spec is None:
            return None
        return (spec.family, spec.n_components, tuple(spec.variance), tuple(spec.lengthscale))

    def update_posterior(self, want_grad=False):
        """filter then smoother (basemodels.py:689-706).  want_grad: the same pass also accumulates
        d log-lik / d kernel hyper-parameters, kept for energy_and_grad().  With the scan form and an in-library kernel this is
        ONE fused call (ops.update_posterior); the filter log-likelihood it produces on the way is kept and
        served to compute_log_lik() for as long as the sites and hyper-parameters it was computed from stand
        (the reference evaluates the identical filter a second time inside energy(), basemodels.py:733)."""
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        self._grad_cache = None
        if want_grad and self._hyper_key() is None:
            raise NotImplementedError('the hyper-gradient needs a kernel with an in-library discretisation (kernel.spec())')
        # the gradient pass always runs the fused scan-form update: for parallel=False models too (same posterior and
        # gradient to rounding; the sequential form has no adjoint kernel of its own)
        if (self.parallel or want_grad) and self._hyper_key() is not None:
            out = ops.update_posterior(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y,
                                       want_ell=True, want_grad=want_grad)
            ell, sm, sP = out[:3]
            self._ell_cache = (ell, self.pseudo_likelihood.version, self._hyper_key())
            if want_grad:
                self._grad_cache = (out[3], self.pseudo_likelihood.version, self._hyper_key())
        else:
            _, (fm, fP) = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y,
                                      parallel=self.parallel, want_ell=False)
            sm, sP, _ = self.smoother(self.dt_smoother, self.kernel, fm, fP, parallel=self.parallel,
                                      want_gains=False)
        self.posterior_mean, self.posterior_variance = sm, sP

    def compute_log_lik(self, pseudo_y=None, pseudo_var=None):
        """log normaliser of the pseudo model = the filter's log-likelihood (basemodels.py:726-741)"""
        if pseudo_y is None:
            cache = getattr(self, '_ell_cache', None)
            if cache is not None and cache[1] == self.pseudo_likelihood.version and cache[2] == self._hyper_key():
                return cache[0]
            pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        ell, _ = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y,
                             parallel=self.parallel, want_states=False)
        return ell

    def log_lik_grad(self):
        """d compute_log_lik() / d [variance_c...; lengthscale_c...] ([2, NC], untransformed hyper-parameters);
        served from the last update_posterior(want_grad=True) while sites and hyper-parameters stand"""
        cache = getattr(self, '_grad_cache', None)
        if cache is None or cache[1] != self.pseudo_likelihood.version or cache[2] != self._hyper_key():
            self.update_posterior(want_grad=True)
        return self._grad_cache[0]

    def expected_density_pseudo(self):
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        N, D = pseudo_y.shape[0], pseudo_y.shape[1]
        out = torch.zeros((), dtype=torch.float64, device=pseudo_y.device)
        ws, nb = workspace(N, self.state_dim, D)
        pm, pv = as_dev(self.posterior_mean), as_dev(self.posterior_variance)  # (float32 after set_precision('float32'))
        _lib.check(_lib.lib().bn_gaussian_expected_log_lik(
            N, D, ptr(pseudo_y), ptr(pm), ptr(pv), ptr(pseudo_var),
            ptr(self.mask_pseudo_y), None, ptr(out), ptr(ws), nb, stream_ptr()))
        return out

    def _energy_terms_fused(self, cubature=None):
        """(sum_n likelihood term, sum_n E_q[log N(pseudo_y_n | f_n, pseudo_var_n)]) in one kernel for single-latent
        VI / Newton models; None when the model has more than one latent"""
        if self.func_dim != 1 or self.method not in (_lib.BN_METHOD_VI, _lib.BN_METHOD_NEWTON):
            return None
        cache = getattr(self, '_energy_cache', None)
        if cache is not None and cache[2] == self._energy_key(cubature):  # the closing pass of the fused inference()
            return cache[0], cache[1]
        a, keep = self._site_args(cubature)
        pl = self.pseudo_likelihood
        a.site_mean, a.site_cov = pl.mean.data_ptr(), pl.covariance.data_ptr()
        parts = torch.zeros(2, dtype=torch.float64, device=self.posterior_mean.device)
        ws, nb = workspace(a.N, self.state_dim, 1)
        _lib.check(_lib.lib().bn_energy_terms(a, ptr(self.mask_pseudo_y), parts.data_ptr(), ptr(ws), nb, stream_ptr()))
        return parts[0], parts[1]

    def compute_kl(self):
        """KL[q || p] = sum_n E_q[log N(pseudo_y_n | f_n, pseudo_var_n)] - log Z_pseudo  (basemodels.py:708-724)"""
        return self.expected_density_pseudo() - self.compute_log_lik()

    def conditional_posterior_to_data(self, batch_ind=None, post_mean=None, post_cov=None):
        return (self.posterior_mean if post_mean is None else post_mean,
                self.posterior_variance if post_cov is None else post_cov)

    @staticmethod
    def temporal_conditional(*args, **kwargs):
        return ops.temporal_conditional(*args, **kwargs)

    def predict(self, X=None, R=None, pseudo_lik_params=None):
        """posterior of the latent function(s) at test inputs X (basemodels.py:766-816): filter, full-state smoother
        with gains, then the two-sided conditional on the neighbouring training states (utils.py:99-215) and the
        measurement model.  Returns (mean, var) squeezed like the reference: [N*] each for one latent."""
        if R is not None:
            raise NotImplementedError('spatial test inputs need the spatio-temporal model')
        X = self.X if X is None else np.asarray(X, dtype=np.float64).reshape(-1)
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik() if pseudo_lik_params is None else pseudo_lik_params
        _, (fm, fP) = self.filter(self.dt, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y,
                                  parallel=self.parallel, want_ell=False)
        sm, sP, gain = self.smoother(self.dt_smoother, self.kernel, fm, fP, return_full=True, parallel=self.parallel)
        test_mean, test_var = self.temporal_conditional(self.X, X, sm, sP, gain, self.kernel, return_full=False)
        return test_mean.squeeze(), test_var.squeeze()

    def predict_y(self, X, R=None, cubature=None):
        """predictive mean and variance of the observations at X (basemodels.py:165-175)"""
        if getattr(self.likelihood, 'multi_latent', False):
            raise NotImplementedError('predict_y for multi-latent likelihoods')
        mean_f, var_f = self.predict(X, R)
        return self.likelihood.predict(mean_f.reshape(-1), var_f.reshape(-1), cubature)

    def negative_log_predictive_density(self, X, Y, R=None, cubature=None):
        """-nanmean_n log E_q[p(y_n | f_n)] at test inputs (basemodels.py:177-193), single-latent likelihoods"""
        if getattr(self.likelihood, 'multi_latent', False):
            raise NotImplementedError('negative_log_predictive_density for multi-latent likelihoods')
        mean_f, var_f = self.predict(X, R)
        Yt = as_dev(np.asarray(Y, dtype=np.float64).reshape(-1))
        ld = self.likelihood.log_density(Yt, mean_f.reshape(-1), var_f.reshape(-1), cubature)
        # a missing test target yields NaN in the reference's log_density_cubature and is dropped by nanmean; the raw EP
        # kernel substitutes y := m for it (the moment_match rule), so the NaN is restored here
        ld = torch.where(torch.isnan(Yt), torch.full_like(ld, float('nan')), ld)
        return -torch.nanmean(ld)


class MarkovMeanFieldGaussianProcess(MarkovGaussianProcess):
    """basemodels.py:1155-1175: mean-field across the spatial (latent) blocks"""
    _st_mixin = 'MeanFieldMixin'


MarkovMeanFieldGP = MarkovMeanFieldGaussianProcess


class InfiniteHorizonGaussianProcess(MarkovGaussianProcess):
    """steady-state Markov GP (basemodels.py:1257-1300): the state covariance is the fixed point of the Riccati recursion
    for the AVERAGED site precision -- 20 iterations per filter call, warm-started from the previous call's result
    (:1272-1289) -- so filter and smoother are affine recursions in the mean (ops.kalman_filter_infinite_horizon).
    Evenly spaced inputs, one latent; the spatio-temporal and sparse variants of the reference are not built."""

    def __init__(self, kernel, likelihood, X, Y, R=None, dare_iters=20, parallel=None):
        from .likelihoods import Gaussian
        super().__init__(kernel, likelihood, X, Y, R=R, parallel=parallel)
        if self.func_dim != 1 or self.state_dim > 4:
            raise NotImplementedError('the infinite-horizon model is built for one latent with state dimension <= 4')
        if self.num_data > 2 and np.max(np.abs(np.diff(np.diff(self.X)))) >= 1e-6:
            raise AssertionError('the infinite-horizon model needs equidistant time steps (basemodels.py:1269)')
        self.heteroscedastic = bool(np.isnan(self.Y_host).any()) or not isinstance(likelihood, Gaussian)
        self.dare_iters = dare_iters
        Pinf = np.asarray(kernel.stationary_covariance(), dtype=np.float64)
        self.dare_init_filter, self.dare_init_smoother = Pinf, Pinf
        self._dt_host = np.concatenate([[0.0], np.diff(self.X)])

    def _fused_ok(self):
        return False

    def filter(self, dt, kernel, y, noise_cov, mask=None, parallel=False, want_ell=True, **kw):
        tied = 1.0 / float(self.pseudo_likelihood.nat2.mean())  # inv(mean of the site precisions), basemodels.py:1280-1281
        out = ops.kalman_filter_infinite_horizon(dt, kernel, y, noise_cov, mask, parallel=parallel,
                                                 heteroscedastic=self.heteroscedastic, noise_cov_tied=tied,
                                                 dare_iters=self.dare_iters, dare_init=self.dare_init_filter,
                                                 want_ell=want_ell)
        self.dare_init_filter = out[1][1][0]
        return out

    def smoother(self, dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False, **kw):
        means, covs, gains, dare_cov = ops.rauch_tung_striebel_smoother_infinite_horizon(
            dt, kernel, filter_mean, filter_cov, return_full=return_full, parallel=parallel, dare_iters=self.dare_iters,
            dare_init=self.dare_init_smoother)
        self.dare_init_smoother = dare_cov
        return means, covs, gains

    def update_posterior(self, want_grad=False):
        if want_grad:
            raise NotImplementedError('no hyper-gradient pass for the infinite-horizon model')
        pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        # the smoother's transition is the step OUT of a state: on the even grid that is dt[1] everywhere (only element 0 is
        # read; building the shifted series on the host cost 15 ms per update at N = 1e7)
        dts = self._dt_host[1:2] if self.num_data > 1 else np.zeros(1)
        _, (fm, fcov) = self.filter(self._dt_host, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y,
                                    parallel=self.parallel, want_ell=False)
        self.posterior_mean, self.posterior_variance, _ = self.smoother(dts, self.kernel, fm, fcov, parallel=self.parallel)

    def compute_log_lik(self, pseudo_y=None, pseudo_var=None):
        """every call runs the filter again: the Riccati warm start makes the result depend on the number of calls, and the
        reference's sequence of calls is mirrored (no caching)"""
        if pseudo_y is None:
            pseudo_y, pseudo_var = self.compute_full_pseudo_lik()
        ell, _ = self.filter(self._dt_host, self.kernel, pseudo_y, pseudo_var, mask=self.mask_pseudo_y, parallel=self.parallel)
        return ell


IHGP = InfiniteHorizonGaussianProcess
