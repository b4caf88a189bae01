"""Sparse Markov GP: host mirror of SparseMarkovGaussianProcess (bayesnewton/basemodels.py:928-1152) and of
kalman_filter_pairs (ops.py:383-426), executed by libbn_b200 (csrc/sparse.cu + the array-level filter).

The model keeps one Gaussian site per TRANSITION between neighbouring inducing inputs (2n-dimensional: the pair of
states [u_-, u_+]); data points reach their transition through the two-sided conditional of
compute_conditional_statistics (utils.py:173-215).  Full-batch VI with a single-latent likelihood is what is built
(the reference's SparseMarkovVariationalGP, models.py); the O(N) data pass is one fused kernel
(bn_sparse_site_update), everything else is O(number of inducing points).
"""
import numpy as np
import torch

from . import _lib, ops
from ._util import as_dev, device, ptr, stream_ptr, workspace
from .basemodels import input_admin
from .cubature import host_table


def kalman_filter_pairs(dt, kernel, y, noise_cov, mask=None, parallel=False, want_ell=True):
    """Kalman filter over pairs of states (ops.py:383-426): y [N,2n,1], noise_cov [N,2n,2n] ->
    ell, (means [N-1,n,1], covs [N-1,n,n])"""
    spec = kernel.spec() if hasattr(kernel, 'spec') else None
    if spec is None or spec.n_components != 1:
        raise NotImplementedError('the pairs filter needs a single-component kernel with kernel.spec()')
    dt = as_dev(dt).reshape(-1)
    Mt = dt.shape[0]
    n = _lib.lib().bn_state_dim(spec)
    p = 2 * n
    Ap = torch.empty((Mt, p, p), dtype=torch.float64, device=dt.device)
    Qp = torch.empty((Mt, p, p), dtype=torch.float64, device=dt.device)
    _lib.check(_lib.lib().bn_pairs_discretise(spec, Mt, ptr(dt), ptr(Ap), ptr(Qp), stream_ptr()))
    Pinf = as_dev(kernel.stationary_covariance())
    Pinfpair = torch.zeros((p, p), dtype=torch.float64, device=dt.device)
    Pinfpair[:n, :n] = Pinf
    Pinfpair[n:, n:] = Pinf
    minfpair = torch.zeros((p, 1), dtype=torch.float64, device=dt.device)
    H = torch.eye(p, dtype=torch.float64, device=dt.device)
    ell, means, covs = ops._kf_arrays(ops._form(parallel), Ap, Qp, H, y, noise_cov, minfpair, Pinfpair, mask, False,
                                      want_ell=want_ell)
    return ell, (means[1:, :n].contiguous(), covs[1:, :n, :n].contiguous())


def _sparse_workspace(Mz):
    need = int(_lib.lib().bn_sparse_workspace_bytes(int(Mz)))
    ws = torch.empty(need, dtype=torch.uint8, device=device())
    return ws, need


class _PairSites:
    """GaussianDistribution (basemodels.py:52-100) over the transitions; the fused site kernel rewrites all four arrays"""

    def __init__(self, mean, cov, nat1, nat2):
        self.version = 0
        self.mean_, self.covariance_, self.nat1_, self.nat2_ = mean, cov, nat1, nat2

    mean = property(lambda self: self.mean_)
    covariance = property(lambda self: self.covariance_)
    nat1 = property(lambda self: self.nat1_)
    nat2 = property(lambda self: self.nat2_)

    def __call__(self):
        return self.mean, self.covariance


class SparseMarkovGaussianProcess:
    """basemodels.py:928-1152 (temporal inputs, full batch)"""
    method = _lib.BN_METHOD_VI
    power = 1.0

    def __init__(self, kernel, likelihood, X, Y, Z, R=None, parallel=None):
        if R is not None:
            raise NotImplementedError('spatio-temporal sparse Markov models are not on this path')
        if getattr(likelihood, 'multi_latent', False):
            raise NotImplementedError('multi-latent likelihoods on the sparse Markov path')
        spec = kernel.spec() if hasattr(kernel, 'spec') else None
        if spec is None or spec.n_components != 1 or spec.family == _lib.BN_MATERN72:
            raise NotImplementedError('single-component Matern-1/2, -3/2, -5/2 kernels are supported')
        self.kernel, self.likelihood = kernel, likelihood
        self.parallel = True if parallel is None else parallel
        t, Yh, _ = input_admin(X, Y)
        if Yh.shape[1] != 1:
            raise NotImplementedError('one observation per input')
        self.num_data = t.shape[0]
        self.X, self.Y_host = t, Yh
        self.x_dev, self.Y = as_dev(t), as_dev(Yh[:, 0])
        self.state_dim = n = kernel.state_dim
        Zs = np.sort(np.asarray(Z, dtype=np.float64).reshape(-1))
        self.Z = np.concatenate([[-1e10], Zs, [1e10]])          # basemodels.py:943-945
        self.z_dev = as_dev(Zs)
        self.dz_host = np.diff(self.Z)
        self.dz = as_dev(self.dz_host)
        self.dz_smoother = as_dev(self.dz_host[1:])
        self.num_transitions = Mt = self.dz_host.shape[0]
        self.Mz = Zs.shape[0]
        # set_z_stats (utils.py:556-559): transition of every data point; sorted inputs => contiguous ranges
        self.ind = np.searchsorted(self.Z, t) - 1
        self.num_neighbours = np.bincount(self.ind, minlength=Mt)[:Mt]
        start = np.searchsorted(self.ind, np.arange(Mt + 1), side='left').astype(np.int64)
        self.start = torch.as_tensor(start, device=device())
        dev = device()
        p = 2 * n
        nat2 = 1e-8 * torch.eye(p, dtype=torch.float64, device=dev).repeat(Mt, 1, 1)
        nat2[:-1, n, n] = 1e-2                                   # basemodels.py:954
        cov = torch.diag_embed(1.0 / torch.diagonal(nat2, dim1=1, dim2=2))   # inv_vmap of a diagonal matrix
        zeros = torch.zeros((Mt, p, 1), dtype=torch.float64, device=dev)
        self.pseudo_likelihood = _PairSites(zeros.clone(), cov, zeros.clone(), nat2)
        self.posterior_mean = zeros.clone()
        self.posterior_variance = torch.eye(p, dtype=torch.float64, device=dev).repeat(Mt, 1, 1)
        self.mask_pseudo_y = None
        self.func_dim, self.obs_dim = 1, 1

    @staticmethod
    def filter(*args, **kwargs):
        return kalman_filter_pairs(*args, **kwargs)

    @staticmethod
    def smoother(*args, **kwargs):
        return ops.rauch_tung_striebel_smoother(*args, **kwargs)

    def compute_full_pseudo_lik(self):
        return self.pseudo_likelihood()

    def _hyper_key(self):
        return (self.kernel.variance, self.kernel.lengthscale)

    def _smoothed_states(self):
        pl = self.pseudo_likelihood
        ell, (fm, fP) = self.filter(self.dz, self.kernel, pl.mean, pl.covariance, parallel=self.parallel)
        self._ell_cache = (ell, pl.version, self._hyper_key())
        return self.smoother(self.dz_smoother, self.kernel, fm, fP, return_full=True, parallel=self.parallel)

    def update_posterior(self):
        """pairs filter, smoother, joint of neighbouring states (basemodels.py:980-1008)"""
        sm, sP, gain = self._smoothed_states()
        Mt, p = self.num_transitions, 2 * self.state_dim
        pm = torch.empty((Mt, p, 1), dtype=torch.float64, device=sm.device)
        pV = torch.empty((Mt, p, p), dtype=torch.float64, device=sm.device)
        _lib.check(_lib.lib().bn_build_joint(self.kernel.spec(), Mt, ptr(sm), ptr(sP), ptr(gain), ptr(pm), ptr(pV),
                                             stream_ptr()))
        self.posterior_mean, self.posterior_variance = pm, pV

    def compute_log_lik(self, pseudo_y=None, pseudo_var=None):
        c = getattr(self, '_ell_cache', None)
        if c is not None and c[1] == self.pseudo_likelihood.version and c[2] == self._hyper_key():
            return c[0]
        pl = self.pseudo_likelihood
        ell, _ = self.filter(self.dz, self.kernel, pl.mean, pl.covariance, parallel=self.parallel)
        return ell

    def expected_density_pseudo(self):
        """sum over transitions of gaussian_expected_log_lik with full 2n x 2n blocks (basemodels.py:215-223)"""
        pl = self.pseudo_likelihood
        Mt, p = self.num_transitions, 2 * self.state_dim
        out = torch.zeros((), dtype=torch.float64, device=pl.mean.device)
        ws, nb = workspace(Mt, 1, 1)
        _lib.check(_lib.lib().bn_gaussian_expected_log_lik(Mt, p, ptr(pl.mean), ptr(self.posterior_mean),
                                                           ptr(self.posterior_variance), ptr(pl.covariance), None, None,
                                                           ptr(out), ptr(ws), nb, stream_ptr()))
        return out

    def compute_kl(self):
        return self.expected_density_pseudo() - self.compute_log_lik()

    def _cub(self, cubature):
        if self.likelihood.lik_id in (_lib.BN_LIK_GAUSSIAN, _lib.BN_LIK_POISSON_EXP):  # closed-form VI statistics
            return 0, None, None
        cx, cw, Q = host_table(cubature, 1)
        return Q, cx, cw

    def inference(self, lr=1., batch_ind=None, cubature=None, ensure_psd=True, **kwargs):
        """one VI iteration (inference.py:65-90, 170-195 with the sparse overrides basemodels.py:1071-1138)"""
        if batch_ind is not None and len(batch_ind) != self.num_data:
            raise NotImplementedError('mini-batched site updates are outside the hot-path scope')
        self.update_posterior()
        pl = self.pseudo_likelihood
        Q, cx, cw = self._cub(cubature)
        diffs = torch.zeros((2,), dtype=torch.float64, device=pl.mean.device)
        ws, nb = _sparse_workspace(self.Mz)
        lik = self.likelihood
        _lib.check(_lib.lib().bn_sparse_site_update(
            self.kernel.spec(), lik.lik_id, float(lik.lik_param), self.num_data, self.Mz, ptr(self.x_dev), ptr(self.Y),
            ptr(self.z_dev), ptr(self.start), ptr(self.posterior_mean), ptr(self.posterior_variance), Q,
            None if cx is None else cx.ctypes.data, None if cw is None else cw.ctypes.data, float(lr), int(bool(ensure_psd)),
            ptr(pl.nat1_), ptr(pl.nat2_), ptr(pl.mean_), ptr(pl.covariance_), ptr(diffs), ptr(ws), nb, stream_ptr()))
        pl.version += 1
        self.update_posterior()
        return None, (diffs[0], diffs[1])

    def expected_density(self, cubature=None):
        Q, cx, cw = self._cub(cubature)
        out = torch.zeros((), dtype=torch.float64, device=self.posterior_mean.device)
        ws, nb = _sparse_workspace(self.Mz)
        lik = self.likelihood
        _lib.check(_lib.lib().bn_sparse_expected_density(
            self.kernel.spec(), lik.lik_id, float(lik.lik_param), self.num_data, self.Mz, ptr(self.x_dev), ptr(self.Y),
            ptr(self.z_dev), ptr(self.start), ptr(self.posterior_mean), ptr(self.posterior_variance), Q,
            None if cx is None else cx.ctypes.data, None if cw is None else cw.ctypes.data, ptr(out), ptr(ws), nb,
            stream_ptr()))
        return out

    def energy(self, batch_ind=None, cubature=None, **kwargs):
        """variational free energy (inference.py:197-222)"""
        return -(self.expected_density(cubature) - self.compute_kl())

    def predict(self, X, R=None):
        """basemodels.py:1033-1069: latent mean and variance at test inputs"""
        sm, sP, gain = self._smoothed_states()
        tm, tv = ops.temporal_conditional(self.Z, np.asarray(X, dtype=np.float64).reshape(-1), sm, sP, gain, self.kernel,
                                          return_full=False)
        return tm.squeeze(), tv.squeeze()


SparseMarkovGP = SparseMarkovGaussianProcess
