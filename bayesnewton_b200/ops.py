"""Markov-GP ops with the reference's signatures (bayesnewton/ops.py:149-380), executed by libbn_b200.

    kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict=False)   ops.py:256
    rauch_tung_striebel_smoother(dt, kernel, filter_mean, filter_cov, return_full, parallel)    ops.py:357
    _sequential_kf / _parallel_kf (As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict)    ops.py:154,237
    _sequential_rts / _parallel_rts (fms, fPs, As, Qs, H, return_full)                          ops.py:288,338
    process_noise_covariance(A, Pinf)                                                           ops.py:149

Inputs may be numpy arrays or torch tensors (host or device); outputs are float64 CUDA tensors
with the reference's shapes ([N,d,1], [N,d,d], ...).  `parallel=False` runs the sequential
recursion on one GPU thread (the reference's lax.scan order); `parallel=True` runs the blocked
temporally-parallel scan.  Extra keyword `want_ell=False` skips the log-likelihood, the dead
code XLA would eliminate in update_posterior (basemodels.py:694,701).
"""
import numpy as np
import torch

from . import _lib
from ._util import as_dev, as_mask, ptr, stream_ptr, up_workspace, workspace
from .kernels import discretise

__all__ = ['temporal_conditional', 'update_posterior', 'kalman_filter', 'rauch_tung_striebel_smoother', '_sequential_kf', '_parallel_kf', '_sequential_rts',
           '_parallel_rts', 'process_noise_covariance', 'dare', 'rts_dare', 'kalman_filter_infinite_horizon',
           'rauch_tung_striebel_smoother_infinite_horizon']


def process_noise_covariance(A, Pinf):
    A, Pinf = as_dev(A), as_dev(Pinf)
    return Pinf - A @ Pinf @ A.transpose(-1, -2)


def _form(parallel):
    return _lib.BN_SCAN if parallel else _lib.BN_SEQUENTIAL


def _kf_arrays(form, As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict, want_ell=True, want_states=True):
    As, Qs, H, ys, Rs, m0, P0 = (as_dev(a) for a in (As, Qs, H, ys, noise_covs, m0, P0))
    N, d = As.shape[0], As.shape[-1]
    D = H.shape[0]
    mk = as_mask(masks)
    ell = torch.zeros((), dtype=torch.float64, device=As.device) if want_ell else None
    fms = torch.empty((N, d, 1), dtype=torch.float64, device=As.device) if want_states else None
    fPs = torch.empty((N, d, d), dtype=torch.float64, device=As.device) if want_states else None
    ws, nb = workspace(N, d, D)
    _lib.check(_lib.lib().bn_kf_arrays(form, N, d, D, ptr(As), ptr(Qs), ptr(H), ptr(ys), ptr(Rs), ptr(m0), ptr(P0),
                                       ptr(mk), int(bool(return_predict)), ptr(ell), ptr(fms), ptr(fPs),
                                       ptr(ws), nb, stream_ptr()))
    return ell, fms, fPs


def _sequential_kf(As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict=False):
    return _kf_arrays(_lib.BN_SEQUENTIAL, As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict)


def _parallel_kf(As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict=False):
    return _kf_arrays(_lib.BN_SCAN, As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict)


def _rts_arrays(form, fms, fPs, As, Qs, H, return_full, want_gains=True):
    fms, fPs, As, Qs, H = (as_dev(a) for a in (fms, fPs, As, Qs, H))
    N, d = As.shape[0], As.shape[-1]
    Df = H.shape[0]
    od = d if return_full else Df
    sms = torch.empty((N, od, 1), dtype=torch.float64, device=As.device)
    sPs = torch.empty((N, od, od), dtype=torch.float64, device=As.device)
    gains = torch.empty((N, d, d), dtype=torch.float64, device=As.device) if want_gains else None
    ws, nb = workspace(N, d, Df)
    _lib.check(_lib.lib().bn_rts_arrays(form, N, d, Df, ptr(fms), ptr(fPs), ptr(As), ptr(Qs), ptr(H),
                                        int(bool(return_full)), ptr(sms), ptr(sPs), ptr(gains), ptr(ws), nb,
                                        stream_ptr()))
    return sms, sPs, gains


def _sequential_rts(fms, fPs, As, Qs, H, return_full):
    return _rts_arrays(_lib.BN_SEQUENTIAL, fms, fPs, As, Qs, H, return_full)


def _parallel_rts(fms, fPs, As, Qs, H, return_full):
    return _rts_arrays(_lib.BN_SCAN, fms, fPs, As, Qs, H, return_full)


def kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict=False, want_ell=True,
                  want_states=True):
    """p(f_n | y_1..y_n) for all n; returns ell, (means [N,d,1], covs [N,d,d])"""
    if hasattr(kernel, 'temporal_kernel'):  # SpatioTemporalKernel: dense d = M n state (csrc/st.cu)
        from .spacetime import st_kalman_filter
        return st_kalman_filter(dt, kernel, y, noise_cov, mask, parallel, return_predict, want_ell)
    dt = as_dev(dt).reshape(-1)
    N = dt.shape[0]
    spec = kernel.spec() if hasattr(kernel, 'spec') else None
    if spec is None:  # generic path of the reference: materialise As, Qs, then the array-level filter
        As, Qs = discretise(kernel, dt)
        Pinf = as_dev(kernel.stationary_covariance())
        minf = torch.zeros((Pinf.shape[0], 1), dtype=torch.float64, device=dt.device)
        ell, m, P = _kf_arrays(_form(parallel), As, Qs, kernel.measurement_model(), y, noise_cov, minf, Pinf, mask,
                               return_predict, want_ell, want_states)
        return ell, (m, P)
    y, R = as_dev(y), as_dev(noise_cov)
    d = _lib.lib().bn_state_dim(spec)
    D = spec.n_components
    if y.numel() != N * D or R.numel() != N * D * D:
        raise ValueError('y must be [N,%d,1] and noise_cov [N,%d,%d] for N = %d steps' % (D, D, D, N))
    mk = as_mask(mask)
    ell = torch.zeros((), dtype=torch.float64, device=dt.device) if want_ell else None
    means = torch.empty((N, d, 1), dtype=torch.float64, device=dt.device) if want_states else None
    covs = torch.empty((N, d, d), dtype=torch.float64, device=dt.device) if want_states else None
    ws, nb = workspace(N, d, D)
    _lib.check(_lib.lib().bn_kalman_filter(spec, _form(parallel), N, ptr(dt), ptr(y), ptr(R), ptr(mk),
                                           int(bool(return_predict)), ptr(ell), ptr(means), ptr(covs), ptr(ws), nb,
                                           stream_ptr()))
    return ell, (means, covs)


def rauch_tung_striebel_smoother(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False,
                                 want_gains=True):
    """p(f_n | y_1..y_N); dt is the step OUT OF n (basemodels.py:700).  Returns (means, covs, gains)"""
    if hasattr(kernel, 'temporal_kernel'):
        from .spacetime import st_rts_smoother
        return st_rts_smoother(dt, kernel, filter_mean, filter_cov, return_full, parallel, want_gains)
    dt = as_dev(dt).reshape(-1)
    N = dt.shape[0]
    spec = kernel.spec() if hasattr(kernel, 'spec') else None
    if spec is None:
        As, Qs = discretise(kernel, dt)
        return _rts_arrays(_form(parallel), filter_mean, filter_cov, As, Qs, kernel.measurement_model(), return_full,
                           want_gains)
    fm, fP = as_dev(filter_mean), as_dev(filter_cov)
    d = _lib.lib().bn_state_dim(spec)
    Df = spec.n_components
    if fm.numel() != N * d or fP.numel() != N * d * d:
        raise ValueError('filter_mean must be [N,%d,1] and filter_cov [N,%d,%d] for N = %d' % (d, d, d, N))
    od = d if return_full else Df
    means = torch.empty((N, od, 1), dtype=torch.float64, device=dt.device)
    covs = torch.empty((N, od, od), dtype=torch.float64, device=dt.device)
    gains = torch.empty((N, d, d), dtype=torch.float64, device=dt.device) if want_gains else None
    ws, nb = workspace(N, d, Df)
    _lib.check(_lib.lib().bn_rts_smoother(spec, _form(parallel), N, ptr(dt), ptr(fm), ptr(fP),
                                          int(bool(return_full)), ptr(means), ptr(covs), ptr(gains), ptr(ws), nb,
                                          stream_ptr()))
    return means, covs, gains


def update_posterior(dt, kernel, y, noise_cov, mask=None, want_ell=False, want_grad=False):
    """kalman_filter followed by rauch_tung_striebel_smoother, `parallel=True` form, as ONE library call
    (MarkovGaussianProcess.update_posterior, basemodels.py:689-706).  Returns (ell or None, means [N,D,1],
    covs [N,D,D]) = (filter log-likelihood, H sm, H sP H^T).  Needs a kernel with an in-library
    discretisation (kernel.spec()); other kernels take the two stand-alone calls.
    want_grad appends d ell / d [variance_c...; lengthscale_c...] as a [2, NC] tensor (bn_update_posterior_grad:
    the reverse-mode pass objax.GradValues runs through compute_log_lik, basemodels.py:726-741)."""
    spec = kernel.spec() if hasattr(kernel, 'spec') else None
    if spec is None:
        raise NotImplementedError('the fused update needs kernel.spec(); use kalman_filter + rauch_tung_striebel_smoother')
    dt = as_dev(dt).reshape(-1)
    N = dt.shape[0]
    y, R = as_dev(y), as_dev(noise_cov)
    D = spec.n_components
    if y.numel() != N * D or R.numel() != N * D * D:
        raise ValueError('y must be [N,%d,1] and noise_cov [N,%d,%d] for N = %d steps' % (D, D, D, N))
    mk = as_mask(mask)
    ell = torch.zeros((), dtype=torch.float64, device=dt.device) if want_ell else None
    means = torch.empty((N, D, 1), dtype=torch.float64, device=dt.device)
    covs = torch.empty((N, D, D), dtype=torch.float64, device=dt.device)
    ws, nb = up_workspace(spec, N)
    if want_grad:
        if mk is not None:
            raise NotImplementedError('hyper-gradient with missing-data masks is not available (see include/bn_b200.h)')
        g = torch.zeros((2, D), dtype=torch.float64, device=dt.device)
        _lib.check(_lib.lib().bn_update_posterior_grad(spec, N, ptr(dt), ptr(y), ptr(R), ptr(ell), ptr(means),
                                                       ptr(covs), g[0].data_ptr(), g[1].data_ptr(), ptr(ws), nb,
                                                       stream_ptr()))
        return ell, means, covs, g
    _lib.check(_lib.lib().bn_update_posterior(spec, N, ptr(dt), ptr(y), ptr(R), ptr(mk), ptr(ell), ptr(means),
                                              ptr(covs), ptr(ws), nb, stream_ptr()))
    return ell, means, covs


def temporal_conditional(X, X_test, mean, cov, gain, kernel, return_full=True):
    """state distribution at the test inputs from the smoothed states of the neighbouring training inputs
    (utils.py:122-136).  X: the training inputs, with or without the dummy states at -1e10 / +1e10 the reference's
    predict() adds (basemodels.py:793-794) -- they are implied either way; mean [N,d,1], cov [N,d,d], gain [N,d,d]
    from rauch_tung_striebel_smoother(..., return_full=True).  Returns (test_mean [N*,d,1], test_cov [N*,d,d]);
    return_full=False applies the measurement model as predict() does: ([N*,Df,1], [N*,Df,Df])."""
    spec = kernel.spec() if hasattr(kernel, 'spec') else None
    if spec is None:
        raise NotImplementedError('prediction needs a kernel with an in-library discretisation (kernel.spec())')
    X = as_dev(X).reshape(-1)
    if X.numel() >= 2 and float(X[0]) <= -1e10 and float(X[-1]) >= 1e10:
        X = X[1:-1].contiguous()
    Xs = as_dev(X_test).reshape(-1)
    mean, cov, gain = as_dev(mean), as_dev(cov), as_dev(gain)
    N, Ns = X.shape[0], Xs.shape[0]
    d = _lib.lib().bn_state_dim(spec)
    if mean.numel() != N * d or cov.numel() != N * d * d or gain.numel() != N * d * d:
        raise ValueError('mean, cov, gain must be the full-state smoother output for the %d training inputs' % N)
    od = d if return_full else spec.n_components
    tm = torch.empty((Ns, od, 1), dtype=torch.float64, device=X.device)
    tc = torch.empty((Ns, od, od), dtype=torch.float64, device=X.device)
    _lib.check(_lib.lib().bn_temporal_conditional(spec, N, ptr(X), Ns, ptr(Xs), ptr(mean), ptr(cov), ptr(gain),
                                                  int(bool(return_full)), ptr(tm), ptr(tc), stream_ptr()))
    return tm, tc


def kalman_filter_pairs(dt, kernel, y, noise_cov, mask=None, parallel=False):
    """ops.py:383-426 (see sparse.kalman_filter_pairs)"""
    from .sparse import kalman_filter_pairs as kfp
    return kfp(dt, kernel, y, noise_cov, mask, parallel)


# ---------------------------------------------------------------------------------------------- infinite horizon
def _np64(x):
    return x.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(x) else np.asarray(x, dtype=np.float64)


def _chol_solve_host(P, B):
    """solve(P, B) by Cholesky (utils.py:14-19) on small host matrices"""
    L = np.linalg.cholesky(P)
    return np.linalg.solve(L.T, np.linalg.solve(L, B))


def dare(A, H, Q, R, Pinit, num_iters=20):
    """fixed-point iterations of the discrete algebraic Riccati equation (ops.py:796-824); d x d host algebra"""
    X = _np64(Pinit)
    A, H, Q, R = _np64(A), _np64(H), _np64(Q), _np64(R)
    for _ in range(num_iters):
        HX = H @ X
        S = HX @ H.T + R
        K = _chol_solve_host(S, HX).T
        X = A @ (X - K @ HX) @ A.T + Q
    return X


def rts_dare(A, Q, Pinf, num_iters=20):
    """ops.py:955-975"""
    X = _np64(Pinf)
    A, Q = _np64(A), _np64(Q)
    for _ in range(num_iters):
        X = A @ X @ A.T + Q
    return X


def _ih_ws(d, N):
    nb = int(_lib.lib().bn_ih_workspace_bytes(int(d), int(N)))
    return torch.empty(nb, dtype=torch.uint8, device=torch.device('cuda', torch.cuda.current_device())), nb


def _dt_at(dt, i):
    """one step length as a host float (ONE element crosses PCIe, not the series)"""
    flat = dt.reshape(-1) if torch.is_tensor(dt) else np.asarray(dt, dtype=np.float64).reshape(-1)
    return float(flat[min(i, flat.shape[0] - 1)])


def kalman_filter_infinite_horizon(dt, kernel, y, noise_cov, mask=None, parallel=False, heteroscedastic=False,
                                   noise_cov_tied=None, dare_iters=20, dare_init=None, want_ell=True):
    """ops.py:881-952 for one latent with one site per step: ell, (means [N,d,1], (Pdare, cov)).  The Riccati fixed point
    and the stationary gain are d x d host algebra; the O(N) mean recursion and the log-likelihood are bn_ih_filter."""
    y, R = as_dev(y).reshape(-1), as_dev(noise_cov).reshape(-1)
    N = y.shape[0]
    Pinf = _np64(kernel.stationary_covariance())
    A = _np64(kernel.state_transition(_dt_at(dt, 1)))
    Q = Pinf - A @ Pinf @ A.T
    H = _np64(kernel.measurement_model())
    if H.shape[0] != 1 or not (H[0, 1:] == 0).all():
        raise NotImplementedError('the infinite-horizon entries are built for one latent (H = e_0^T)')
    d = A.shape[0]
    tied = _np64(noise_cov_tied).reshape(1, 1)
    Pdare = dare(A, H, Q, tied, Pinf if dare_init is None else dare_init, dare_iters)
    S = H @ Pdare @ H.T + tied
    K = Pdare @ _chol_solve_host(S, H).T
    cov = Pdare - K @ H @ Pdare
    mk = as_mask(mask)
    ell = torch.zeros((), dtype=torch.float64, device=y.device) if want_ell else None
    means = torch.empty((N, d, 1), dtype=torch.float64, device=y.device)
    Rv = R if heteroscedastic else as_dev(tied.reshape(-1))
    ws, nb = _ih_ws(d, N)
    Ac, Pc = np.ascontiguousarray(A), np.ascontiguousarray(Pdare)
    _lib.check(_lib.lib().bn_ih_filter(_form(parallel), d, N, Ac.ctypes.data, Pc.ctypes.data, ptr(y), ptr(Rv),
                                       int(not heteroscedastic), ptr(mk), ptr(ell), ptr(means), ptr(ws), nb, stream_ptr()))
    return ell, (means, (Pdare, cov))


def rauch_tung_striebel_smoother_infinite_horizon(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False,
                                                  dare_iters=20, dare_init=None):
    """ops.py:1018-1068: means, covs (the fixed point, tiled over time), gains, dare_cov"""
    fm = as_dev(filter_mean)
    N, d = fm.shape[0], fm.shape[1]
    Pinf = _np64(kernel.stationary_covariance())
    A = _np64(kernel.state_transition(_dt_at(dt, 0)))
    H = _np64(kernel.measurement_model())
    Pdare, fcov = (_np64(c) for c in filter_cov)
    gain = fcov @ _chol_solve_host(Pdare, A).T
    Qdare = fcov - gain @ Pdare @ gain.T
    dare_cov = rts_dare(gain, Qdare, Pinf if dare_init is None else dare_init, dare_iters)
    od = d if return_full else 1
    means = torch.empty((N, od, 1), dtype=torch.float64, device=fm.device)
    ws, nb = _ih_ws(d, N)
    Ac, Gc = np.ascontiguousarray(A), np.ascontiguousarray(gain)
    _lib.check(_lib.lib().bn_ih_smoother(_form(parallel), d, N, Ac.ctypes.data, Gc.ctypes.data, ptr(fm), int(bool(return_full)),
                                         ptr(means), ptr(ws), nb, stream_ptr()))
    cov = dare_cov if return_full else H @ dare_cov @ H.T
    covs = as_dev(cov).reshape(1, od, od).expand(N, od, od).contiguous()
    gains = as_dev(gain).reshape(1, d, d).expand(N, d, d)
    return means, covs, gains, dare_cov
