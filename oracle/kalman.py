"""Kalman filter / RTS smoother, sequential and associative-scan forms (oracle; test infrastructure).

Function-by-function restatement of ``bayesnewton/ops.py:149-380``.  The
sequential forms are explicit Python loops over time (what ``lax.scan`` does);
the scan forms offer two evaluation orders of the same associative operator:
``order='tree'`` follows the recursive odd/even schedule of
``jax.lax.associative_scan`` (jax 0.4.14, ``lax/control_flow/loops.py``), and
``order='fold'`` is the plain left fold.  Any blocked GPU scan is a third
order; the spread between these two bounds what rounding may legally do.
"""
import math
import numpy as np
from .linalg import T, chol, cho_solve, solve, inv
from .ssm import discretise

LOG2PI = math.log(2 * math.pi)
INV2PI = (2 * math.pi) ** -1


def mvn_logpdf(x, mean, cov, mask=None):
    """utils.py:376-396 (batched over leading axes of cov)"""
    cov = np.asarray(cov)
    n = cov.shape[-1]
    x = np.asarray(x).reshape(cov.shape[:-2] + (n, 1))
    mean = np.asarray(mean).reshape(cov.shape[:-2] + (n, 1))
    if mask is not None:
        maskv = np.asarray(mask).reshape(cov.shape[:-2] + (n, 1))
        x = np.where(maskv, 0., x)
        mean = np.where(maskv, 0., mean)
        cov = np.where(maskv | T(maskv), 0., cov)
        eye = np.eye(n, dtype=bool)
        cov = np.where(eye & maskv, cov.dtype.type(INV2PI), cov)  # np.diag(mask) in the reference
    L = chol(cov)
    log_det = 2 * np.sum(np.log(np.abs(np.diagonal(L, axis1=-2, axis2=-1))), axis=-1)
    diff = x - mean
    dist = np.sum(diff * cho_solve(L, diff), axis=(-2, -1))
    return -0.5 * (dist + n * LOG2PI + log_det)


def sequential_kf(As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict=False):
    """ops.py:154-180"""
    N = ys.shape[0]
    d = P0.shape[0]
    dt = np.result_type(As, ys, P0)
    m, P, ell = m0.astype(dt), P0.astype(dt), dt.type(0.)
    fms = np.zeros((N, d, 1), dtype=dt)
    fPs = np.zeros((N, d, d), dtype=dt)
    for k in range(N):
        A, Q, R, y = As[k], Qs[k], noise_covs[k], ys[k]
        m_ = A @ m
        P_ = A @ P @ A.T + Q
        obs_mean = H @ m_
        HP = H @ P_
        S = HP @ H.T + R
        ell = ell + mvn_logpdf(y, obs_mean, S, masks[k])
        K = solve(S, HP).T
        m = m_ + K @ (y - obs_mean)
        P = P_ - K @ HP
        if return_predict:
            fms[k], fPs[k] = m_, P_
        else:
            fms[k], fPs[k] = m, P
    return ell, fms, fPs


def filtering_elements(As, Qs, H, ys, noise_covs, m0, P0):
    """ops.py:183-200 (vmapped) and :222-229 (first-element fix-up)"""
    Qs = Qs.copy()
    Qs[0] = P0
    HQ, HA = H @ Qs, H @ As
    S = HQ @ H.T + noise_covs
    SinvH = solve(S, np.broadcast_to(H, HQ.shape))
    K = Qs @ T(SinvH)
    AA = As - K @ HA
    b = K @ ys
    C = Qs - K @ HQ
    SinvHA = T(SinvH @ As)
    eta = SinvHA @ ys
    J = SinvHA @ HA
    S0 = H @ Qs[0] @ H.T + noise_covs[0]
    K0 = solve(S0, H @ Qs[0]).T
    b[0] = b[0] + (m0 - K0 @ H @ m0)
    return AA, b, C, J, eta


def filtering_operator(e1, e2):
    """ops.py:203-219; e1 = earlier, e2 = later; batched"""
    A1, b1, C1, J1, eta1 = e1
    A2, b2, C2, J2, eta2 = e2
    C1inv = inv(C1)
    temp = solve(C1inv + J2, C1inv)
    A2temp = A2 @ temp
    AA = A2temp @ A1
    b = A2temp @ (b1 + C1 @ eta2) + b2
    C = A2temp @ C1 @ T(A2) + C2
    A1temp = T(A1) @ T(temp)
    eta = A1temp @ (eta2 - J2 @ b1) + eta1
    J = A1temp @ J2 @ A1 + J1
    return AA, b, C, J, eta


def smoothing_elements(As, Qs, fms, fPs):
    """ops.py:318-325 and :314-315,343-345"""
    Pp = As @ fPs @ T(As) + Qs
    E = T(solve(Pp, As @ fPs))
    g = fms - E @ As @ fms
    L = fPs - E @ Pp @ T(E)
    gains = E.copy()
    E[-1] = 0.
    g[-1] = fms[-1]
    L[-1] = fPs[-1]
    return (E, g, L), gains


def smoothing_operator(e1, e2):
    """ops.py:328-335; called by the reversed scan with e1 = accumulated later part, e2 = earlier element"""
    E1, g1, L1 = e1
    E2, g2, L2 = e2
    return E2 @ E1, E2 @ g1 + g2, E2 @ L1 @ T(E2) + L2


def associative_scan(fn, elems, reverse=False, order='tree'):
    """inclusive scan along axis 0 in the evaluation order of jax.lax.associative_scan ('tree') or as a left fold"""
    if reverse:
        elems = tuple(e[::-1] for e in elems)
    if order == 'fold':
        n = elems[0].shape[0]
        out = [np.empty_like(e) for e in elems]
        acc = tuple(e[0:1] for e in elems)
        for o, a in zip(out, acc):
            o[0] = a[0]
        for k in range(1, n):
            acc = fn(acc, tuple(e[k:k + 1] for e in elems))
            for o, a in zip(out, acc):
                o[k] = a[0]
        res = tuple(out)
    else:
        res = _tree_scan(fn, tuple(elems))
    if reverse:
        res = tuple(e[::-1] for e in res)
    return res


def _tree_scan(fn, elems):
    n = elems[0].shape[0]
    if n < 2:
        return elems
    reduced = fn(tuple(e[0:-1:2] for e in elems), tuple(e[1::2] for e in elems))
    odd = _tree_scan(fn, reduced)
    if n % 2 == 0:
        even = fn(tuple(e[:-1] for e in odd), tuple(e[2::2] for e in elems))
    else:
        even = fn(odd, tuple(e[2::2] for e in elems))
    even = tuple(np.concatenate([e[0:1], r], axis=0) for e, r in zip(elems, even))
    out = []
    for ev, od in zip(even, odd):
        o = np.empty((n,) + ev.shape[1:], dtype=ev.dtype)
        o[0::2] = ev
        o[1::2] = od
        out.append(o)
    return tuple(out)


def parallel_kf(As, Qs, H, ys, noise_covs, m0, P0, masks, return_predict=False, order='tree'):
    """ops.py:237-253"""
    elems = filtering_elements(As, Qs, H, ys, noise_covs, m0, P0)
    final = associative_scan(filtering_operator, elems, order=order)
    fms, fPs = final[1], final[2]
    mpred = As @ np.concatenate([m0[None], fms[:-1]])
    Ppred = As @ np.concatenate([P0[None], fPs[:-1]]) @ T(As) + Qs
    ell = np.sum(mvn_logpdf(ys, H @ mpred, H @ Ppred @ H.T + noise_covs, masks))
    if return_predict:
        return ell, mpred, Ppred
    return ell, fms, fPs


def sequential_rts(fms, fPs, As, Qs, H, return_full):
    """ops.py:288-311"""
    N, d = fms.shape[0], fms.shape[1]
    Df = d if return_full else H.shape[0]
    sms = np.zeros((N, Df, 1), dtype=fms.dtype)
    sPs = np.zeros((N, Df, Df), dtype=fms.dtype)
    gains = np.zeros((N, d, d), dtype=fms.dtype)
    sm, sP = fms[-1], fPs[-1]
    for k in range(N - 1, -1, -1):
        fm, fP, A, Q = fms[k], fPs[k], As[k], Qs[k]
        pm = A @ fm
        AfP = A @ fP
        pP = AfP @ A.T + Q
        C = solve(pP, AfP).T
        sm = fm + C @ (sm - pm)
        sP = fP + C @ (sP - pP) @ C.T
        gains[k] = C
        if return_full:
            sms[k], sPs[k] = sm, sP
        else:
            sms[k], sPs[k] = H @ sm, H @ sP @ H.T
    return sms, sPs, gains


def parallel_rts(fms, fPs, As, Qs, H, return_full, order='tree'):
    """ops.py:338-354"""
    elems, gains = smoothing_elements(As, Qs, fms, fPs)
    final = associative_scan(smoothing_operator, elems, reverse=True, order=order)
    sms, sPs = final[1], final[2]
    if return_full:
        return sms, sPs, gains
    return H @ sms, H @ sPs @ H.T, gains


def kalman_filter(dt, kernel, y, noise_cov, mask=None, parallel=False, return_predict=False, order='tree'):
    """ops.py:256-285"""
    if mask is None:
        mask = np.zeros_like(y, dtype=bool)
    Pinf = kernel.stationary_covariance()
    minf = np.zeros((Pinf.shape[0], 1), dtype=Pinf.dtype)
    As, Qs = discretise(kernel, dt)
    H = kernel.measurement_model()
    if parallel:
        ell, means, covs = parallel_kf(As, Qs, H, y, noise_cov, minf, Pinf, mask, return_predict, order=order)
    else:
        ell, means, covs = sequential_kf(As, Qs, H, y, noise_cov, minf, Pinf, mask, return_predict)
    return ell, (means, covs)


def rauch_tung_striebel_smoother(dt, kernel, filter_mean, filter_cov, return_full=False, parallel=False,
                                 order='tree'):
    """ops.py:357-380"""
    As, Qs = discretise(kernel, dt)
    H = kernel.measurement_model()
    if parallel:
        return parallel_rts(filter_mean, filter_cov, As, Qs, H, return_full, order=order)
    return sequential_rts(filter_mean, filter_cov, As, Qs, H, return_full)
