"""Hyper-parameter gradient of the filter log-likelihood (oracle; test infrastructure).

The reference obtains d energy / d hyper-parameters by reverse-mode autodiff
(``objax.GradValues(model.energy, model.vars())``, ``README.md:56-70``,
``demos/regression.py:63-70``) of ``energy()`` (``inference.py:130-154,197-222,286-325``).
Sites and posterior are ``StateVar`` s (``basemodels.py:61-64,134-135``) -- constants -- so the
kernel hyper-parameters reach the energy of a temporal model only through the filter
log-likelihood ``compute_log_lik`` (``basemodels.py:726-741``): d E / d theta = - d ell / d theta.

Three independent evaluations of d ell / d theta live here:

* ``kf_vjp``                the literal reverse sweep of ``_sequential_kf`` (``ops.py:154-180``) --
                            what autodiff of the ``lax.scan`` computes, masks included;
* ``kf_grad_smoother``      the closed form the CUDA path evaluates inside its smoother sweep:
                            the adjoint of the predicted state is available from the smoothed state,
                            d ell / d m^-_k = (P^-_k)^-1 (sm_k - m^-_k),
                            d ell / d P^-_k = 1/2 (P^-_k)^-1 (sP_k - P^-_k + delta delta^T) (P^-_k)^-1
                            (Fisher's identity applied to the prediction as the prior of steps k..N);
* ``ell_grad_fd``           central finite differences (``longdouble`` capable).

Derivatives of the discretisation (dA_k, dQ_k, dPinf per hyper-parameter) are obtained by complex-step
differentiation of the closed forms in ``ssm.py`` -- exact to rounding, no hand-written derivative shared
with the CUDA code.
"""
import numpy as np
from . import kalman, ssm
from .linalg import T


# ------------------------------------------------------------------------------------------ reverse sweep
def _masked_logpdf_grads(e, S, mask):
    """d logpdf / d e and d logpdf / d S of utils.py:376-396 (masked rows/cols carry no density)"""
    D = S.shape[0]
    keep = np.ones(D, dtype=bool) if mask is None else ~np.asarray(mask).reshape(-1)
    ge, gS = np.zeros((D, 1)), np.zeros((D, D))
    if keep.any():
        idx = np.where(keep)[0]
        Suu = S[np.ix_(idx, idx)]
        Si = np.linalg.inv(Suu)
        v = Si @ e[idx]
        ge[idx] = -v
        gS[np.ix_(idx, idx)] = 0.5 * (v @ v.T - Si)
    return ge, gS


def kf_vjp(As, Qs, H, ys, Rs, m0, P0, masks=None):
    """cotangents of ell = _sequential_kf(...)[0] w.r.t. As, Qs, ys, Rs, m0, P0 (symmetric form, see below).

    Forward step (ops.py:156-175): m- = A m, P- = A P A^T + Q, G = H P-, S = G H^T + R, e = y - H m-,
    K = G^T S^-1, m+ = m- + K e, P+ = P- - K G, ell += logpdf(e; S, mask).
    The adjoints of symmetric quantities are kept symmetric; contracted with symmetric perturbations
    (all a hyper-parameter can produce) they give exactly what autodiff gives."""
    N, d = ys.shape[0], P0.shape[0]
    # forward, storing what the reverse sweep needs
    ms, Ps = np.zeros((N + 1, d, 1)), np.zeros((N + 1, d, d))
    ms[0], Ps[0] = m0, P0
    ell = 0.0
    for k in range(N):
        A = As[k]
        m_, P_ = A @ ms[k], A @ Ps[k] @ A.T + Qs[k]
        G = H @ P_
        S = G @ H.T + Rs[k]
        e = ys[k] - H @ m_
        K = np.linalg.solve(S, G).T
        ms[k + 1], Ps[k + 1] = m_ + K @ e, P_ - K @ G
        mk = None if masks is None else masks[k]
        ell += float(kalman.mvn_logpdf(ys[k], H @ m_, S, mk))
    Ab, Qb = np.zeros_like(As), np.zeros_like(Qs)
    yb, Rb = np.zeros_like(ys), np.zeros_like(Rs)
    mb, Pb = np.zeros((d, 1)), np.zeros((d, d))
    for k in range(N - 1, -1, -1):
        A, m, P = As[k], ms[k], Ps[k]
        m_, P_ = A @ m, A @ P @ A.T + Qs[k]
        G = H @ P_
        S = G @ H.T + Rs[k]
        e = ys[k] - H @ m_
        Si = np.linalg.inv(S)
        K = G.T @ Si
        mk = None if masks is None else masks[k]
        ge, gS = _masked_logpdf_grads(e, S, mk)
        Kb = mb @ e.T - Pb @ G.T
        eb = ge + K.T @ mb
        Sb = gS - 0.5 * (Si @ Kb.T @ K + K.T @ Kb @ Si)
        Gb = -K.T @ Pb + Si @ Kb.T + Sb @ H
        Rb[k], yb[k] = Sb, eb
        mb_ = mb - H.T @ eb
        Pb_ = Pb + 0.5 * (H.T @ Gb + Gb.T @ H)
        Qb[k] = Pb_
        Ab[k] = 2.0 * Pb_ @ A @ P + mb_ @ m.T
        Pb = A.T @ Pb_ @ A
        mb = A.T @ mb_
    return ell, dict(As=Ab, Qs=Qb, ys=yb, Rs=Rb, m0=mb, P0=Pb)


# ------------------------------------------------------------------------------------------ d(A, Q, Pinf)/d theta
def _with_params(kernel, params, dtype):
    """a copy of `kernel` with hyper-parameters `params` = [(variance, lengthscale), ...] in `dtype`"""
    if isinstance(kernel, ssm.Independent):
        return ssm.Independent([type(k)(p[0], p[1], dtype=dtype) for k, p in zip(kernel.kernels, params)])
    return type(kernel)(params[0][0], params[0][1], dtype=dtype)


def kernel_params(kernel):
    ks = kernel.kernels if isinstance(kernel, ssm.Independent) else [kernel]
    return [(float(k.variance), float(k.lengthscale)) for k in ks]


def discretisation_derivatives(kernel, dt):
    """[(dAs, dQs, dPinf)] per hyper-parameter, ordered (variance_0, lengthscale_0, variance_1, ...);
    complex-step differentiation (h = 1e-30) of kernels.py:158-365 / ops.py:149-151"""
    p0 = kernel_params(kernel)
    out = []
    h = 1e-30
    for c in range(len(p0)):
        for which in (0, 1):
            p = [list(map(complex, q)) for q in p0]
            p[c][which] += 1j * h
            kc = _with_params(kernel, p, np.complex128)
            Pinf = kc.stationary_covariance()
            As = np.stack([kc.state_transition(x) for x in np.asarray(dt).reshape(-1)])
            Qs = Pinf - As @ Pinf @ T(As)
            out.append((As.imag / h, Qs.imag / h, Pinf.imag / h))
    return out


def ell_grad_adjoint(kernel, dt, ys, Rs, masks=None):
    """(ell, d ell / d [variance_c, lengthscale_c ...]) by the reverse sweep"""
    As, Qs = ssm.discretise(kernel, dt)
    Pinf = kernel.stationary_covariance()
    H = kernel.measurement_model()
    m0 = np.zeros((Pinf.shape[0], 1))
    ell, bar = kf_vjp(As, Qs, H, ys, Rs, m0, Pinf, masks)
    g = []
    for dA, dQ, dP in discretisation_derivatives(kernel, dt):
        g.append(np.sum(bar['As'] * dA) + np.sum(bar['Qs'] * dQ) + np.sum(bar['P0'] * dP))
    return ell, np.array(g)


# ------------------------------------------------------------------------------------------ smoother identity
def kf_grad_smoother(kernel, dt, ys, Rs):
    """d ell / d theta from filtered + smoothed states (no mask: the identity needs ell to be the true
    marginal likelihood of the model the recursion runs, which the reference's mask rule breaks)"""
    As, Qs = ssm.discretise(kernel, dt)
    Pinf = kernel.stationary_covariance()
    H = kernel.measurement_model()
    d = Pinf.shape[0]
    m0 = np.zeros((d, 1))
    N = ys.shape[0]
    ell, fms, fPs = kalman.sequential_kf(As, Qs, H, ys, Rs, m0, Pinf, np.zeros_like(ys, dtype=bool))
    dts = np.concatenate([np.asarray(dt)[1:], [0.0]])
    As_s, Qs_s = ssm.discretise(kernel, dts)
    sms, sPs, _ = kalman.sequential_rts(fms, fPs, As_s, Qs_s, H, True)
    Gamma = np.zeros((d, d))          # sum_k (M_k - A_k^T M_k A_k)  (+ A_0^T M_0 A_0: the prior itself is Pinf)
    Abar = np.zeros((N, d, d))
    for k in range(N):
        A = As[k]
        mprev, Pprev = (m0, Pinf) if k == 0 else (fms[k - 1], fPs[k - 1])
        m_, P_ = A @ mprev, A @ Pprev @ A.T + Qs[k]
        delta = sms[k] - m_
        Pi = np.linalg.inv(P_)
        v = Pi @ delta
        M = 0.5 * Pi @ (sPs[k] - P_ + delta @ delta.T) @ Pi
        Gamma += M - A.T @ M @ A
        if k == 0:
            Gamma += A.T @ M @ A
        Abar[k] = 2.0 * M @ A @ (Pprev - Pinf) + v @ mprev.T
    g = []
    for dA, dQ, dP in discretisation_derivatives(kernel, dt):
        g.append(np.sum(Gamma * dP) + np.sum(Abar * dA))
    return float(ell), np.array(g)


# ------------------------------------------------------------------------------------------ finite differences
def ell_grad_fd(kernel, dt, ys, Rs, masks=None, h=1e-6, dtype=np.longdouble):
    p0 = kernel_params(kernel)
    g = []

    def ell_at(p):
        kk = _with_params(kernel, p, dtype)
        mk = None if masks is None else masks
        ell, _ = kalman.kalman_filter(np.asarray(dt, dtype=dtype), kk, np.asarray(ys, dtype=dtype),
                                      np.asarray(Rs, dtype=dtype), mk)
        return ell

    for c in range(len(p0)):
        for which in (0, 1):
            pp = [list(q) for q in p0]
            pm = [list(q) for q in p0]
            pp[c][which] = dtype(pp[c][which]) + dtype(h)
            pm[c][which] = dtype(pm[c][which]) - dtype(h)
            g.append(float((ell_at(pp) - ell_at(pm)) / (2 * dtype(h))))
    return np.array(g)
