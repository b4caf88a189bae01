"""ctypes front-end of the host emulation harness (tests/hostemu/hostemu.cu).  Test infrastructure."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostemu', 'hostemu.cu')
LIB = os.path.join(HERE, 'hostemu', 'libbn_hostemu.so')
CSRC = os.path.join(os.path.dirname(HERE), 'bayesnewton_b200', 'csrc')


class KernelSpec(C.Structure):
    _fields_ = [('family', C.c_int32), ('n_components', C.c_int32),
                ('variance', C.c_double * 4), ('lengthscale', C.c_double * 4)]


def _digest():
    import hashlib
    h = hashlib.sha256()
    deps = [SRC] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'bn_b200.h'))
    for d in deps:
        h.update(open(d, 'rb').read())
    return h.hexdigest()


def load():
    """builds the harness when its sources changed (content hash, so a copied tree does not rebuild)"""
    stamp = LIB + '.sha256'
    dig = _digest()
    if not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig):
        subprocess.run(['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O2', '-std=c++17',
                        '-shared', '-Xcompiler', '-fPIC', SRC, '-o', LIB], check=True, capture_output=True)
        open(stamp, 'w').write(dig)
    return C.CDLL(LIB)


def spec(family, variances, lengthscales):
    s = KernelSpec()
    s.family, s.n_components = family, len(variances)
    for i, (v, l) in enumerate(zip(variances, lengthscales)):
        s.variance[i], s.lengthscale[i] = v, l
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def kalman_filter(lib, sp, form, dt, y, R, mask=None, L=8, world=1, return_predict=False, want_ell=True):
    N = dt.shape[0]
    nc = sp.n_components
    d = {1: 1, 2: 2, 3: 3, 4: 4}[sp.family] * nc
    dt, y, R = (np.ascontiguousarray(a, dtype=np.float64) for a in (dt, y, R))
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    ell = np.zeros(1)
    m = np.zeros((N, d, 1))
    P = np.zeros((N, d, d))
    rc = lib.emu_kalman_filter(C.byref(sp), form, C.c_longlong(N), L, world, _p(dt), _p(y), _p(R), _p(mk),
                               int(return_predict), _p(ell) if want_ell else None, _p(m), _p(P))
    assert rc == 0
    return ell[0], m, P


def rts_smoother(lib, sp, form, dt, fm, fP, L=8, world=1, return_full=False):
    N = dt.shape[0]
    nc = sp.n_components
    d = {1: 1, 2: 2, 3: 3, 4: 4}[sp.family] * nc
    od = d if return_full else nc
    dt, fm, fP = (np.ascontiguousarray(a, dtype=np.float64) for a in (dt, fm, fP))
    sm = np.zeros((N, od, 1))
    sP = np.zeros((N, od, od))
    G = np.zeros((N, d, d))
    rc = lib.emu_rts_smoother(C.byref(sp), form, C.c_longlong(N), L, world, _p(dt), _p(fm), _p(fP), int(return_full),
                              _p(sm), _p(sP), _p(G))
    assert rc == 0
    return sm, sP, G


def update_posterior(lib, sp, dt, y, R, mask=None, L=8, world=1, want_ell=True, want_grad=False):
    """fused filter + smoother (csrc/up_impl.cuh): returns ell, post_mean [N,D,1], post_cov [N,D,D]
    (+ d ell / d variance[NC], d ell / d lengthscale[NC] with want_grad)"""
    N = dt.shape[0]
    D = sp.n_components
    dt, y, R = (np.ascontiguousarray(a, dtype=np.float64) for a in (dt, y, R))
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    ell = np.zeros(1)
    pm = np.zeros((N, D, 1))
    pc = np.zeros((N, D, D))
    dvar, dlen = np.zeros(D), np.zeros(D)
    rc = lib.emu_update_posterior(C.byref(sp), C.c_longlong(N), L, world, _p(dt), _p(y), _p(R), _p(mk),
                                  _p(ell) if want_ell else None, _p(pm), _p(pc),
                                  _p(dvar) if want_grad else None, _p(dlen) if want_grad else None)
    assert rc == 0, rc
    if want_grad:
        return ell[0], pm, pc, dvar, dlen
    return ell[0], pm, pc


class SiteArgs(C.Structure):
    _fields_ = [('method', C.c_int32), ('likelihood', C.c_int32), ('lik_param', C.c_double), ('N', C.c_int64),
                ('D', C.c_int32), ('Q', C.c_int32), ('cub_x', C.c_void_p), ('cub_w', C.c_void_p), ('y', C.c_void_p),
                ('post_mean', C.c_void_p), ('post_cov', C.c_void_p), ('lr', C.c_double), ('power', C.c_double),
                ('ensure_psd', C.c_int32), ('pad_', C.c_int32), ('nat1', C.c_void_p), ('nat2', C.c_void_p),
                ('site_mean', C.c_void_p), ('site_cov', C.c_void_p), ('out_mean', C.c_void_p),
                ('out_jac', C.c_void_p), ('out_hess', C.c_void_p), ('diffs', C.c_void_p), ('lik_param2', C.c_double)]


METHODS = {'vi': 1, 'ep': 2, 'newton': 3, 'pl': 4}
LIKS = {'gaussian': 1, 'probit': 2, 'logit': 3, 'het_softplus': 4, 'het_exp': 5, 'poisson': 6, 'studentst': 7, 'gamma': 8,
        'negbin': 9, 'beta': 10}


def site_update(lib, method, lik, lik_param, y, post_mean, post_cov, nat1, nat2, lr=1.0, power=1.0, ensure_psd=True,
                cub=None, use_table=True, lik_param2=0.0):
    """returns dict with new nat1, nat2, site_mean, site_cov, mean, jac, hess, diffs"""
    N, D = post_mean.shape[0], post_mean.shape[1]
    keep = []

    def arr(a):
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        keep.append(a)
        return a
    a = SiteArgs()
    a.method, a.likelihood, a.lik_param, a.N, a.D = METHODS[method], LIKS[lik], lik_param, N, D
    a.lik_param2 = lik_param2
    if cub is not None:
        cx, cw = arr(cub[0]), arr(cub[1])
        a.Q, a.cub_x, a.cub_w = cw.shape[0], cx.ctypes.data, cw.ctypes.data
    yv, pm, pc = arr(y), arr(post_mean), arr(post_cov)
    a.y, a.post_mean, a.post_cov = yv.ctypes.data, pm.ctypes.data, pc.ctypes.data
    a.lr, a.power, a.ensure_psd = lr, power, int(ensure_psd)
    out = dict(nat1=arr(nat1), nat2=arr(nat2), site_mean=np.zeros((N, D, 1)), site_cov=np.zeros((N, D, D)),
               mean=np.zeros((N, D, 1)), jac=np.zeros((N, D, 1)), hess=np.zeros((N, D, D)), diffs=np.zeros(2))
    a.nat1, a.nat2 = out['nat1'].ctypes.data, out['nat2'].ctypes.data
    a.site_mean, a.site_cov = out['site_mean'].ctypes.data, out['site_cov'].ctypes.data
    a.out_mean, a.out_jac, a.out_hess = out['mean'].ctypes.data, out['jac'].ctypes.data, out['hess'].ctypes.data
    a.diffs = out['diffs'].ctypes.data
    rc = lib.emu_site_update(C.byref(a), int(use_table))
    assert rc == 0
    return out


def expected_density(lib, method, lik, lik_param, y, post_mean, post_cov, nat1=None, nat2=None, power=1.0, cub=None,
                     use_table=True):
    N, D = post_mean.shape[0], post_mean.shape[1]
    keep = []

    def arr(x):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        keep.append(x)
        return x
    a = SiteArgs()
    a.method, a.likelihood, a.lik_param, a.N, a.D = METHODS[method], LIKS[lik], lik_param, N, D
    if cub is not None:
        cx, cw = arr(cub[0]), arr(cub[1])
        a.Q, a.cub_x, a.cub_w = cw.shape[0], cx.ctypes.data, cw.ctypes.data
    yv, pm, pc = arr(y), arr(post_mean), arr(post_cov)
    a.y, a.post_mean, a.post_cov = yv.ctypes.data, pm.ctypes.data, pc.ctypes.data
    a.power = power
    if nat1 is not None:
        n1, n2 = arr(nat1), arr(nat2)
        a.nat1, a.nat2 = n1.ctypes.data, n2.ctypes.data
    vals, s = np.zeros(N), np.zeros(1)
    rc = lib.emu_expected_density(C.byref(a), _p(vals), _p(s), int(use_table))
    assert rc == 0
    return vals, s[0]


def iter_pass(lib, sp, dt, y, site_mean, site_cov, mode, method='vi', lik='probit', lik_param=0.0, cub=None, mask=None,
              lr=1.0, power=1.0, ensure_psd=True, L=8, world=1, use_table=True, spec=False):
    """fused iteration pass on tiled state (csrc/iter_impl.cuh).  mode 0 plain / 1 sites / 2 energy.
    Returns dict(ell, sums, site_mean, site_cov, post_mean, post_cov) with linear [N] arrays."""
    N = dt.shape[0]
    dt, y = (np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (dt, y))
    sy = np.ascontiguousarray(site_mean, dtype=np.float64).reshape(-1).copy()
    sR = np.ascontiguousarray(site_cov, dtype=np.float64).reshape(-1).copy()
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
    ell, sums, jm = np.zeros(1), np.zeros(2), np.zeros(1)
    pm, pc = np.zeros(N), np.zeros(N)
    Q, cx, cw = 0, None, None
    if cub is not None:
        cx = np.ascontiguousarray(cub[0], dtype=np.float64).reshape(-1)
        cw = np.ascontiguousarray(cub[1], dtype=np.float64).reshape(-1)
        Q = cw.shape[0]
    liks = dict(LIKS, poisson=6)
    rc = lib.emu_iter_pass(C.byref(sp), C.c_longlong(N), L, world, _p(dt), _p(y), _p(sy), _p(sR), _p(mk), mode,
                           METHODS[method], liks[lik], C.c_double(lik_param), Q, _p(cx), _p(cw), C.c_double(lr),
                           C.c_double(power), int(ensure_psd), int(use_table), _p(ell), _p(sums), _p(pm), _p(pc),
                           int(spec), _p(jm))
    assert rc == 0, rc
    return dict(ell=ell[0], sums=sums, site_mean=sy, site_cov=sR, post_mean=pm, post_cov=pc, jstar_mean=jm[0])


def gd_kf(lib, form, As, Qs, H, ys, Rs, m0, P0, masks=None, L=16, return_predict=False, want_ell=True):
    """warp-cooperative small-d filter bodies (csrc/gd_impl.cuh) on the host: ell, fms [N,d,1], fPs [N,d,d]"""
    As, Qs, H = (np.ascontiguousarray(a, dtype=np.float64) for a in (As, Qs, H))
    ys, Rs = np.ascontiguousarray(ys, dtype=np.float64), np.ascontiguousarray(Rs, dtype=np.float64)
    m0, P0 = np.ascontiguousarray(m0, dtype=np.float64), np.ascontiguousarray(P0, dtype=np.float64)
    N, d, D = As.shape[0], As.shape[1], H.shape[0]
    mk = None if masks is None else np.ascontiguousarray(masks, dtype=np.uint8)
    ell, fms, fPs = np.zeros(1), np.zeros((N, d, 1)), np.zeros((N, d, d))
    lib.emu_gd_kf.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int] + [C.c_void_p] * 3
    rc = lib.emu_gd_kf(form, N, L, d, D, _p(As), _p(Qs), _p(H), _p(ys), _p(Rs), _p(m0), _p(P0), _p(mk), int(return_predict),
                       _p(ell) if want_ell else None, _p(fms), _p(fPs))
    assert rc == 0
    return ell[0], fms, fPs


def gd_rts(lib, form, fms, fPs, As, Qs, H, L=16, return_full=False):
    fms, fPs, As, Qs, H = (np.ascontiguousarray(a, dtype=np.float64) for a in (fms, fPs, As, Qs, H))
    N, d, Df = As.shape[0], As.shape[1], H.shape[0]
    Do = d if return_full else Df
    sms, sPs, gains = np.zeros((N, Do, 1)), np.zeros((N, Do, Do)), np.zeros((N, d, d))
    lib.emu_gd_rts.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3
    rc = lib.emu_gd_rts(form, N, L, d, Df, _p(fms), _p(fPs), _p(As), _p(Qs), _p(H), int(return_full), _p(sms), _p(sPs), _p(gains))
    assert rc == 0
    return sms, sPs, gains
