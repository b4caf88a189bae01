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


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.run(['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O2', '-std=c++17',
                        '-shared', '-Xcompiler', '-fPIC', SRC, '-o', LIB], check=True, capture_output=True)
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
