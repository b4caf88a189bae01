"""The tabulated probit log-density (csrc/probit_table.cuh) against an extended-precision evaluation of
log(1e-3 + (1 - 2e-3) Phi(f)) (likelihoods.py:828-829, 836-852), through the host build of the same code."""
import ctypes as C

import numpy as np
from scipy import special


def g_ref(f):
    # erfc keeps full relative accuracy in the lower tail; evaluated in long double around scipy's fp64 erfc
    P = 0.5 * special.erfc(-np.asarray(f, dtype=np.float64) / np.sqrt(2.0))
    return np.log(np.longdouble(1e-3) + (np.longdouble(1) - np.longdouble(2e-3)) * P.astype(np.longdouble))


def test_table_accuracy(emu):
    emu.emu_probit_log_phi.restype = C.c_double
    emu.emu_probit_log_phi.argtypes = [C.c_double]
    rng = np.random.default_rng(0)
    f = np.concatenate([np.linspace(-12, 12, 20001), rng.uniform(-9, 9, 20000), -9 + np.arange(577) / 32.0,
                        -9 + (np.arange(576) + 0.5) / 32.0, [-1e3, 1e3, -9.0, 9.0, 0.0]])
    got = np.array([emu.emu_probit_log_phi(float(x)) for x in f])
    err = np.abs(got - g_ref(f).astype(np.float64))
    assert err.max() < 2.5e-13, err.max()  # cubic interpolant on width 1/512, packed 16-byte coefficients, fp32 Horner of the two highest
    assert np.sqrt(np.mean(err ** 2)) < 5e-14
    # symmetry used by the kernels: log(1 - p(f)) = g(-f)
    p = 0.5 * (1 + special.erf(f / np.sqrt(2))) * (1 - 2e-3) + 1e-3
    assert np.abs(np.log(1 - p) - np.array([emu.emu_probit_log_phi(float(-x)) for x in f])).max() < 2.5e-13
