"""Seeded synthetic inputs shared by the CPU and GPU tests (the reference's generators are unseeded:
tests/test_gp_vs_markovgp_reg.py:15-27, tests/test_gp_vs_markovgp_class.py:9-18)."""
import numpy as np


def regression_data(N, seed=12345):
    rng = np.random.default_rng(seed)
    x = np.sort(100 * rng.random(N))
    f = 6 * np.sin(np.pi * x / 10.0) / (np.pi * x / 10.0 + 1)
    y = f + np.sqrt(0.05) * rng.standard_normal(N)
    return x, y


def classification_data(N, seed=12345):
    x, y_ = regression_data(N, seed)
    y = np.sign(y_)
    y[y == -1] = 0
    return x, y


def filter_problem(N, D=1, seed=0, missing=0.1, dt_lo=0.1, dt_hi=0.5, offdiag=True):
    """dt[N] (dt[0]=0), y[N,D,1], R[N,D,D] SPD, mask[N,D,1]"""
    rng = np.random.default_rng(seed)
    dt = np.concatenate([[0.0], dt_lo + (dt_hi - dt_lo) * rng.uniform(size=N - 1)])
    y = rng.standard_normal((N, D, 1))
    R = np.zeros((N, D, D))
    for i in range(D):
        R[:, i, i] = 0.3 + rng.uniform(size=N)
    if D == 2 and offdiag:
        off = 0.1 * rng.standard_normal(N)
        R[:, 0, 1] = off
        R[:, 1, 0] = off
    mask = rng.uniform(size=(N, D, 1)) < missing
    return dt, y, R, mask


def bench_inputs(N, seed=0):
    """C2/C5 workload of SURVEY section 8(d): dt_0 = 0, dt_k = 0.1 + 0.2 u_k, y = 1[2 sin(.3t)+sin(.05t)+.5 eps > 0]"""
    rng = np.random.default_rng(seed)
    dt = 0.1 + 0.2 * rng.random(N)
    dt[0] = 0.0
    t = np.cumsum(dt)
    eps = np.random.default_rng(seed + 1).standard_normal(N)
    y = (2 * np.sin(0.3 * t) + np.sin(0.05 * t) + 0.5 * eps > 0).astype(np.float64)
    return t, dt, y


def rel_err(a, b):
    """normwise relative error max|a-b| / max|b| with identical NaN patterns required"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert (np.isnan(a) == np.isnan(b)).all(), 'NaN patterns differ'
    if np.isnan(b).all():
        return 0.0
    return float(np.nanmax(np.abs(a - b)) / max(np.nanmax(np.abs(b)), 1e-300))
