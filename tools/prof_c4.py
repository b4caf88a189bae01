"""Phase profile of the dense filter / smoother at the C4 shape (SM cycles per step from bn_st_profile)."""
import ctypes, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib
Nt, G = int(os.environ.get('C4_NT', 200)), 16
M = G * G
a = np.linspace(-3, 3, G)
r = np.array([[u, v] for u in a for v in a])
K = bn.kernels
kern = bn.spacetime.SpatioTemporalKernel(K.Matern32(1.0, 5.0), bn.spacetime.Separable([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)]), z=r)
rng = np.random.default_rng(0)
dt = np.ones(Nt); dt[0] = 0
y = rng.standard_normal((Nt, M, 1))
A = rng.standard_normal((M, M))
Rn = np.tile((A @ A.T / M + np.eye(M))[None], (Nt, 1, 1))
mask = rng.uniform(size=(Nt, M, 1)) < 0.05
names = ['assemble', 'sync_assemble', 'cholesky', 'tiles', 'sync_tiles', 'panel_solve(last owner)', 'lookahead(last owner)', 'panel_wait(last owner)', 'factor_invert(sum)']
out = {}
def prof():
    buf = (ctypes.c_int64 * 16)()
    _lib.check(_lib.lib().bn_st_profile(buf, 16))
    return {n: buf[i] / Nt for i, n in enumerate(names)}
for mk in (None, mask):
    for _ in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ell, (fm, fP) = bn.ops.kalman_filter(dt, kern, y, Rn, mk)
        e1.record(); torch.cuda.synchronize()
    out['filter' + ('_masked' if mk is not None else '')] = {'us_per_step': e0.elapsed_time(e1) * 1e3 / Nt, 'cycles_per_step': prof()}
dts = np.concatenate([dt[1:], [0.0]])
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sm, sP, _ = bn.ops.rauch_tung_striebel_smoother(dts, kern, fm, fP, want_gains=False)
    e1.record(); torch.cuda.synchronize()
out['smoother'] = {'us_per_step': e0.elapsed_time(e1) * 1e3 / Nt, 'cycles_per_step': prof()}
print(json.dumps(out, indent=1))
