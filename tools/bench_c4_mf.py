"""C4 shape with the mean-field filter / smoother (MarkovVariationalMeanFieldGP): timing aid, one JSON line."""
import ctypes, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib
Nt, G = int(os.environ.get('C4_NT', 2000)), 16
torch.cuda.set_device(0)
t = np.arange(Nt, dtype=np.float64)
a = np.linspace(-3, 3, G)
r = np.array([[u, v] for u in a for v in a])
R = np.tile(r[None], (Nt, 1, 1))
Y = np.sin(t / 10)[:, None] + np.sin(r[:, 0])[None] + np.cos(r[:, 1])[None] + 0.1 * np.random.default_rng(1).standard_normal((Nt, G * G))
Y[np.random.default_rng(2).uniform(size=Y.shape) < 0.05] = np.nan
K = bn.kernels
kern = bn.spacetime.SpatioTemporalKernel(K.Matern32(1.0, 5.0), bn.spacetime.Separable([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)]), z=r)
m = bn.models.MarkovVariationalMeanFieldGP(kernel=kern, likelihood=bn.likelihoods.Gaussian(1.0), X=t, Y=Y, R=R)
step = lambda: (m.inference(lr=1.0, return_state=False), m.energy())[1]
E = step()
torch.cuda.synchronize()
L = _lib.lib()
L.bn_timing_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
E = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
buf = ctypes.create_string_buffer(1 << 16)
L.bn_timing_report(buf, len(buf))
L.bn_timing_enable(0)
kern_ms = {ln.split()[0]: float(ln.split()[2]) for ln in buf.value.decode().strip().splitlines()}
print(json.dumps({'config': 'C4 mean-field N_t=%d M=%d' % (Nt, G * G), 'ms_per_iter': ms, 'time_steps_per_s': Nt / (ms * 1e-3), 'energy': float(E), 'kernels_ms': kern_ms}))
