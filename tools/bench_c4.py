"""C4 timing aid (not the bench line): MarkovVariationalGP, SpatioTemporalKernel(Matern32 time, Separable Matern32 x2
space), 16 x 16 grid (M = 256, d = 512), Gaussian likelihood, 5 % missing; one iteration = inference(lr=1) + energy().
`C4_NT=10000 python tools/bench_c4.py` prints one JSON line with the iteration time, the per-kernel device times and
the fp64 roofline fraction (structure-exploiting FLOP count of SURVEY section 8d: F 0.209, S 0.852 GFLOP per step)."""
import ctypes, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib

Nt = int(os.environ.get('C4_NT', 1000))
G = int(os.environ.get('C4_GRID', 16))
iters = int(os.environ.get('C4_ITERS', 2))
torch.cuda.set_device(0)
rng = np.random.default_rng(0)
t = np.arange(Nt, dtype=np.float64)
a = np.linspace(-3, 3, G)
r = np.array([[u, v] for u in a for v in a])
R = np.tile(r[None], (Nt, 1, 1))
Y = np.sin(t / 10)[:, None] + np.sin(r[:, 0])[None] + np.cos(r[:, 1])[None] + 0.1 * np.random.default_rng(1).standard_normal((Nt, G * G))
Y[np.random.default_rng(2).uniform(size=Y.shape) < 0.05] = np.nan
K = bn.kernels
kern = bn.spacetime.SpatioTemporalKernel(K.Matern32(1.0, 5.0), bn.spacetime.Separable([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)]), z=r)
m = bn.models.MarkovVariationalGP(kernel=kern, likelihood=bn.likelihoods.Gaussian(1.0), X=t, Y=Y, R=R)
step = lambda: (m.inference(lr=1.0, return_state=False), m.energy())[1]
E = step()  # warm-up
torch.cuda.synchronize()
L = _lib.lib()
L.bn_timing_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    E = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
buf = ctypes.create_string_buffer(1 << 16)
L.bn_timing_report(buf, len(buf))
L.bn_timing_enable(0)
kern_ms = {}
for line in buf.value.decode().strip().splitlines():
    name, cnt, tot = line.split()
    kern_ms[name] = {'launches_per_iter': int(cnt) / iters, 'ms_per_iter': float(tot) / iters}
scratch = torch.empty(8 * 148 * 256 * 2, dtype=torch.float64, device='cuda')
peak = ctypes.c_double()
L.bn_measure_dfma_peak(scratch.data_ptr(), scratch.numel(), ctypes.byref(peak))
M = G * G
d = 2 * M
flop_F = M ** 3 / 3 + 2 * M * M * d + 2 * d * d * M + 8 * d * d
flop_S = d ** 3 / 3 + 2 * d ** 3 + 4 * d ** 3 + 8 * d * d
flop_iter = Nt * (3 * flop_F + 2 * flop_S)  # filter runs inside both posterior updates and once more in energy() unless served from cache
out = {'config': 'C4 N_t=%d, M=%d, d=%d' % (Nt, M, d), 'ms_per_iter': ms, 'time_steps_per_s': Nt / (ms * 1e-3), 'energy': float(E),
       'kernels': kern_ms, 'fp64_peak_tflops': 2 * peak.value / 1e12,
       'algorithmic_tflop_per_iter': flop_iter / 1e12, 'achieved_tflops': flop_iter / (ms * 1e-3) / 1e12,
       'frac_of_fp64_peak': flop_iter / (ms * 1e-3) / (2 * peak.value)}
for k, fl in (('st_filter', flop_F), ('st_smoother', flop_S)):
    if k in kern_ms:
        per = kern_ms[k]['ms_per_iter'] / kern_ms[k]['launches_per_iter']
        out[k + '_us_per_step'] = per * 1e3 / Nt
        out[k + '_frac_of_fp64_peak'] = Nt * fl / (per * 1e-3) / (2 * peak.value)
print(json.dumps(out))
