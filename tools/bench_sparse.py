"""Sparse Markov GP timing aid (not the bench line): N data points, Mz inducing points, Matern-5/2, Bernoulli-probit,
one iteration = inference(lr=1) + energy().  SPARSE_N=10000000 SPARSE_MZ=100000 python tools/bench_sparse.py"""
import ctypes, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib
from _data import bench_inputs

N, Mz = int(os.environ.get('SPARSE_N', 10_000_000)), int(os.environ.get('SPARSE_MZ', 100_000))
torch.cuda.set_device(0)
t, dt, y = bench_inputs(N)
z = np.linspace(t[0], t[-1], Mz)
m = bn.models.SparseMarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, Z=z, parallel=True)
step = lambda: (m.inference(lr=1.0), m.energy())[1]
for _ in range(2):
    E = step()
torch.cuda.synchronize()
L = _lib.lib()
L.bn_timing_enable(1)
iters = 5
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
kern = {}
for line in buf.value.decode().strip().splitlines():
    name, cnt, tot = line.split()
    kern[name] = {'launches_per_iter': int(cnt) / iters, 'ms_per_iter': float(tot) / iters}
print(json.dumps({'config': 'SparseMarkovVariationalGP Matern52 Bernoulli-probit N=%d Mz=%d' % (N, Mz), 'ms_per_iter': ms,
                  'data_points_per_s': N / (ms * 1e-3), 'energy': float(E), 'kernels': kern}))
