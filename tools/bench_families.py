"""iteration time (inference + energy, fused path) of every Matern family at N = 1e7, probit VI:
python tools/bench_families.py [N]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bayesnewton_b200 as bn
from _data import bench_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t, dt, y = bench_inputs(N)
out = {'N': N}
K = bn.kernels
for name, kern in (('matern12_d1', K.Matern12(1.0, 1.0)), ('matern32_d2', K.Matern32(1.0, 1.0)), ('matern52_d3', K.Matern52(1.0, 1.0)),
                   ('matern72_d4', K.Matern72(1.0, 1.0))):
    m = bn.models.MarkovVariationalGP(kernel=kern, likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
    for _ in range(3):
        m.inference(lr=1.0); E = m.energy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.inference(lr=1.0); E = m.energy()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[name] = {'ms_per_iteration': ms, 'steps_per_s': N / ms * 1e3, 'energy': float(E)}
    del m
    torch.cuda.empty_cache()
print(json.dumps(out))
