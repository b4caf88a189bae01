"""one C2 inference iteration for profiling: python tools/prof_iter.py [N] [iters]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bayesnewton_b200 as bn
from _data import bench_inputs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t, dt, y = bench_inputs(N)
m = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
for it in range(iters):
    m.inference(lr=1.0)
    E = m.energy()
torch.cuda.synchronize()
print('energy', float(E))
