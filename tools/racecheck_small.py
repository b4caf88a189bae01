"""small-N pass over the kernels that synchronise through shared memory or a grid barrier, for
`compute-sanitizer --tool racecheck` (and memcheck / synccheck): the chunk scan, the fused iteration with its warp
prescan and staged table, the unfused filter / smoother, the dense spatio-temporal persistent kernels"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import bayesnewton_b200 as bn
from _data import bench_inputs

N = 200_003  # more than 32 x 288 chunks: the one-launch upper scan levels and the ticketed sums are on the path
t, dt, y = bench_inputs(N)
y[::17] = np.nan
for par in (True, False):
    m = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=par)
    m.inference(lr=0.7)
    print('markov', par, float(m.energy()))
m = bn.models.MarkovExpectationPropagationGP(kernel=bn.kernels.Matern32(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
m.inference(lr=0.7)
print('ep', float(m.energy()))
# dense spatio-temporal model: persistent filter / smoother with the grid barrier
rng = np.random.default_rng(0)
Nt, M = 12, 16
X = np.linspace(0.0, 3.0, Nt)
R = np.tile(np.stack(np.meshgrid(np.linspace(0, 1, 4), np.linspace(0, 1, 4)), -1).reshape(1, M, 2), (Nt, 1, 1))
Y = rng.standard_normal((Nt, M))
kern = bn.spacetime.SpatioTemporalKernel(bn.kernels.Matern32(1.0, 1.0), bn.kernels.Matern32(1.0, 0.5), z=R[0])
st = bn.models.MarkovVariationalGP(kernel=kern, likelihood=bn.likelihoods.Gaussian(0.3), X=X, R=R, Y=Y)
st.inference(lr=1.0)
print('st', float(st.energy()))
torch.cuda.synchronize()
