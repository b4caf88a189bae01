"""A/B of an inference iteration with the probit table on/off (BN_B200_PROBIT_TABLE env), run on the GPU box."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bayesnewton_b200 as bn
from _data import bench_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t, dt, y = bench_inputs(N)
m = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
for it in range(iters):
    m.inference(lr=1.0, return_state=False)
    V = m.expected_density()
    X = m.expected_density_pseudo()
    L = m.compute_log_lik()
    print('it %d  V %.17g  X %.17g  L %.17g  energy %.17g' % (it, float(V), float(X), float(L), float(-(V - (X - L)))))
pm = m.posterior_mean.cpu().numpy().reshape(-1); pv = m.posterior_variance.cpu().numpy().reshape(-1)
n1 = m.pseudo_likelihood.nat1.cpu().numpy().reshape(-1); n2 = m.pseudo_likelihood.nat2.cpu().numpy().reshape(-1)
np.save(os.path.join(ROOT, 'gpurun_out', 'ab_%s.npy' % os.environ.get('BN_B200_PROBIT_TABLE', '1')), np.stack([pm, pv, n1, n2]))
