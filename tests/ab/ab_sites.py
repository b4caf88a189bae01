"""A/B check of the GPU site kernels against the oracle at large N (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/ab -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bayesnewton_b200 as bn
from oracle import sites

rng = np.random.default_rng(0)
N = 1_000_000
y = (rng.uniform(size=N) < 0.5).astype(float)
m = 1.2 * rng.standard_normal(N)
v = 0.15 + 0.3 * rng.uniform(size=N)
E0 = sites.variational_expectation(sites.Bernoulli('probit'), y, m, v)
lg = bn.likelihoods.Bernoulli('probit')
E1 = lg.variational_expectation(y, m[:, None, None], v[:, None, None])
for k, nm in enumerate(['E', 'dE', 'd2E']):
    a = E1[k].cpu().numpy().reshape(-1)
    d = np.abs(a - E0[k])
    i = int(np.argmax(d))
    print(nm, 'max abs err %.3e at i=%d (m=%.4f v=%.4f y=%g got=%.17g ref=%.17g)  sum diff %.3e' %
          (d.max(), i, m[i], v[i], y[i], a[i], E0[k][i], a.sum() - E0[k].sum()))
# where are the large errors?
a = E1[0].cpu().numpy().reshape(-1)
bad = np.where(np.abs(a - E0[0]) > 1e-12)[0]
print('n bad', bad.size)
if bad.size:
    print('bad m range', m[bad].min(), m[bad].max(), 'v', v[bad].min(), v[bad].max(), 'first idx', bad[:10], 'idx mod 256', bad[:10] % 256)
