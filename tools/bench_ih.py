"""infinite-horizon model (csrc/ih.cu): iteration time at N = 1e7, evenly spaced inputs, Gaussian and probit, scan form:
python tools/bench_ih.py [N]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bayesnewton_b200 as bn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
x = 0.2 * np.arange(N)
f = 2 * np.sin(0.3 * x) + np.sin(0.05 * x)
out = {'N': N}
for name, lik, y in (('gaussian', bn.likelihoods.Gaussian(0.5), f + 0.7 * rng.standard_normal(N)),
                     ('probit', bn.likelihoods.Bernoulli(), (f + 0.5 * rng.standard_normal(N) > 0).astype(np.float64))):
    m = bn.models.InfiniteHorizonVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=lik, X=x, Y=y, parallel=True)
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
