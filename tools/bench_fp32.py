"""iteration time of the fused path in fp64 and fp32 (FusedShard): python tools/bench_fp32.py [N] [iters]
One iteration = site pass + energy pass (the two fused passes of inference() + energy()), CUDA-event timed."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib, fused
from _data import bench_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
t, dt, y = bench_inputs(N)
dev = torch.device('cuda')
dt_d, y_d = torch.as_tensor(dt, device=dev), torch.as_tensor(y, device=dev)
kern, lik = bn.kernels.Matern52(1.0, 1.0), bn.likelihoods.Bernoulli()
res = {}
post = {}
for name, dtype in (('f64', torch.float64), ('f32', torch.float32)):
    sh = fused.FusedShard(kern, dt_d, y_d, dtype=dtype)
    sh.load_sites(torch.zeros(N, device=dev), torch.full((N,), 100.0, device=dev))

    def step():
        sh.run(fused.SITES, lik, _lib.BN_METHOD_VI, None, 1.0, 1.0, True, want_ell=False)
        return sh.run(fused.ENERGY, lik, _lib.BN_METHOD_VI, None, 1.0, 1.0, True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ell, s = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    pm, pc = sh.posterior()
    post[name] = (pm.double().cpu().numpy().reshape(-1), pc.double().cpu().numpy().reshape(-1), float(ell))
    res[name] = {'ms_per_iteration': ms, 'steps_per_s': N / ms * 1e3, 'ell': float(ell)}
    del sh
    torch.cuda.empty_cache()
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
res['f32_vs_f64'] = {'post_mean_rel': rel(post['f32'][0], post['f64'][0]), 'post_var_rel': rel(post['f32'][1], post['f64'][1]),
                     'ell_rel': abs(post['f32'][2] - post['f64'][2]) / abs(post['f64'][2]),
                     'speedup': res['f64']['ms_per_iteration'] / res['f32']['ms_per_iteration']}
res['N'] = N
print(json.dumps(res))
