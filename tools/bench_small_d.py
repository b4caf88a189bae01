"""time of the warp-cooperative small-d filter + smoother (csrc/gd.cu) through bn_kf_arrays / bn_rts_arrays, scan form:
python tools/bench_small_d.py [N]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from bayesnewton_b200 import ops
from oracle import ssm
from test_small_d_generic import STACKS, smoother_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
out = {'N': N}
for name in ('m72x2_d8', 'm52x3_d9', 'm72x4_d16'):
    k = ssm.Independent(STACKS[name])
    rng = np.random.default_rng(0)
    dt = np.concatenate([[0.0], 0.05 + 0.3 * rng.random(N - 1)])
    uniq = np.stack([k.state_transition(h) for h in dt[:2000]])
    As = np.tile(uniq, (N // 2000 + 1, 1, 1))[:N].copy()
    Pinf = k.stationary_covariance()
    Qs = Pinf[None] - As @ Pinf[None] @ np.transpose(As, (0, 2, 1))
    H = k.measurement_model()
    d, D = H.shape[1], H.shape[0]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    ys, Rs = t(rng.standard_normal((N, D, 1))), t(np.tile(0.5 * np.eye(D)[None], (N, 1, 1)))
    A2, Q2 = smoother_inputs(As, Qs)
    As_d, Qs_d, H_d, m0, P0, A2, Q2 = t(As), t(Qs), t(H), t(np.zeros((d, 1))), t(Pinf), t(A2), t(Q2)
    res = {}
    for it in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ell, fm, fP = ops._parallel_kf(As_d, Qs_d, H_d, ys, Rs, m0, P0, None)
        e1.record()
        sm, sP, G = ops._parallel_rts(fm, fP, A2, Q2, H_d, False)
        e2.record()
        torch.cuda.synchronize()
        res = {'filter_ms': e0.elapsed_time(e1), 'smoother_ms': e1.elapsed_time(e2)}
    res['steps_per_s_filter_plus_smoother'] = N / ((res['filter_ms'] + res['smoother_ms']) * 1e-3)
    res['d'], res['D'] = d, D
    out[name] = res
print(json.dumps(out))
