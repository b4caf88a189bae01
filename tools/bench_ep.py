"""EP iteration (inference + energy) at N = 1e7, fused (default) and through the stage-level kernels (BN_B200_FUSED=0):
python tools/bench_ep.py [N]"""
import json
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
if len(sys.argv) > 2:  # child: one configuration
    import torch
    import bayesnewton_b200 as bn
    from _data import bench_inputs
    t, dt, y = bench_inputs(N)
    m = bn.models.MarkovExpectationPropagationGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(),
                                                 X=t, Y=y, power=0.5, parallel=True)
    for _ in range(3):
        m.inference(lr=1.0); E = m.energy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.inference(lr=1.0); E = m.energy()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({'ms_per_iteration': e0.elapsed_time(e1) / 10, 'energy': float(E), 'fused': os.environ.get('BN_B200_FUSED', '1')}))
else:
    out = {'N': N}
    for fused in ('1', '0'):
        r = subprocess.run([sys.executable, __file__, str(N), 'child'], env=dict(os.environ, BN_B200_FUSED=fused),
                           capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith('{')]
        out['fused' if fused == '1' else 'stage_level'] = json.loads(line[-1]) if line else {'error': r.stderr[-400:]}
    if 'ms_per_iteration' in out.get('fused', {}) and 'ms_per_iteration' in out.get('stage_level', {}):
        out['speedup'] = out['stage_level']['ms_per_iteration'] / out['fused']['ms_per_iteration']
        out['energy_rel_diff'] = abs(out['fused']['energy'] - out['stage_level']['energy']) / abs(out['stage_level']['energy'])
    print(json.dumps(out))
