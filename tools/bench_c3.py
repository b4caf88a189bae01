"""C3 timing aid (not the bench line): Independent[Matern32 x2] + HeteroscedasticNoise, N = 10^6, one iteration =
inference(lr=0.3) + energy().  `python tools/bench_c3.py` = joint d=4 model on one GPU;
`torchrun --nproc-per-node 2 tools/bench_c3.py` = latents sharded over two GPUs (NCCL all-to-all)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bayesnewton_b200 as bn
from bayesnewton_b200 import _lib, latent_sharding as ls

world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
N = int(os.environ.get('C3_N', 1_000_000))
rng = np.random.default_rng(0)
dt = 0.05 + 0.1 * rng.random(N); dt[0] = 0
t = np.cumsum(dt)
y = np.sin(0.5 * t) + np.log1p(np.exp(np.cos(0.2 * t))) * np.random.default_rng(1).standard_normal(N)
y = (y - y.mean()) / y.std()
K = bn.kernels
out = {}
for name, mid in (('newton', _lib.BN_METHOD_NEWTON), ('ep', _lib.BN_METHOD_EP), ('vi', _lib.BN_METHOD_VI)):
    kern = K.Independent([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)])
    lik = bn.likelihoods.HeteroscedasticNoise()
    if world == 1 and os.environ.get('C3_JOINT', '1') == '1':
        cls = {'newton': bn.models.MarkovNewtonGP, 'ep': bn.models.MarkovExpectationPropagationGP, 'vi': bn.models.MarkovVariationalGP}[name]
        m = cls(kernel=kern, likelihood=lik, X=t, Y=y, parallel=True, **(dict(power=0.5) if name == 'ep' else {}))
        step = lambda: (m.inference(lr=0.3, return_state=False), m.energy())[1]
    else:
        m = ls.LatentShardedMarkovGP(kern, lik, t, y, mid, rank, world, power=0.5)
        step = lambda: (m.inference(lr=0.3), m.energy())[1]
    for _ in range(3):
        E = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        E = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[name] = {'ms_per_iter': float(ms), 'steps_per_s': N / (float(ms) * 1e-3), 'energy': float(E)}
if rank == 0:
    print(json.dumps({'config': 'C3 N=%d' % N, 'n_gpus': world, 'mode': 'joint d=4' if (world == 1 and os.environ.get('C3_JOINT', '1') == '1') else 'latent-sharded', 'results': out}))
if world > 1:
    dist.destroy_process_group()
